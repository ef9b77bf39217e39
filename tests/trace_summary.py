"""Reads an EESEN_B200_TRACE_FILE (context.h:prof_collect) and prints the timeline of the LAST collected block: launches of
the main stream in order (start, duration, gap to the previous one), per-category busy time per stream, idle time."""
import sys
CATS = ("gemm", "lstm_fwd", "lstm_bwd", "softmax", "ctc", "sgd", "allreduce", "misc", "gemm_side")
blocks, cur = [], None
for ln in open(sys.argv[1]):
    if ln.startswith("#"):
        cur = []; blocks.append(cur); continue
    st, cat, t0, ms = ln.split()
    cur.append((int(st), int(cat), float(t0), float(ms)))
ev = blocks[-1]
end = max(t0 + ms for _, _, t0, ms in ev)
print(f"{len(ev)} launches, span {end:.3f} ms")
for stream in (0, 1):
    busy = {}
    for st, cat, t0, ms in ev:
        if st == stream:
            busy[CATS[cat]] = busy.get(CATS[cat], 0.0) + ms
    print(("main" if stream == 0 else "side"), "busy:", {k: round(v, 3) for k, v in busy.items()}, "sum", round(sum(busy.values()), 3))
verbose = len(sys.argv) > 2
prev_end, gaps = 0.0, 0.0
seg = []   # merge consecutive same-category launches of the main stream
for st, cat, t0, ms in ev:
    if st != 0:
        continue
    gap = t0 - prev_end
    if gap > 0:
        gaps += gap
    if seg and seg[-1][0] == cat and gap < 0.02:
        seg[-1][2] = t0 + ms; seg[-1][3] += 1; seg[-1][4] += ms
    else:
        seg.append([cat, t0, t0 + ms, 1, ms, gap])
    prev_end = max(prev_end, t0 + ms)
print(f"main stream: idle between launches {gaps:.3f} ms")
for cat, a, b, n, ms, gap in seg:
    print(f"  {a:8.3f} .. {b:8.3f}  {CATS[cat]:9s} x{n:<3d} busy {ms:7.3f}   (gap before {gap:6.3f})")
