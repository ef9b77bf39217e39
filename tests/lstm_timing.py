"""Debug helper (needs `make TIMING=1`): per-phase clock64 breakdown of the recurrent kernels on C2."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eesen_b200 import binding, synth
from util import model_file
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32x3"
w = synth.WORKLOADS["c2"]
ctx = binding.Context(0, prec, prec)
net = binding.Net(ctx, model_file(synth.make_model(w, seed=0)))
net.set_train_options(w.learn_rate, w.momentum)
b = synth.make_batch(w, seed=1)
for _ in range(2):
    net.train_step(b.feats, b.frames, b.labels, True)
buf = (C.c_longlong * 32)()
assert ctx.lib.eesen_b200_debug_lstm_timing(ctx.h, buf, 1) == 1, "not built with TIMING=1"
net.train_step(b.feats, b.frames, b.labels, True)
ctx.lib.eesen_b200_debug_lstm_timing(ctx.h, buf, 0)
steps = (b.T - 1) * w.layers
if os.environ.get("EESEN_B200_LSTM_ENGINE") == "legacy":
    names_f = ["poll", "barA", "stage_ld", "barB", "mma", "scratch+barD", "elementwise+publish", "signal", "gate_stores+prefetch", "scratch reduce"]
    names_b = ["poll", "barA", "partials+elementwise", "barC", "mma+P stores", "signal"]
else:   # tcgen05 engine (lstm_tc.cu)
    names_f = ["stage (poll + B tile) + arrive", "MMA issue + commit wait", "TMEM ld + staging + barrier", "gates + publish", "saved-state stores + prefetch", "(re-polls of thread 0, count)"]
    names_b = ["gather partial d_m", "gate math + DG stores + B tile + arrive", "MMA issue", "TMEM ld + publish partials", "prefetch issue", "(re-polls of thread 0, count)", "waiting for the tiles' commits", "(unused)", "(unused)", "(unused)", "waiting for the other warps (b_full)", "(unused)", "(warp 4: cycles from MMA issue start to commit 0 seen)", "(warp 4: ... commit 1 seen)", "(warp 4: ... commit 2 seen)", "(warp 4: ... commit 3 seen)"]
print(f"forward ({prec}) cycles/step (thread 0 of CTA 0; {steps} steps):")
tot = 0
for i, n in enumerate(names_f):
    if n.startswith("("):
        print(f"  {n:24s} {buf[i] / steps:9.2f}"); continue
    print(f"  {n:24s} {buf[i] / steps:9.0f}"); tot += buf[i] / steps
print(f"  {'total':24s} {tot:9.0f}")
print("backward, second observer (warp 4, both utterance halves of its quadrant): cycles/step in commit waits / TMEM loads (+ hand-over barrier) / arithmetic / pushes:")
print("  ", "  ".join(f"{buf[12 + i] / steps:8.0f}" for i in range(4)))
print("backward cycles/step:")
tot = 0
for i, n in enumerate(names_b):
    if n.startswith("("):
        print(f"  {n:24s} {buf[16 + i] / steps:9.2f}"); continue
    print(f"  {n:24s} {buf[16 + i] / steps:9.0f}"); tot += buf[16 + i] / steps
print(f"  {'total':24s} {tot:9.0f}")
