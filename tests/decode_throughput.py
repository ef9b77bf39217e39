"""Measurement helper (one GPU, plain python): one-best WFST search (BASELINE config 5) of 64
utterances over a synthetic ~10 M-arc CTC-topology lexicon graph built without OpenFst (eesen_b200/wfst.py).

    python tests/decode_throughput.py [words=400000] [T=300] [beam=7.0]

Prints frames/s over the batch, device ms per frame step, epsilon-closure rounds per frame and the graph's footprint in
HBM (the search is HBM/L2-latency-bound integer work: per expanded arc 16 bytes of arc record + one 8-byte atomicMin)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from eesen_b200 import binding, wfst  # noqa: E402

words = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
beam = float(sys.argv[3]) if len(sys.argv) > 3 else 7.0
K, S = 46, 64
t0 = time.time()
g = wfst.synthetic_tlg(2, words, K)
t_build = time.time() - t0
rng = np.random.default_rng(5)
frames = np.sort(rng.integers(T * 2 // 3, T + 1, size=S))[::-1].copy()
frames[0] = T
ll = np.log(rng.dirichlet(np.ones(K) * 0.15, size=T * S)).astype(np.float32)
ld = (K + 3) // 4 * 4
pad = np.zeros((T * S, ld), np.float32)
pad[:, :K] = ll
d_ll = torch.from_numpy(pad).cuda()
ctx = binding.Context(0)
dg = wfst.DeviceGraph(ctx, g)
out = None
for rep in range(2):                       # first call: allocation of the workspace
    torch.cuda.synchronize()
    t0 = time.time()
    wordsq, cost, st = dg.decode(d_ll, ld, K, frames, T, 1.0, beam, frame_cap=1 << 17, tok_cap=1 << 24)
    wall = time.time() - t0
    out = {"graph": {"states": int(g.num_states), "arcs": g.num_arcs, "bytes": g.num_arcs * 16 + g.num_states * 12,
                     "build_s": round(t_build, 1)},
           "utterances": S, "T": T, "beam": beam, "frames": int(frames.sum()),
           "device_ms": st["device_ms"], "wall_ms": wall * 1e3, "frames_per_s": float(frames.sum()) / (st["device_ms"] * 1e-3),
           "ms_per_frame_step": st["device_ms"] / T, "closure_rounds_per_frame": st["closure_rounds"] / T,
           "decoded_nonempty": int(sum(bool(w) for w in wordsq))}
print(json.dumps(out))
dg.close()
