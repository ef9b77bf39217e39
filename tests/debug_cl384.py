"""Debug helper: one BiLSTM layer forward/backward with the cluster (DSMEM) exchange against the L2 exchange on the same
inputs; prints where d(gates) differ (time step, utterance, gate, cell / 32 = producing slice)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eesen_b200 import binding, kaldi_io

S, T, I, C = [int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (3, 9, 40, 384))]
rng = np.random.default_rng(S * 1000 + T)
frames = np.sort(rng.integers(max(1, T // 2), T + 1, size=S))[::-1].astype(np.int32); frames[0] = T
x = rng.standard_normal((T * S, I)).astype(np.float32)
for s in range(S):
    x[np.arange(frames[s], T) * S + s] = 0
l = kaldi_io.LayerSpec("bilstm", I, 2 * C)
params = [rng.uniform(-0.3, 0.3, size=l.param_shapes()[n]).astype(np.float32) for n in l.param_names()]
dout = rng.standard_normal((T * S, 2 * C)).astype(np.float32)
for s in range(S):
    dout[np.arange(frames[s], T) * S + s] = 0
res = {}
for ex in ("l2", "dsmem", "dsmem"):
    os.environ["EESEN_B200_LSTM_EXCHANGE"] = ex
    ctx = binding.Context(0)
    dp = [torch.from_numpy(p).cuda() for p in params]
    dg = [torch.zeros_like(p) for p in dp]
    d_x = torch.from_numpy(x).cuda(); d_len = torch.from_numpy(frames).cuda()
    gates = torch.zeros((T * S, 8 * C), device="cuda"); cell = torch.zeros((T * S, 2 * C), device="cuda")
    out = torch.zeros((T * S, 2 * C), device="cuda"); dgates = torch.zeros((T * S, 8 * C), device="cuda")
    d_dout = torch.from_numpy(dout).cuda(); dx = torch.zeros((T * S, I), device="cuda")
    torch.cuda.synchronize()
    ctx.bilstm_forward(T, S, I, C, d_len, d_x, I, dp, gates, cell, out, 2 * C)
    ctx.bilstm_backward(T, S, I, C, d_x, I, dp, gates, cell, out, 2 * C, d_dout, 2 * C, dgates, dx, I, dg)
    ctx.synchronize()
    r = {"out": out.cpu().numpy(), "dgates": dgates.cpu().numpy()}
    if "l2" in res and ex == "dsmem":
        for k in ("out", "dgates"):
            d = np.abs(r[k] - res["l2"][k])
            print(f"{k}: max |dsmem - l2| = {d.max():.3e}")
            if d.max() > 1e-4:
                rows, cols = np.nonzero(d > 1e-4)
                ts = sorted(set((rows // S).tolist())); us = sorted(set((rows % S).tolist()))
                ncol = r[k].shape[1] // 2
                dirs = sorted(set((cols // ncol).tolist()))
                cc = cols % ncol
                gate = sorted(set((cc // C).tolist())) if k == "dgates" else []
                sl = sorted(set(((cc % C) // 32).tolist()))
                print(f"   {len(rows)} elements; t in {ts}; utts {us}; dirs {dirs}; gates {gate}; slices (cell/32) {sl}")
    res[ex] = r
    ctx.close()
