"""First-light script (not a pytest file): level-2 train step on tiny/small vs the oracle."""
import sys, os, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eesen_b200 import binding, kaldi_io, synth
from oracle import oracle

def cmp(name, a, b, tol=None):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    err = np.abs(a - b).max() if a.size else 0.0
    print(f"  {name:12s} max|ref|={np.abs(b).max():.3e} maxabs err={err:.3e}", flush=True)
    return err

ctx = binding.Context(0)
print("SMs", ctx.sm_count, flush=True)
for wl in sys.argv[1:] or ["tiny", "small", "mid"]:
    w = synth.WORKLOADS[wl]
    net = synth.make_model(w, seed=3)
    b = synth.make_batch(w, seed=5)
    d = tempfile.mkdtemp()
    kaldi_io.write_model(d + "/model", net)
    lr, mom = 1e-3, 0.9
    on = oracle.OracleNet(net, np.float64)
    ro = on.train_step(b, lr, mom)
    n = binding.Net(ctx, d + "/model")
    n.set_train_options(lr, mom)
    n.get(102)  # request in_diff
    t0 = time.time()
    st = n.train_step(b.feats, b.frames, b.labels)
    print(wl, "step", time.time() - t0, "s", st, "oracle obj", ro["pzx"].sum(), flush=True)
    for i in range(len(net.layers) + 1):
        cmp(f"out_l{i}", n.get(i), on.acts[i])
    cmp("pzx", n.get(101).ravel(), ro["pzx"])
    cmp("obj_diff", n.get(100), ro["obj_diff"])
    cmp("in_diff", n.get(102), ro["in_diff"])
    cmp("corr", n.corr(), on.flat_corr())
    cmp("params", n.params(), on.flat_params())
    st = n.train_step(b.feats, b.frames, b.labels)
    ro = on.train_step(b, lr, mom)
    print(" step2", st, "oracle obj", ro["pzx"].sum())
    cmp("params2", n.params(), on.flat_params())
    n.close()
print("DONE")
