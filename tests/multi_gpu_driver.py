"""2-GPU check of the C++ driver with UNEVEN shards (run on a 2-GPU box, plain python):

    python tests/multi_gpu_driver.py

Job 1 holds 8 utterances (two minibatches of 4), job 2 holds 4 (one minibatch).  With
--num-jobs=2 the drivers all-reduce the gradient every step; in step 2 job 2 has no data and votes
"idle" (Net::BackpropagateShared) instead of dead-locking.  The written model must equal a
single-GPU run through the level-2 API over [A0-3 + B0-3] then [A4-7] (gradients are sums over rows).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eesen_b200 import binding, kaldi_io, synth  # noqa: E402


def main():
    w = synth.WORKLOADS["small"]
    net = synth.make_model(w, seed=3)
    a = synth.make_batch(w, seed=41)          # 8 utterances
    b = synth.make_batch(w, seed=42, S=4)     # 4 utterances
    d = tempfile.mkdtemp()
    model = os.path.join(d, "nnet.in")
    kaldi_io.write_model(model, net)

    def utts(batch):
        return [batch.feats[np.arange(batch.frames[s]) * batch.S + s] for s in range(batch.S)]

    ua, ub = utts(a), utts(b)
    for name, us, labs in (("A", ua, a.labels), ("B", ub, b.labels)):
        keys = [f"{name}{i:02d}" for i in range(len(us))]
        kaldi_io.write_feature_ark(os.path.join(d, f"feats{name}.ark"), keys, us)
        kaldi_io.write_label_ark(os.path.join(d, f"labels{name}.ark"), keys, labs)
    out = os.path.join(d, "nnet.out")
    exe = os.path.join(ROOT, "eesen_b200", "bin", "train-ctc-parallel")
    procs = []
    for job, name in ((1, "A"), (2, "B")):
        cmd = [exe, "--learn-rate=0.001", "--momentum=0.9", "--num-sequence=4", "--num-jobs=2", f"--job-id={job}",
               f"ark:{d}/feats{name}.ark", f"ark,t:{d}/labels{name}.ark", model, out]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    ok = True
    for p in procs:
        try:
            so, se = p.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            p.kill()
            print("driver timed out (dead-lock?)")
            ok = False
            continue
        if p.returncode != 0:
            print("driver failed:", se[-2000:])
            ok = False
    if ok:
        got = kaldi_io.read_model(out).flat_params()
        ctx = binding.Context(0)
        n = binding.Net(ctx, model)
        n.set_train_options(1e-3, 0.9)
        f1, fr1 = kaldi_io.pack_utterances(ua[:4] + ub[:4])
        n.train_step(f1, fr1, list(a.labels[:4]) + list(b.labels[:4]), True)
        f2, fr2 = kaldi_io.pack_utterances(ua[4:8])
        n.train_step(f2, fr2, list(a.labels[4:8]), True)
        err = np.abs(got - n.params()).max()
        print(f"2-job driver (8 + 4 utterances) vs single-GPU equivalent: max |dparam| = {err:.3e}")
        ok = err < 2e-6
    print("MULTI_GPU_DRIVER", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
