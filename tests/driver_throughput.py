"""N4 measurement (run on a GPU box): sustained rate of the C++ driver eesen_b200/bin/train-ctc-parallel on C2-shaped
Kaldi archives -- ark/scp reading, minibatch assembly on the producer thread, pinned staging, H2D, the step -- against
the rate of the device-resident loop of bench.py.  Reference loop: src/netbin/train-ctc-parallel.cc:144-218.

    python tests/driver_throughput.py [num_utts]      -> one JSON line
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eesen_b200 import kaldi_io, synth  # noqa: E402


def main():
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    w = synth.WORKLOADS["c2"]
    rng = np.random.default_rng(7)
    d = tempfile.mkdtemp(prefix="eesen_b200_drv_")
    lens = np.sort(rng.integers(w.t_lo, w.t_hi + 1, size=n_utts))[::-1]      # sorted by length, as the recipes do
    keys = [f"utt{i:06d}" for i in range(n_utts)]
    utts = [rng.standard_normal((int(t), w.in_dim)).astype(np.float32) for t in lens]
    labels = [rng.integers(1, w.classes, size=int(rng.integers(w.lab_lo, w.lab_hi + 1))).astype(np.int32) for _ in lens]
    kaldi_io.write_feature_ark(os.path.join(d, "feats.ark"), keys, utts)
    kaldi_io.write_label_ark(os.path.join(d, "labels.ark"), keys, labels)
    kaldi_io.write_model(os.path.join(d, "nnet.in"), synth.make_model(w, seed=0))
    valid = int(lens.sum())
    exe = os.path.join(ROOT, "eesen_b200", "bin", "train-ctc-parallel")
    cmd = [exe, f"--learn-rate={w.learn_rate}", f"--momentum={w.momentum}", f"--num-sequence={w.S}", "--frame-limit=1000000",
           "--report-step=1000", f"ark:{d}/feats.ark", f"ark,t:{d}/labels.ark", f"{d}/nnet.in", f"{d}/nnet.out"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        print(r.stderr[-3000:])
        sys.exit(1)
    m = re.search(r"([0-9.eE+-]+) min, fps([0-9.eE+-]+)\]", r.stderr)
    drv_min, drv_fps = float(m.group(1)), float(m.group(2))
    print(json.dumps({"driver": "eesen_b200/bin/train-ctc-parallel", "utterances": n_utts, "minibatches": (n_utts + w.S - 1) // w.S,
                      "valid_frames": valid, "driver_reported_padded_fps": drv_fps, "driver_reported_minutes": drv_min,
                      "valid_frames_per_sec_training_loop": valid / (drv_min * 60.0),
                      "wall_seconds_incl_process_start_and_model_io": wall,
                      "archive_mb": os.path.getsize(os.path.join(d, "feats.ark")) / 1e6}))


if __name__ == "__main__":
    main()
