"""Shared helpers for the parity tests."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from eesen_b200 import kaldi_io, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def assert_close(name, got, ref, atol, rtol=0.0):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, f"{name}: shape {got.shape} vs {ref.shape}"
    err = np.abs(got - ref)
    bound = atol + rtol * np.abs(ref)
    worst = (err - bound).max() if err.size else 0.0
    assert worst <= 0, (f"{name}: max abs err {err.max():.3e} (|ref| max {np.abs(ref).max():.3e}) exceeds "
                        f"atol={atol:g} rtol={rtol:g}")


def model_file(net):
    d = tempfile.mkdtemp(prefix="eesen_b200_test_")
    p = os.path.join(d, "model")
    kaldi_io.write_model(p, net)
    return p


def case(wl, mseed=3, bseed=5, bidirectional=True):
    w = synth.WORKLOADS[wl]
    return w, synth.make_model(w, seed=mseed, bidirectional=bidirectional), synth.make_batch(w, seed=bseed)


def golden_arrays(dump, net):
    """Flatten the reference's grad_<layer>_<name> dumps into model-file (arena) order."""
    names = {"wx": "wx", "wm": "wm", "b": "b", "pi": "pi", "pf": "pf", "po": "po"}
    out = []
    for li, l in enumerate(net.layers):
        for n in l.param_names():
            out.append(np.asarray(dump[f"grad_{li}_{n}"]).ravel())
    return np.concatenate(out)
