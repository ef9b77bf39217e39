"""Relative-to-norm check of every gradient tensor after one training step of `mid` (2x320 BiLSTM, 16 utterances) against
the fp64 oracle: ||g_gpu - g_fp64|| / ||g_fp64|| per parameter tensor.

The element-wise bounds of test_gpu_parity.py (abs 2e-3 + rel 2e-3 on the momentum buffers) would not notice a small
systematic bias of a whole tensor; a norm-relative bound does.  Scale of the bound: the fp32 build of the oracle itself sits
3e-6 .. 1.3e-5 from its fp64 build on these tensors (summation order, fp32 CTC; measured with
tests/test_zz_gradient_norms.py::test_fp32_oracle_norm_distance_is_the_yardstick, no GPU needed), and the fp16x3 dense products
add ~1e-5 (profiles/r02_gemm_shapes_fp16x3.txt); 5e-4 leaves an order of magnitude for both and still flags a 1e-3 relative bias.
(Runs last -- file name -- so that a failure here cannot hide another test under `pytest -x`.)"""
import numpy as np
import pytest

from util import case, model_file
from oracle import oracle

LR, MOM = 1e-3, 0.9


def _tensors(net):
    """(layer index, name, size) of every parameter tensor in arena order."""
    out = []
    for li, l in enumerate(net.layers):
        try:
            names, shapes = l.param_names(), l.param_shapes()
        except Exception:
            continue
        for n in names:
            out.append((li, n, int(np.prod(shapes[n]))))
    return out


def _rel_norms(got, ref, net):
    off, res = 0, []
    for li, name, n in _tensors(net):
        a, r = got[off:off + n].astype(np.float64), ref[off:off + n].astype(np.float64)
        res.append((li, name, np.linalg.norm(a - r) / max(np.linalg.norm(r), 1e-30)))
        off += n
    assert off == len(ref)
    return res


def test_fp32_oracle_norm_distance_is_the_yardstick():
    w, net, b = case("mid")
    o64 = oracle.OracleNet(net, np.float64); o64.train_step(b, LR, MOM)
    o32 = oracle.OracleNet(net, np.float32); o32.train_step(b, LR, MOM)
    rel = _rel_norms(o32.flat_corr(), o64.flat_corr(), net)
    worst = max(r for _, _, r in rel)
    assert 1e-7 < worst < 5e-5, rel          # fp32 arithmetic alone: ~1e-5


@pytest.mark.gpu
def test_gradient_tensors_relative_to_norm_vs_fp64_oracle():
    from eesen_b200 import binding
    w, net, b = case("mid")
    ctx = binding.Context(0)
    n = binding.Net(ctx, model_file(net))
    n.set_train_options(LR, MOM)
    n.train_step(b.feats, b.frames, b.labels, True)
    o64 = oracle.OracleNet(net, np.float64); o64.train_step(b, LR, MOM)
    rel = _rel_norms(n.corr(), o64.flat_corr(), net)      # momentum buffer after the first step = the raw gradient
    n.close(); ctx.close()
    bad = [(li, name, r) for li, name, r in rel if not r <= 5e-4]
    print("gradient ||d||/||g|| per tensor: max %.2e, median %.2e" % (max(r for _, _, r in rel), float(np.median([r for _, _, r in rel]))))
    assert not bad, bad
