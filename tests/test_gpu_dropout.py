"""GPU parity of the dropout variants of BiLstmParallel through the C ABI.
Chain of evidence: reference CPU build == restatement on the reference's own masks (tests/test_dropout_cpu.py,
committed fixtures); here: CUDA path == restatement (fp64) on the same masks, the forward pass additionally
straight against the reference's numbers in the fixtures, and the device mask generator against its
specification."""
import os
import subprocess

import numpy as np
import pytest

from util import ROOT, assert_close, model_file
from test_dropout_cpu import GOLDEN_CASES, VARIANTS, drop_case, golden_drop_case
from test_gpu_parity import diff_atol
from eesen_b200 import binding, kaldi_io
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = binding.Context(0)
    yield c
    c.close()


def _inject(n, masks):
    for li, m in enumerate(masks):
        if m:
            n.set_dropout_masks(li, m.get("fmask"), m.get("rmask"))


@pytest.mark.parametrize("wl,variant", GOLDEN_CASES)
def test_dropout_train_step_vs_oracle_on_reference_masks(ctx, wl, variant):
    w, net, b, masks, g = golden_drop_case(wl, variant)
    lr, mom = 1e-3, 0.9
    n = binding.Net(ctx, model_file(net))
    n.set_train_options(lr, mom)
    n.get(102)
    _inject(n, masks)
    st = n.train_step(b.feats, b.frames, b.labels, True)
    # forward pass straight against the reference's CPU numbers (same masks)
    assert_close("last BiLSTM output vs reference", n.get(w.layers), g[f"out_l{w.layers}"], atol=2e-5)
    on = oracle.OracleNet(net, np.float64)
    ro = on.train_step(b, lr, mom, masks=masks)
    for i in range(1, len(net.layers) + 1):
        assert_close(f"out_l{i}", n.get(i), on.acts[i], atol=2e-5)
    assert_close("pzx", n.get(101).ravel(), ro["pzx"], atol=0, rtol=2e-5)
    assert_close("obj_diff", n.get(100), ro["obj_diff"], atol=diff_atol(ro["pzx"]))
    assert_close("in_diff", n.get(102), ro["in_diff"], atol=2e-5, rtol=1e-3)
    assert_close("corr", n.corr(), on.flat_corr(), atol=2e-3, rtol=2e-3)
    assert_close("params", n.params(), on.flat_params(), atol=2e-6)
    # test mode (cross-validation): no dropout at all -- identical to the model without the options
    st0 = n.train_step(b.feats, b.frames, b.labels, False)
    plain = kaldi_io.read_model(model_file(net))
    for l in plain.layers:
        l.dropout = {}
    plain.set_flat_params(n.params())
    n2 = binding.Net(ctx, model_file(plain))
    st1 = n2.train_step(b.feats, b.frames, b.labels, False)
    assert st0["obj"] == st1["obj"]
    n.close(); n2.close()


@pytest.mark.parametrize("variant", ["fwdstep", "fwdseq", "nmlstep", "rnndropseq", "fwd+nml"])
def test_generated_masks_follow_the_specification(ctx, variant):
    w, net, b = drop_case("small", variant)
    lr, mom = 1e-3, 0.9
    opts = VARIANTS[variant]
    N, C2 = b.feats.shape[0], 2 * w.cells

    def run(seed, steps=1):
        n = binding.Net(ctx, model_file(net))
        n.set_train_options(lr, mom)
        n.set_dropout_seed(seed)
        for _ in range(steps):
            n.train_step(b.feats, b.frames, b.labels, True)
        return n

    n = run(1234)
    masks = []
    for li, l in enumerate(net.layers):
        if l.kind != "bilstm":
            masks.append(None)
            continue
        m, draw = {}, 0
        if "recurrent" in opts:    # drawn before the recurrence (InitializeRecurrentMasks), forward mask after it
            rows = N if opts.get("rec_step") else b.S
            m["rmask"] = n.get(400 + li)
            assert m["rmask"].shape == (rows, C2)
            spec = oracle.dropout_mask(rows, C2, opts["recurrent"], bool(opts.get("rec_seq")), 1234, ((li + 1) << 32) + draw)
            assert np.array_equal(m["rmask"], spec)
            draw += 1
        if "forward" in opts:
            m["fmask"] = n.get(300 + li)
            assert m["fmask"].shape == (N, C2)
            spec = oracle.dropout_mask(N, C2, opts["forward"], bool(opts.get("fw_seq")), 1234, ((li + 1) << 32) + draw)
            assert np.array_equal(m["fmask"], spec)
        masks.append(m)
        for key, p in (("fmask", opts.get("forward")), ("rmask", opts.get("recurrent"))):
            if key in m:
                seq = opts.get("fw_seq") if key == "fmask" else opts.get("rec_seq")
                nd = m[key].shape[1] if seq else m[key].size
                assert abs((m[key] == 0).mean() - p) < 5.0 * np.sqrt(p * (1 - p) / nd) + 0.01
                assert set(np.unique(m[key])) <= {np.float32(0.0), np.float32(1.0) / (np.float32(1.0) - np.float32(p))}
                if seq:
                    assert np.all(m[key] == m[key][0])
    # end to end on the generated masks: same numbers as the restatement fed with them
    on = oracle.OracleNet(net, np.float64)
    on.train_step(b, lr, mom, masks=masks)
    assert_close("params", n.params(), on.flat_params(), atol=2e-6)
    # same seed -> same run; another seed or another step -> other masks
    n_same, n_other, n_two = run(1234), run(99), run(1234, steps=2)
    assert np.array_equal(n.params(), n_same.params())
    assert not np.array_equal(n.params(), n_other.params())
    key = 300 if "forward" in opts else 400
    assert not np.array_equal(n.get(key), n_two.get(key))
    for x in (n, n_same, n_other, n_two):
        x.close()


def test_change_dropout_checks_tool_and_twiddle(ctx, tmp_path):
    w, net, b = drop_case("tiny", "fwd+nml")
    plain = kaldi_io.read_model(model_file(net))
    for l in plain.layers:
        l.dropout = {}
    d = str(tmp_path)
    kaldi_io.write_model(d + "/plain", plain)
    kaldi_io.write_model(d + "/want", net)
    # the tool writes the same bytes as a model that carries the options
    r = subprocess.run([os.path.join(ROOT, "eesen_b200", "bin", "net-change-model"), "--forwarddrop=0.2", "--forwardstep=true",
                        "--nmldrop=true", "--recurrentdrop=0.25", "--recurrentseq=true", d + "/plain", d + "/got"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(d + "/got", "rb").read() == open(d + "/want", "rb").read()
    n = binding.Net(ctx, d + "/plain")
    for bad in (dict(forward=0.2), dict(fw_step=True), dict(forward=0.2, fw_step=True, fw_seq=True),
                dict(recurrent=0.2, rec_step=True, rnndrop=True, nml=True), dict(rnndrop=True, rec_step=True),
                dict(recurrent=0.2, rnndrop=True), dict(recurrent=0.2, rec_step=True, rec_seq=True, nml=True)):
        with pytest.raises(binding.EesenB200Error):     # bilstm-layer.h:74-112
            n.change_dropout(**bad)
    # TwiddleForward: every minibatch uses exactly one of the two kinds; deterministic in the seed
    n.change_dropout(forward=0.2, fw_step=True, recurrent=0.25, rec_seq=True, nml=True, twiddle=True)
    n.set_train_options(1e-3, 0.9)
    n.set_dropout_seed(7)
    for _ in range(6):
        n.train_step(b.feats, b.frames, b.labels, True)
    p1 = n.params()
    n2 = binding.Net(ctx, d + "/plain")
    n2.change_dropout(forward=0.2, fw_step=True, recurrent=0.25, rec_seq=True, nml=True, twiddle=True)
    n2.set_train_options(1e-3, 0.9)
    n2.set_dropout_seed(7)
    for _ in range(6):
        n2.train_step(b.feats, b.frames, b.labels, True)
    assert np.array_equal(p1, n2.params()) and np.all(np.isfinite(p1))
    n.close(); n2.close()
