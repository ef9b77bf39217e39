"""Kaldi/Eesen on-disk formats at the boundary of the hot path (numpy, host side).

Byte-compatible with the reference so models and feature archives can be
exchanged with stock Eesen tools:

* binary model  -- ``Net::Write`` (net.cc:325-334), ``Layer::Write`` (layer.cc:209-222),
  ``BiLstm::WriteData`` (bilstm-layer.h:429-493), ``AffineTransform::WriteData``
  (affine-trans-layer.h:117-134); tokens ``WriteToken`` / ``WriteBasicType``
  (base/io-funcs-inl.h:32-60), matrices ``FM``/``FV`` (cpucompute/matrix.cc:968-1010).
* feature ark   -- ``key \\0B FM \\4 rows \\4 cols data`` per utterance.
* label ark (text) -- ``key l1 l2 ...\\n``.
* packed-batch file for the oracle driver (our own format, see oracle/dump_ref.cc).

This module is plumbing for tests/bench; the product's model I/O is the C++
host library (eesen_b200/host), which implements the same formats.
"""
from __future__ import annotations

import io
import struct
from dataclasses import dataclass, field
from typing import BinaryIO, Dict, List, Sequence, Tuple

import numpy as np

BILSTM_TENSORS = ("wx", "wm", "b", "pi", "pf", "po")  # per direction, WriteData order


@dataclass
class LayerSpec:
    kind: str  # "bilstm" | "lstm" (uni-directional) | "affine" | "softmax"
    in_dim: int
    out_dim: int  # <CellDim>: 2 * cells-per-direction for bilstm, the cell count for lstm
    learn_rate_coef: float = 1.0
    max_grad: float = 0.0
    params: Dict[str, np.ndarray] = field(default_factory=dict)
    # Adagrad/RMSProp accumulators (<BiLstmAccus>/<AffineAccus>, bilstm-layer.h:375-395,
    # affine-trans-layer.h:98-106); empty when the model carries none
    accus: Dict[str, np.ndarray] = field(default_factory=dict)
    # dropout options of <BiLstmParallel> (bilstm-layer.h:62-135): forward (float), fw_step, fw_seq, recurrent (float),
    # rec_step, rec_seq, rnndrop, nml, twiddle; missing keys = off
    dropout: Dict[str, float] = field(default_factory=dict)

    @property
    def cells(self) -> int:
        assert self.kind in ("bilstm", "lstm")
        return self.out_dim // 2 if self.kind == "bilstm" else self.out_dim

    def param_names(self) -> List[str]:
        if self.kind == "bilstm":
            return [f"{n}_{d}" for d in ("fw", "bw") for n in BILSTM_TENSORS]
        if self.kind == "lstm":
            return list(BILSTM_TENSORS)
        if self.kind == "affine":
            return ["w", "b"]
        return []

    def param_shapes(self) -> Dict[str, Tuple[int, ...]]:
        if self.kind == "bilstm":
            c, i = self.cells, self.in_dim
            one = {"wx": (4 * c, i), "wm": (4 * c, c), "b": (4 * c,), "pi": (c,), "pf": (c,), "po": (c,)}
            return {f"{n}_{d}": one[n] for d in ("fw", "bw") for n in BILSTM_TENSORS}
        if self.kind == "lstm":
            c, i = self.cells, self.in_dim
            return {"wx": (4 * c, i), "wm": (4 * c, c), "b": (4 * c,), "pi": (c,), "pf": (c,), "po": (c,)}
        if self.kind == "affine":
            return {"w": (self.out_dim, self.in_dim), "b": (self.out_dim,)}
        return {}


@dataclass
class NetSpec:
    layers: List[LayerSpec]

    @property
    def in_dim(self) -> int:
        return self.layers[0].in_dim

    @property
    def out_dim(self) -> int:
        return self.layers[-1].out_dim

    def num_params(self) -> int:
        return sum(int(np.prod(s)) for l in self.layers for s in l.param_shapes().values())

    def flat_params(self) -> np.ndarray:
        """All parameters concatenated in model-file order (the gradient-arena order)."""
        out = [l.params[n].astype(np.float32).ravel() for l in self.layers for n in l.param_names()]
        return np.concatenate(out) if out else np.zeros(0, np.float32)

    def flat_accus(self) -> np.ndarray:
        """Accumulators in the same order; zeros for layers without any."""
        out = []
        for l in self.layers:
            shapes = l.param_shapes()
            for n in l.param_names():
                out.append(l.accus[n].astype(np.float32).ravel() if l.accus else np.zeros(int(np.prod(shapes[n])), np.float32))
        return np.concatenate(out) if out else np.zeros(0, np.float32)

    def set_flat_params(self, flat: np.ndarray) -> None:
        off = 0
        for l in self.layers:
            shapes = l.param_shapes()
            for n in l.param_names():
                k = int(np.prod(shapes[n]))
                l.params[n] = flat[off:off + k].reshape(shapes[n]).astype(np.float32).copy()
                off += k
        assert off == flat.size


def make_net(in_dim: int, cells: int, num_layers: int, num_classes: int, seed: int = 0,
             param_range: float = 0.1, max_grad: float = 50.0, learn_rate_coef: float = 1.0,
             bidirectional: bool = True) -> NetSpec:
    """Random-init BiLSTM stack + affine + softmax, uniform(-range, range) like
    ``InitRandUniform`` (bilstm-layer.h:187-210); proto shape per
    asr_egs/wsj/utils/model_topo.py:80-95."""
    rng = np.random.default_rng(seed)
    layers: List[LayerSpec] = []
    d = in_dim
    for _ in range(num_layers):
        l = LayerSpec("bilstm", d, 2 * cells, learn_rate_coef, max_grad) if bidirectional else \
            LayerSpec("lstm", d, cells, learn_rate_coef, max_grad)
        for n, shp in l.param_shapes().items():
            l.params[n] = rng.uniform(-param_range, param_range, size=shp).astype(np.float32)
        layers.append(l)
        d = l.out_dim
    a = LayerSpec("affine", d, num_classes, learn_rate_coef, max_grad)
    for n, shp in a.param_shapes().items():
        a.params[n] = rng.uniform(-param_range, param_range, size=shp).astype(np.float32)
    layers.append(a)
    layers.append(LayerSpec("softmax", num_classes, num_classes))
    return NetSpec(layers)


# ----------------------------------------------------------------------------- binary tokens
def _wtok(f: BinaryIO, tok: str) -> None:
    f.write(tok.encode() + b" ")


def _wi32(f: BinaryIO, v: int) -> None:
    f.write(b"\x04" + struct.pack("<i", v))


def _wf32(f: BinaryIO, v: float) -> None:
    f.write(b"\x04" + struct.pack("<f", v))


def _wbool(f: BinaryIO, v: bool) -> None:
    f.write(b"T" if v else b"F")


def _wmat(f: BinaryIO, m: np.ndarray) -> None:
    m = np.ascontiguousarray(m, dtype="<f4")
    if m.ndim == 2:
        _wtok(f, "FM")
        _wi32(f, m.shape[0])
        _wi32(f, m.shape[1])
    else:
        _wtok(f, "FV")
        _wi32(f, m.shape[0])
    f.write(m.tobytes())


_MARKER = {"bilstm": "<BiLstmParallel>", "lstm": "<LstmParallel>", "affine": "<AffineTransform>", "softmax": "<Softmax>"}
_BILSTM_FLAGS = ("<ForwardTimeStepDropout>", "<ForwardSequenceDropout>", "<RecurrentTimeStepDropout>",
                 "<RecurrentSequenceDropout>", "<RNNDrop>", "<NoMemLossDropout>")


def write_model(path_or_file, net: NetSpec) -> None:
    f = open(path_or_file, "wb") if isinstance(path_or_file, str) else path_or_file
    f.write(b"\0B")
    _wtok(f, "<Nnet>")
    for l in net.layers:
        _wtok(f, _MARKER[l.kind])
        _wtok(f, "<InputDim>"); _wi32(f, l.in_dim)
        _wtok(f, "<CellDim>" if l.kind in ("bilstm", "lstm") else "<OutputDim>"); _wi32(f, l.out_dim)
        if l.kind == "bilstm":
            _wtok(f, "<LearnRateCoef>"); _wf32(f, l.learn_rate_coef)
            _wtok(f, "<MaxGrad>"); _wf32(f, l.max_grad)
            d = l.dropout
            _wtok(f, "<ForwardDropoutFactor>"); _wf32(f, float(d.get("forward", 0.0)))
            for t, key in zip(_BILSTM_FLAGS[:4], ("fw_step", "fw_seq", "rec_step", "rec_seq")):
                _wtok(f, t); _wbool(f, bool(d.get(key, False)))
            _wtok(f, "<RNNDrop>"); _wbool(f, bool(d.get("rnndrop", False)))
            _wtok(f, "<NoMemLossDropout>"); _wbool(f, bool(d.get("nml", False)))
            _wtok(f, "<RecurrentDropoutFactor>"); _wf32(f, float(d.get("recurrent", 0.0)))
            _wtok(f, "<TwiddleForward>"); _wbool(f, bool(d.get("twiddle", False)))
            if l.accus:
                _wtok(f, "<BiLstmAccus>")
                for n in l.param_names():
                    _wmat(f, l.accus[n])
            for n in l.param_names():
                _wmat(f, l.params[n])
        elif l.kind == "lstm":   # lstm-layer.h:147-172
            _wtok(f, "<LearnRateCoef>"); _wf32(f, l.learn_rate_coef)
            _wtok(f, "<MaxGrad>"); _wf32(f, l.max_grad)
            if l.accus:
                _wtok(f, "<LstmAccus>")
                for n in l.param_names():
                    _wmat(f, l.accus[n])
            for n in l.param_names():
                _wmat(f, l.params[n])
        elif l.kind == "affine":
            _wtok(f, "<LearnRateCoef>"); _wf32(f, l.learn_rate_coef)
            _wtok(f, "<MaxGrad>"); _wf32(f, l.max_grad)
            if l.accus:
                _wtok(f, "<AffineAccus>")
                _wmat(f, l.accus["w"]); _wmat(f, l.accus["b"])
            _wmat(f, l.params["w"]); _wmat(f, l.params["b"])
    _wtok(f, "</Nnet>")
    if isinstance(path_or_file, str):
        f.close()


class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        self.p = 0

    def peek(self) -> int:
        return self.d[self.p] if self.p < len(self.d) else -1

    def tok(self) -> str:
        e = self.d.index(b" ", self.p)
        t = self.d[self.p:e].decode()
        self.p = e + 1
        return t

    def expect(self, t: str) -> None:
        got = self.tok()
        if got != t:
            raise ValueError(f"expected token {t}, got {got}")

    def i32(self) -> int:
        assert self.d[self.p] == 4
        v = struct.unpack_from("<i", self.d, self.p + 1)[0]
        self.p += 5
        return v

    def f32(self) -> float:
        assert self.d[self.p] == 4
        v = struct.unpack_from("<f", self.d, self.p + 1)[0]
        self.p += 5
        return v

    def boolean(self) -> bool:
        c = self.d[self.p:self.p + 1]
        self.p += 1
        if self.peek() == 0x20:  # WriteBasicType<bool> writes 'T'/'F' followed by a space in some versions
            self.p += 1
        return c == b"T"

    def mat(self) -> np.ndarray:
        t = self.tok()
        if t == "FM":
            r, c = self.i32(), self.i32()
            n = r * c
            m = np.frombuffer(self.d, "<f4", n, self.p).reshape(r, c).copy()
        elif t == "FV":
            n = self.i32()
            m = np.frombuffer(self.d, "<f4", n, self.p).copy()
        else:
            raise ValueError(f"bad matrix token {t}")
        self.p += 4 * n
        return m


def read_model(path: str) -> NetSpec:
    data = open(path, "rb").read()
    if data[:2] != b"\0B":
        raise ValueError("only binary models are supported by this reader")
    r = _Reader(data)
    r.p = 2
    layers: List[LayerSpec] = []
    inv = {v: k for k, v in _MARKER.items()}
    inv["<BiLstm>"] = "bilstm"
    inv["<Lstm>"] = "lstm"
    while r.peek() != -1:
        t = r.tok()
        if t == "</Nnet>":
            break
        if t == "<Nnet>":
            t = r.tok()
        kind = inv[t]
        r.expect("<InputDim>"); i = r.i32()
        r.expect("<CellDim>" if kind in ("bilstm", "lstm") else "<OutputDim>"); o = r.i32()
        l = LayerSpec(kind, i, o)
        if kind == "bilstm":
            while r.peek() == ord("<"):
                tk = r.tok()
                if tk == "<LearnRateCoef>": l.learn_rate_coef = r.f32()
                elif tk == "<MaxGrad>": l.max_grad = r.f32()
                elif tk == "<ForwardDropoutFactor>": l.dropout["forward"] = r.f32()
                elif tk == "<RecurrentDropoutFactor>": l.dropout["recurrent"] = r.f32()
                elif tk in _BILSTM_FLAGS or tk == "<TwiddleForward>":
                    key = {"<ForwardTimeStepDropout>": "fw_step", "<ForwardSequenceDropout>": "fw_seq",
                           "<RecurrentTimeStepDropout>": "rec_step", "<RecurrentSequenceDropout>": "rec_seq",
                           "<RNNDrop>": "rnndrop", "<NoMemLossDropout>": "nml", "<TwiddleForward>": "twiddle"}[tk]
                    l.dropout[key] = r.boolean()
                elif tk == "<BiLstmAccus>":
                    for n in l.param_names():
                        l.accus[n] = r.mat()
                    break
                else: raise ValueError(f"unsupported token {tk}")
            for n in l.param_names():
                l.params[n] = r.mat()
        elif kind == "lstm":
            while r.peek() == ord("<"):
                tk = r.tok()
                if tk == "<LearnRateCoef>": l.learn_rate_coef = r.f32()
                elif tk == "<MaxGrad>": l.max_grad = r.f32()
                elif tk == "<LstmAccus>":
                    for n in l.param_names():
                        l.accus[n] = r.mat()
                    break
                else: raise ValueError(f"unsupported token {tk}")
            for n in l.param_names():
                l.params[n] = r.mat()
        elif kind == "affine":
            while r.peek() == ord("<"):
                tk = r.tok()
                if tk == "<LearnRateCoef>": l.learn_rate_coef = r.f32()
                elif tk == "<MaxGrad>": l.max_grad = r.f32()
                elif tk == "<AffineAccus>":
                    l.accus["w"] = r.mat(); l.accus["b"] = r.mat()
                    break
                else: raise ValueError(f"unsupported token {tk}")
            l.params["w"] = r.mat(); l.params["b"] = r.mat()
        layers.append(l)
    return NetSpec(layers)


# ----------------------------------------------------------------------------- batches
@dataclass
class Batch:
    feats: np.ndarray        # [T*S, I] float32, row = t*S + s, zero padded
    frames: np.ndarray       # [S] int32
    labels: List[np.ndarray]  # S arrays of int32 (1-based class ids)

    @property
    def S(self) -> int:
        return int(self.frames.shape[0])

    @property
    def T(self) -> int:
        return self.feats.shape[0] // self.S

    @property
    def valid_frames(self) -> int:
        return int(self.frames.sum())


def pack_utterances(utts: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """train-ctc-parallel.cc:186-193: row t*S+s <- frame t of utterance s, zero padded to Tmax."""
    S = len(utts)
    T = max(u.shape[0] for u in utts)
    I = utts[0].shape[1]
    out = np.zeros((T * S, I), np.float32)
    for s, u in enumerate(utts):
        out[np.arange(u.shape[0]) * S + s] = u
    return out, np.array([u.shape[0] for u in utts], np.int32)


def write_batch_file(path: str, b: Batch) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", 0x45534E42, b.S, b.T, b.feats.shape[1]))
        f.write(b.frames.astype("<i4").tobytes())
        f.write(np.array([len(l) for l in b.labels], "<i4").tobytes())
        for l in b.labels:
            f.write(np.asarray(l, "<i4").tobytes())
        f.write(np.ascontiguousarray(b.feats, "<f4").tobytes())


def write_feature_ark(path: str, keys: Sequence[str], utts: Sequence[np.ndarray]) -> None:
    with open(path, "wb") as f:
        for k, u in zip(keys, utts):
            f.write(k.encode() + b" \0B")
            _wmat(f, u)


def read_feature_ark(path: str):
    """Binary float-matrix archive -> (keys, matrices); the format BaseFloatMatrixWriter produces."""
    data = open(path, "rb").read()
    r = _Reader(data)
    keys, mats = [], []
    while r.p < len(data):
        keys.append(r.tok())
        if data[r.p:r.p + 2] != b"\0B":
            raise ValueError("only binary archives are supported by this reader")
        r.p += 2
        mats.append(r.mat())
    return keys, mats


def write_label_ark(path: str, keys: Sequence[str], labels: Sequence[Sequence[int]]) -> None:
    with open(path, "w") as f:
        for k, l in zip(keys, labels):
            f.write(k + " " + " ".join(str(int(x)) for x in l) + "\n")
