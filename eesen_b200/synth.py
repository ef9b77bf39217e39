"""Seeded synthetic workloads for the BASELINE.json configs.

Features: i.i.d. N(0,1) fp32 of dim 40 (post-CMVN fbank statistics); labels uniform in
[1, K-1] (repeats allowed, exercising the l[j]==l[j-2] branch of the CTC recursion);
utterances sorted by length, longest first (asr_egs/wsj/steps/train_ctc_parallel.sh:84-88).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from .kaldi_io import Batch, NetSpec, make_net, pack_utterances


@dataclass(frozen=True)
class Workload:
    name: str
    in_dim: int
    cells: int
    layers: int
    classes: int
    S: int
    t_lo: int
    t_hi: int
    lab_lo: int
    lab_hi: int
    learn_rate: float = 4e-5
    momentum: float = 0.9
    max_grad: float = 50.0


WORKLOADS = {
    # C1: 2-utterance synthetic fbank, 1x128 BiLSTM, 5-label CTC (K = 6 with blank)
    "c1": Workload("c1_2utt_1x128_k6", 40, 128, 1, 6, 2, 37, 50, 5, 7),
    # C2: 4x320 BiLSTM phone-CTC, 40-dim fbank, 64-utt batch (the metric's config)
    "c2": Workload("c2_64utt_4x320_k46", 40, 320, 4, 46, 64, 400, 600, 30, 60),
    # C4: 5x512 BiLSTM char-CTC, 2000-frame utterances
    "c4": Workload("c4_64utt_5x512_k32_t2000", 40, 512, 5, 32, 64, 2000, 2000, 150, 250),
    # small parity shapes the oracle finishes in seconds
    "tiny": Workload("tiny_4utt_2x16_k5", 8, 16, 2, 5, 4, 6, 12, 2, 4),
    "small": Workload("small_8utt_2x64_k12", 40, 64, 2, 12, 8, 30, 50, 5, 10),
    "mid": Workload("mid_16utt_2x320_k46", 40, 320, 2, 46, 16, 40, 80, 8, 20),
}


def make_batch(w: Workload, seed: int, S: int | None = None) -> Batch:
    rng = np.random.default_rng(seed)
    S = S or w.S
    lens = rng.integers(w.t_lo, w.t_hi + 1, size=S)
    lens = np.sort(lens)[::-1].copy()
    if w.name.startswith("c1"):
        lens = np.array([50, 37][:S])
    utts = [rng.standard_normal((int(t), w.in_dim)).astype(np.float32) for t in lens]
    labels: List[np.ndarray] = []
    for t in lens:
        n = int(rng.integers(w.lab_lo, w.lab_hi + 1))
        n = max(1, min(n, int(t) // 2))  # keep the alignment feasible even with repeats
        labels.append(rng.integers(1, w.classes, size=n).astype(np.int32))
    feats, frames = pack_utterances(utts)
    return Batch(feats, frames, labels)


def make_model(w: Workload, seed: int = 0, bidirectional: bool = True) -> NetSpec:
    return make_net(w.in_dim, w.cells, w.layers, w.classes, seed=seed, param_range=0.1,
                    max_grad=w.max_grad, bidirectional=bidirectional)


def flops_per_frame(w: Workload) -> float:
    """48*C*(I+C) per BiLSTM layer + 12*C*K affine (whole train step)."""
    f = 0.0
    d = w.in_dim
    for _ in range(w.layers):
        f += 48.0 * w.cells * (d + w.cells)
        d = 2 * w.cells
    return f + 12.0 * w.cells * w.classes


def hbm_bytes_per_frame(w: Workload, max_lab: int) -> float:
    """SURVEY.md section 8d minimal-traffic model: per layer 3*I + 68*C floats; output block
    6*K + 4*L' floats (L' = 2*Lmax+1).  fp32 storage."""
    fl = 0.0
    d = w.in_dim
    for _ in range(w.layers):
        fl += 3.0 * d + 68.0 * w.cells
        d = 2 * w.cells
    fl += 6.0 * w.classes + 4.0 * (2 * max_lab + 1)
    return 4.0 * fl
