"""WFST search graphs for the decoder slice: a plain CSR container (no OpenFst), seeded
synthetic graphs, and the ctypes front-end of eesen_b200_graph_create / eesen_b200_decode_best_path.

Graph convention (what `latgen-faster` reads as TLG.fst, reference src/decoderbin/latgen-faster.cc:62-70): input
labels are 1-based CTC token ids (0 = epsilon; the decodable shifts by one, src/decoder/decodable-matrix.h:54-56),
output labels are word ids (0 = none), weights are costs (-log).  The arcs of a state are stored emitting arcs
first, then epsilon-input arcs."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np


@dataclass
class Graph:
    num_states: int
    start: int
    row: np.ndarray        # int32 [num_states + 1]
    eps: np.ndarray        # int32 [num_states]: first epsilon-input arc of each state
    ilabel: np.ndarray     # int32 [num_arcs]
    olabel: np.ndarray     # int32 [num_arcs]
    weight: np.ndarray     # float32 [num_arcs]
    nextstate: np.ndarray  # int32 [num_arcs]
    final: np.ndarray      # float32 [num_states], +inf = not final

    @property
    def num_arcs(self) -> int:
        return int(self.ilabel.shape[0])

    @staticmethod
    def from_arcs(num_states: int, start: int, arcs: Sequence[Tuple[int, int, int, float, int]],
                  finals: Sequence[Tuple[int, float]]) -> "Graph":
        """arcs: (from, ilabel, olabel, weight, to) in any order."""
        a = sorted(arcs, key=lambda x: (x[0], x[1] == 0))       # stable: per state, emitting arcs first
        frm = np.array([x[0] for x in a], np.int64)
        il = np.array([x[1] for x in a], np.int32)
        row = np.zeros(num_states + 1, np.int32)
        np.add.at(row, frm + 1, 1)
        row = np.cumsum(row).astype(np.int32)
        n_emit = np.zeros(num_states, np.int32)
        np.add.at(n_emit, frm[il != 0], 1)
        eps = (row[:-1] + n_emit).astype(np.int32)
        fin = np.full(num_states, np.inf, np.float32)
        for s, w in finals:
            fin[s] = w
        return Graph(num_states, start, row, eps, il, np.array([x[2] for x in a], np.int32),
                     np.array([x[3] for x in a], np.float32), np.array([x[4] for x in a], np.int32), fin)


def random_graph(rng: np.random.Generator, num_states: int, tokens: int, words: int, emit_per_state: int = 3,
                 eps_per_state: float = 0.4) -> Graph:
    """Small random test graph: every state has a self-loop and a few emitting arcs, some states have epsilon-input
    arcs (weights > 0, so there is no zero-cost epsilon cycle), some arcs carry words; random real weights make ties
    between competing paths a measure-zero event."""
    arcs = []
    for s in range(num_states):
        arcs.append((s, int(rng.integers(1, tokens + 1)), 0, float(rng.uniform(0.05, 1.5)), s))
        for _ in range(emit_per_state):
            arcs.append((s, int(rng.integers(1, tokens + 1)), int(rng.integers(1, words + 1)) if rng.random() < 0.3 else 0,
                         float(rng.uniform(0.05, 3.0)), int(rng.integers(0, num_states))))
        if rng.random() < eps_per_state:
            for _ in range(int(rng.integers(1, 3))):
                arcs.append((s, 0, int(rng.integers(1, words + 1)) if rng.random() < 0.5 else 0,
                             float(rng.uniform(0.2, 2.0)), int(rng.integers(0, num_states))))
    finals = [(int(s), float(rng.uniform(0.0, 2.0))) for s in rng.choice(num_states, size=max(1, num_states // 4), replace=False)]
    return Graph.from_arcs(num_states, 0, arcs, finals)


def synthetic_tlg(seed: int, words: int, tokens: int, min_len: int = 3, max_len: int = 8) -> Graph:
    """CTC-topology lexicon graph with a unigram "LM" hub, the shape of the T o L o G graphs of the Eesen recipes
    (asr_egs/wsj/utils/ctc_compile_dict_token.sh: token 1 = <blk>): from the hub an epsilon arc with the word's
    unigram cost enters each word; a word is a chain of its tokens, each token state with a self-loop (repeats) and an
    optional blank state with its own self-loop between tokens; the last token state leaves through an epsilon arc
    that emits the word id and returns to the hub.  ~ (4*len + 2) arcs per word: words = 400 000 gives ~10 M arcs."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_len, max_len + 1, size=words)
    logp = rng.gumbel(size=words)
    cost = (np.log(np.exp(logp - logp.max()).sum()) - (logp - logp.max())).astype(np.float32)   # -log softmax
    # states: 0 = hub (final); per word and token position: token state, blank state
    tok_state_off = 1 + 2 * np.concatenate([[0], np.cumsum(lens)[:-1]])
    num_states = int(1 + 2 * lens.sum())
    frm, il, ol, wt, to = [], [], [], [], []
    toks = rng.integers(2, tokens + 1, size=int(lens.sum())).astype(np.int32)   # token 1 is blank
    pos0 = np.concatenate([[0], np.cumsum(lens)[:-1]])
    # hub -> first blank-or-token of each word: epsilon arc into an entry that is the word's leading blank state
    first_tok_state = tok_state_off
    for arr, v in ((frm, np.zeros(words, np.int64)), (il, np.zeros(words, np.int32)), (ol, np.zeros(words, np.int32)),
                   (wt, cost), (to, first_tok_state - 0)):
        arr.append(np.asarray(v))
    # entry state of word w is its first "blank" state b_0 (index tok_state_off[w]); token state t_i = off + 2*i + 1
    idx = np.arange(int(lens.sum()))
    w_of = np.repeat(np.arange(words), lens)
    i_in = idx - pos0[w_of]
    b = tok_state_off[w_of] + 2 * i_in          # blank state in front of token i
    t = b + 1                                   # token state i
    small = rng.uniform(0.01, 0.2, size=(4, idx.size)).astype(np.float32)
    # blank self-loop; blank -> token; token self-loop
    for f_, i_, t_, w_ in ((b, np.ones_like(toks), b, small[0]), (b, toks, t, small[1]), (t, toks, t, small[2])):
        frm.append(f_); il.append(i_.astype(np.int32)); ol.append(np.zeros_like(toks)); wt.append(w_); to.append(t_)
    # token i -> blank state in front of token i+1 (emitting blank) and directly -> token i+1 when the tokens differ
    nxt = i_in + 1 < lens[w_of]
    f2 = t[nxt]; nb = b[nxt] + 2; nt = nb + 1; ntok = toks[np.flatnonzero(nxt) + 1]
    frm.append(f2); il.append(np.ones_like(ntok)); ol.append(np.zeros_like(ntok)); wt.append(small[3][nxt]); to.append(nb)
    diff = ntok != toks[nxt]
    frm.append(f2[diff]); il.append(ntok[diff]); ol.append(np.zeros_like(ntok[diff])); wt.append(small[1][nxt][diff]); to.append(nt[diff])
    # last token -> hub: epsilon arc that emits the word
    last = ~nxt
    frm.append(t[last]); il.append(np.zeros(words, np.int32)); ol.append((w_of[last] + 1).astype(np.int32))
    wt.append(rng.uniform(0.01, 0.1, size=words).astype(np.float32)); to.append(np.zeros(words, np.int64))
    frm = np.concatenate(frm).astype(np.int64); il = np.concatenate(il).astype(np.int32)
    ol = np.concatenate(ol).astype(np.int32); wt = np.concatenate(wt).astype(np.float32); to = np.concatenate(to).astype(np.int32)
    order = np.lexsort((il == 0, frm))          # by state, emitting arcs first
    frm, il, ol, wt, to = frm[order], il[order], ol[order], wt[order], to[order]
    row = np.zeros(num_states + 1, np.int64)
    np.add.at(row, frm + 1, 1)
    row = np.cumsum(row).astype(np.int32)
    n_emit = np.zeros(num_states, np.int64)
    np.add.at(n_emit, frm[il != 0], 1)
    fin = np.full(num_states, np.inf, np.float32)
    fin[0] = 0.0
    return Graph(num_states, 0, row, (row[:-1] + n_emit).astype(np.int32), il, ol, wt, to, fin)


# ------------------------------------------------------------------------------------------------ device front-end
class DeviceGraph:
    """The graph in HBM (eesen_b200_graph_create); decode() runs the batched one-best search."""

    def __init__(self, ctx, g: Graph):
        self.ctx, self.g = ctx, g
        lib = ctx.lib
        lib.eesen_b200_graph_free.argtypes = [C.c_void_p]
        lib.eesen_b200_graph_free.restype = None
        self.h = C.c_void_p()
        arr = lambda a, t: np.ascontiguousarray(a, t)
        self._keep = [arr(g.row, np.int32), arr(g.eps, np.int32), arr(g.ilabel, np.int32), arr(g.olabel, np.int32),
                      arr(g.weight, np.float32), arr(g.nextstate, np.int32), arr(g.final, np.float32)]
        p = [a.ctypes.data_as(C.c_void_p) for a in self._keep]
        ctx.check(lib.eesen_b200_graph_create(ctx.h, g.num_states, g.num_arcs, g.start, p[0], p[1], p[2], p[3], p[4], p[5],
                                              p[6], C.byref(self.h)), "graph_create")

    def decode(self, d_loglikes, ld: int, K: int, frames: Sequence[int], T: int, acoustic_scale: float, beam: float,
               frame_cap: int = 1 << 15, tok_cap: int | None = None, max_out: int = 512,
               max_active: int = 2147483647, min_active: int = 0):
        S = len(frames)
        tok_cap = tok_cap or min(frame_cap * (T + 1), 1 << 24)
        fr = np.ascontiguousarray(frames, np.int32)
        labels = np.zeros((S, max_out), np.int32); n = np.zeros(S, np.int32); cost = np.zeros(S, np.float32)
        stats = (C.c_double * 2)()
        self.ctx.check(self.ctx.lib.eesen_b200_decode_best_path(
            self.ctx.h, self.h, S, T, fr.ctypes.data_as(C.c_void_p), C.c_void_p(d_loglikes.data_ptr()), ld, K,
            C.c_float(acoustic_scale), C.c_float(beam), int(max_active), int(min_active), frame_cap, tok_cap,
            labels.ctypes.data_as(C.c_void_p), max_out, n.ctypes.data_as(C.c_void_p), cost.ctypes.data_as(C.c_void_p), stats),
            "decode_best_path")
        return [labels[s, :max(0, n[s])].tolist() if n[s] >= 0 else None for s in range(S)], cost, \
            {"closure_rounds": stats[0], "device_ms": stats[1]}

    def close(self):
        if self.h:
            self.ctx.lib.eesen_b200_graph_free(self.h)
            self.h = C.c_void_p()
