"""WFST one-best search (BASELINE config 5).

The reference decoder cannot be built here (OpenFst 1.4.1 is un-vendored and absent) and ships no decoder tests:
PARITY IS UNPINNED against the reference binary.  The CPU restatement (oracle/decoder_ref.c, following
src/decoder/lattice-faster-decoder.cc) is pinned on small graphs against an independent exhaustive dynamic
programme over (frame, state); the CUDA search is then held to the restatement: identical word sequences,
costs to fp32 rounding, on random graphs and on a CTC-topology lexicon graph, 64 utterances in one batch."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eesen_b200 import wfst  # noqa: E402
from oracle import oracle  # noqa: E402


def exhaustive_best_path(g, ll, scale):
    """Independent check: exact Viterbi over (frame, state) in float64, epsilon arcs by Bellman-Ford, no pruning."""
    INF = float("inf")
    ns = g.num_states
    cost = [INF] * ns
    words = [None] * ns
    cost[g.start] = 0.0
    words[g.start] = []

    def closure(cost, words):
        changed = True
        while changed:
            changed = False
            for s in range(ns):
                if cost[s] == INF:
                    continue
                for a in range(g.eps[s], g.row[s + 1]):
                    c = cost[s] + float(g.weight[a])
                    d = int(g.nextstate[a])
                    if c < cost[d] - 1e-12:
                        cost[d] = c
                        words[d] = words[s] + ([int(g.olabel[a])] if g.olabel[a] else [])
                        changed = True
    closure(cost, words)
    for t in range(ll.shape[0]):
        nc, nw = [INF] * ns, [None] * ns
        for s in range(ns):
            if cost[s] == INF:
                continue
            for a in range(g.row[s], g.eps[s]):
                c = cost[s] + float(g.weight[a]) - scale * float(ll[t, g.ilabel[a] - 1])
                d = int(g.nextstate[a])
                if c < nc[d]:
                    nc[d] = c
                    nw[d] = words[s] + ([int(g.olabel[a])] if g.olabel[a] else [])
        cost, words = nc, nw
        closure(cost, words)
    best, bw = INF, None
    any_final = any(cost[s] < INF and np.isfinite(g.final[s]) for s in range(ns))
    for s in range(ns):
        if cost[s] == INF:
            continue
        c = cost[s] + (float(g.final[s]) if any_final else 0.0)
        if any_final and not np.isfinite(g.final[s]):
            continue
        if c < best:
            best, bw = c, words[s]
    return bw, best


def test_hand_built_graph_known_answer():
    """Two competing words over a 6-state graph: 'A' = tokens 2 2 3, 'B' = tokens 2 4; blank = token 1."""
    arcs = [(0, 0, 0, 0.5, 1), (0, 0, 0, 0.7, 3),                       # hub -> word entries (LM costs)
            (1, 2, 0, 0.0, 1), (1, 3, 0, 0.0, 2), (2, 3, 0, 0.0, 2), (2, 0, 11, 0.0, 0),    # word A (id 11)
            (3, 2, 0, 0.0, 3), (3, 4, 0, 0.0, 4), (4, 4, 0, 0.0, 4), (4, 0, 22, 0.0, 0),    # word B (id 22)
            (0, 1, 0, 0.0, 0)]                                           # blank self-loop on the hub
    g = wfst.Graph.from_arcs(5, 0, arcs, [(0, 0.0)])
    K = 4
    def frames(seq):
        ll = np.full((len(seq), K), -8.0, np.float32)
        for t, k in enumerate(seq):
            ll[t, k - 1] = -0.1
        return ll
    w, c, nf, _ = oracle.decode_best_path(g, frames([2, 2, 3, 1, 2, 4, 4]), 1.0, 16.0)
    assert w == [11, 22] and nf == 7
    assert abs(c - (0.5 + 0.7 + 7 * 0.1)) < 1e-5
    w, c, _, _ = oracle.decode_best_path(g, frames([1, 2, 4, 1]), 1.0, 16.0)
    assert w == [22] and abs(c - (0.7 + 0.4)) < 1e-5


@pytest.mark.parametrize("seed", range(8))
def test_restatement_vs_exhaustive_dynamic_programme(seed):
    rng = np.random.default_rng(seed)
    g = wfst.random_graph(rng, num_states=int(rng.integers(6, 14)), tokens=4, words=6)
    T = int(rng.integers(3, 9))
    ll = np.log(rng.dirichlet(np.ones(4), size=T)).astype(np.float32)
    w, c, nf, _ = oracle.decode_best_path(g, ll, 0.8, 1e4)      # beam wide open: no search errors possible
    we, ce = exhaustive_best_path(g, ll, 0.8)
    assert nf == T and w == we, (w, we)
    assert abs(c - ce) < 1e-4 * max(1.0, abs(ce))


def test_beam_prunes_and_min_max_active_follow_get_cutoff():
    rng = np.random.default_rng(3)
    g = wfst.random_graph(rng, 200, 8, 30)
    ll = np.log(rng.dirichlet(np.ones(8) * 0.3, size=40)).astype(np.float32)
    w_wide, c_wide, _, a_wide = oracle.decode_best_path(g, ll, 1.0, 1e4)
    w_beam, c_beam, _, a_beam = oracle.decode_best_path(g, ll, 1.0, 6.0)
    assert a_beam < a_wide and c_beam >= c_wide - 1e-3          # pruning expands fewer arcs and can only lose
    w_max, c_max, _, a_max = oracle.decode_best_path(g, ll, 1.0, 1e4, max_active=20)
    assert a_max < a_wide and c_max >= c_wide - 1e-3
    w_min, c_min, _, a_min = oracle.decode_best_path(g, ll, 1.0, 0.5, min_active=50)
    w_nar, c_nar, _, a_nar = oracle.decode_best_path(g, ll, 1.0, 0.5)
    assert a_min > a_nar and c_min <= c_nar + 1e-3              # min_active keeps the search alive under a tiny beam


def test_synthetic_tlg_structure():
    g = wfst.synthetic_tlg(0, 500, 46)
    assert g.row[-1] == g.num_arcs and np.all(g.eps >= g.row[:-1]) and np.all(g.eps <= g.row[1:])
    il = g.ilabel
    for s in range(g.num_states):
        k = g.eps[s] - g.row[s]
        seg = il[g.row[s]:g.row[s + 1]]
        assert np.all(seg[:k] > 0) and np.all(seg[k:] == 0)
    assert (g.olabel > 0).sum() == 500                          # one word-emitting arc per word
    # a clean realisation of word 7 decodes to word 7
    rng = np.random.default_rng(1)
    w = 6
    # walk the word's chain: entry blank state -> token states
    s = int(g.nextstate[g.row[0] + w])
    seq = []
    while True:
        arcs = range(g.row[s], g.eps[s])
        nxt = [a for a in arcs if g.nextstate[a] != s]
        if not nxt:
            break
        a = max(nxt, key=lambda a: g.nextstate[a])
        seq.append(int(g.ilabel[a])); seq.append(int(g.ilabel[a]))
        s = int(g.nextstate[a])
    ll = np.full((len(seq), 46), -9.0, np.float32)
    for t, k in enumerate(seq):
        ll[t, k - 1] = -0.05
    words, c, _, _ = oracle.decode_best_path(g, ll, 1.0, 12.0)
    assert words == [w + 1]


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def ctx():
    from eesen_b200 import binding
    c = binding.Context(0)
    yield c
    c.close()


def _batch_loglikes(rng, S, T, K, frames, peaky=0.3):
    import torch
    ll = np.log(rng.dirichlet(np.ones(K) * peaky, size=T * S)).astype(np.float32)   # packed rows t*S + s
    ld = (K + 3) // 4 * 4
    pad = np.zeros((T * S, ld), np.float32); pad[:, :K] = ll; pad[:, K:] = np.nan
    return ll, torch.from_numpy(pad).cuda(), ld


@pytest.mark.gpu
@pytest.mark.parametrize("seed,states,beam", [(0, 40, 1e4), (1, 300, 8.0), (2, 2000, 6.0), (3, 2000, 12.0)])
def test_gpu_search_vs_restatement_random_graphs(ctx, seed, states, beam):
    rng = np.random.default_rng(seed)
    K, S, T = 9, 16, 30
    g = wfst.random_graph(rng, states, K, 50)
    frames = np.sort(rng.integers(T // 2, T + 1, size=S))[::-1].copy(); frames[0] = T
    ll, d_ll, ld = _batch_loglikes(rng, S, T, K, frames)
    dg = wfst.DeviceGraph(ctx, g)
    words, cost, st = dg.decode(d_ll, ld, K, frames, T, 0.9, beam, frame_cap=1 << 13)
    for s in range(S):
        rows = np.arange(frames[s]) * S + s
        w, c, nf, _ = oracle.decode_best_path(g, ll[rows], 0.9, beam)
        assert nf == frames[s]
        assert words[s] == w, (s, words[s], w)                  # bit-exact word sequence
        assert abs(cost[s] - c) <= 2e-5 * max(1.0, abs(c)), (s, cost[s], c)
    dg.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,states,beam,max_active,min_active", [(4, 2000, 1e4, 25, 0), (5, 2000, 0.5, 2147483647, 40),
                                                                     (6, 3000, 9.0, 60, 20), (7, 300, 6.0, 7, 3)])
def test_gpu_search_max_and_min_active_follow_get_cutoff(ctx, seed, states, beam, max_active, min_active):
    """GetCutoff with --max-active / --min-active (lattice-faster-decoder.cc:594-658): the CUDA search selects the k-th best
    cost exactly (radix select), so words agree bit for bit with the restatement and costs to rounding."""
    rng = np.random.default_rng(seed)
    K, S, T = 9, 12, 24
    g = wfst.random_graph(rng, states, K, 50)
    frames = np.sort(rng.integers(T // 2, T + 1, size=S))[::-1].copy(); frames[0] = T
    ll, d_ll, ld = _batch_loglikes(rng, S, T, K, frames)
    dg = wfst.DeviceGraph(ctx, g)
    words, cost, st = dg.decode(d_ll, ld, K, frames, T, 0.9, beam, frame_cap=1 << 13, max_active=max_active, min_active=min_active)
    limited = 0
    for s in range(S):
        rows = np.arange(frames[s]) * S + s
        w, c, nf, a_lim = oracle.decode_best_path(g, ll[rows], 0.9, beam, max_active=max_active, min_active=min_active)
        _, _, _, a_free = oracle.decode_best_path(g, ll[rows], 0.9, beam)
        limited += a_lim != a_free
        assert words[s] == w, (s, words[s], w)
        if w is not None:
            assert abs(cost[s] - c) <= 2e-5 * max(1.0, abs(c)), (s, cost[s], c)
    assert limited > 0          # the limits actually changed the search
    dg.close()


@pytest.mark.gpu
def test_gpu_search_synthetic_tlg_64_utterances(ctx):
    """CTC-topology lexicon graph (20 k words, ~0.5 M arcs), 64 utterances of noisy realisations of word sequences."""
    rng = np.random.default_rng(5)
    K, S = 46, 64
    g = wfst.synthetic_tlg(2, 20000, K)
    T = 120
    frames = np.sort(rng.integers(80, T + 1, size=S))[::-1].copy(); frames[0] = T
    ll, d_ll, ld = _batch_loglikes(rng, S, T, K, frames, peaky=0.15)
    dg = wfst.DeviceGraph(ctx, g)
    words, cost, st = dg.decode(d_ll, ld, K, frames, T, 1.0, 7.0, frame_cap=1 << 16, tok_cap=1 << 22)
    nonempty = 0
    for s in range(0, S, 7):                                      # the restatement is serial: check every 7th utterance
        rows = np.arange(frames[s]) * S + s
        w, c, nf, _ = oracle.decode_best_path(g, ll[rows], 1.0, 7.0)
        assert words[s] == w, (s, words[s][:10], (w or [])[:10])
        assert abs(cost[s] - c) <= 2e-5 * max(1.0, abs(c))
        nonempty += bool(w)
    assert nonempty > 0 and st["closure_rounds"] >= T
    dg.close()


@pytest.mark.gpu
def test_gpu_search_rejects_unsupported_pruning_and_reports_overflow(ctx):
    from eesen_b200 import binding
    rng = np.random.default_rng(0)
    g = wfst.random_graph(rng, 500, 5, 10)
    ll, d_ll, ld = _batch_loglikes(rng, 2, 10, 5, [10, 10])
    dg = wfst.DeviceGraph(ctx, g)
    with pytest.raises(binding.EesenB200Error):                   # far too few token slots for a wide-open beam
        dg.decode(d_ll, ld, 5, [10, 10], 10, 1.0, 1e4, frame_cap=8, tok_cap=64)
    dg.close()
