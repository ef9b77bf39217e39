"""CPU test (gloo, world_size 2) of the data-parallel host logic:

* utterances shard across ranks with NO data-path collective except one all-reduce(SUM) of the raw
  gradient; with an unchanged learning rate that reproduces a single-rank batch of all utterances
  (gradients are sums over rows in the reference, bilstm-parallel-layer.h:505);
* the momentum / clip / update must be applied AFTER the all-reduce, identically on every rank.

The per-rank compute here is the fp64 oracle (no GPU in this container); the collective is real
(torch.distributed, gloo).  The same ordering is what Net::Backpropagate does with NCCL on the B200.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT, case
from eesen_b200 import kaldi_io, synth


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w = synth.WORKLOADS["tiny"]
    net = synth.make_model(w, seed=3)
    full = synth.make_batch(w, seed=5)
    # contiguous blocks of the length-sorted utterance list, re-packed to the shard's own T
    per = full.S // world
    idx = list(range(rank * per, (rank + 1) * per))
    T = full.T
    utts = [full.feats[np.arange(full.frames[s]) * full.S + s] for s in idx]
    feats, frames = kaldi_io.pack_utterances(utts)
    shard = kaldi_io.Batch(feats, frames, [full.labels[s] for s in idx])
    lr, mom, steps = 1e-2, 0.9, 2
    on = oracle.OracleNet(net, np.float64)
    zero_corr = [{n: np.zeros_like(v) for n, v in c.items()} for c in on.corr]
    corr = [{n: np.zeros_like(v) for n, v in c.items()} for c in on.corr]
    for _ in range(steps):
        # raw gradient of this shard: momentum 0 into zeroed accumulators, lr 0 (no update yet)
        on.corr = [{n: np.zeros_like(v) for n, v in c.items()} for c in zero_corr]
        for l in on.spec.layers:
            l_max = l.max_grad
            l.max_grad = 0.0   # clipping acts on the momentum buffer after the all-reduce, not on the shard gradient
        on.train_step(shard, 0.0, 0.0)
        g = torch.from_numpy(on.flat_corr().copy())
        dist.all_reduce(g, op=dist.ReduceOp.SUM)          # the ONE data-path collective of a step
        g = g.numpy()
        # identical update on every rank: corr = g + mom*corr ; clip ; w -= lr*coef*corr
        off = 0
        for li, l in enumerate(on.spec.layers):
            if l.kind != "softmax":
                l.max_grad = w.max_grad
            for n in l.param_names():
                k = on.params[li][n].size
                c = g[off:off + k].reshape(on.params[li][n].shape) + mom * corr[li][n]
                if l.max_grad > 0:
                    c = np.clip(c, -l.max_grad, l.max_grad)
                corr[li][n] = c
                on.params[li][n] -= lr * l.learn_rate_coef * c
                off += k
    q.put((rank, on.flat_params()))
    dist.destroy_process_group()


def test_sharded_allreduce_equals_single_rank_batch():
    from oracle import oracle
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single rank, all utterances in one packed batch, the reference's fused order
    w, net, full = case("tiny")
    on = oracle.OracleNet(net, np.float64)
    for _ in range(2):
        on.train_step(full, 1e-2, 0.9)
    ref = on.flat_params()
    assert np.abs(res[0] - res[1]).max() == 0.0          # replicas stay bit-identical
    assert np.abs(res[0] - ref).max() < 1e-10, np.abs(res[0] - ref).max()
