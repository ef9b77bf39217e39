"""Multi-GPU equivalence check (run under torchrun on an N-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/multi_gpu_equiv.py [workload]

Every rank trains `steps` steps on ITS shard of utterances through the public level-2 API (the NCCL
all-reduce of the raw gradient happens inside Net::Backpropagate); rank 0 then trains a single-GPU
net on the concatenated batch of all N*S utterances and compares the parameters: with gradients
that are sums over rows and an unchanged learning rate the two must agree to fp32 reduction-order
tolerance (per-parameter abs 2e-6 after the update; SURVEY.md asks rel 1e-5 after one step).
Also checks that all replicas stay bit-identical.
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eesen_b200 import binding, kaldi_io, synth  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "small"
    steps = 2
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = synth.WORKLOADS[wl]
    net = synth.make_model(w, seed=3)
    lr, mom = 1e-3, 0.9
    ctx = binding.Context(local)
    obj = [ctx.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    ctx.nccl_init(rank, world, obj[0])
    d = tempfile.mkdtemp()
    path = os.path.join(d, "model")
    kaldi_io.write_model(path, net)
    shards = [synth.make_batch(w, seed=100 + r) for r in range(world)]
    n = binding.Net(ctx, path)
    n.set_train_options(lr, mom)
    for _ in range(steps):
        st = n.train_step(shards[rank].feats, shards[rank].frames, shards[rank].labels, True)
    p = torch.from_numpy(n.params()).cuda()
    gathered = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    ok = True
    if rank == 0:
        for r in range(1, world):
            same = bool(torch.equal(gathered[0], gathered[r]))
            print(f"replica {r} bit-identical to replica 0: {same}")
            ok &= same
        # single-GPU reference: all utterances in one packed batch
        utts, labels = [], []
        for b in shards:
            for s in range(b.S):
                utts.append(b.feats[np.arange(b.frames[s]) * b.S + s])
                labels.append(b.labels[s])
        feats, frames = kaldi_io.pack_utterances(utts)
        ctx1 = binding.Context(local)   # no NCCL: world of one
        n1 = binding.Net(ctx1, path)
        n1.set_train_options(lr, mom)
        for _ in range(steps):
            n1.train_step(feats, frames, labels, True)
        err = np.abs(n1.params() - n.params()).max()
        print(f"{world}-GPU sharded vs 1-GPU batch of {len(utts)} utterances after {steps} steps: max |dparam| = {err:.3e}")
        ok &= err < 2e-6
        print("MULTI_GPU_EQUIV", "PASS" if ok else "FAIL")
    dist.barrier()
    n.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
