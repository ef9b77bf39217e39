"""N-GPU == 1-GPU equivalence under the driver's `pytest -m gpu`.

On a box with >= 2 visible GPUs this spawns tests/multi_gpu_equiv.py under torch.distributed.run for
N = 2 and N = all visible GPUs: every rank trains on its own utterance shard through Net::Backpropagate (one
NCCL all-reduce of the raw gradient per step, reference semantics bilstm-parallel-layer.h:505: gradients are
SUMS over rows), then rank 0 checks (a) all replicas hold bit-identical parameters, (b) the N-shard run equals
one GPU training on the union minibatch to fp32 reduction-order tolerance (2e-6 per parameter).
On a 1-GPU box the test is skipped (the gloo world_size-2 CPU test covers the host logic)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _spawn(n, workload, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_equiv.py"), workload]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "MULTI_GPU_EQUIV PASS" in out, out[-4000:]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("workload", ["small", "mid"])
def test_two_gpus_equal_one_gpu_union_batch(workload):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 visible GPUs")
    _spawn(2, workload, 29541)


@pytest.mark.gpu
def test_all_gpus_equal_one_gpu_union_batch():
    n = _ngpu()
    if n < 3:
        pytest.skip("needs > 2 visible GPUs (the 2-GPU case is covered above)")
    _spawn(n, "small", 29543)
