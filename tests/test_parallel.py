"""Data-parallel path: world_size-2 gloo run on CPU (host-side sharding logic, driven through the oracle)
and, on a box with >= 2 GPUs, the CUDA update sharded over NCCL vs the full-batch oracle."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dp_worker.py")


def _torchrun(mode, nproc, port):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)


def test_shard_helpers():
    from d3rlpy_b200 import parallel

    assert parallel.shard_rows(8, 2, 1) == (4, 8)
    with pytest.raises(ValueError):
        parallel.shard_rows(9, 2, 0)
    t = torch.arange(2 * 4 * 3).view(2, 4, 3)
    assert torch.equal(parallel.shard_noise(t, "NB*", 4, 2, 2, 1), t[:, 2:4])
    u = torch.arange(8 * 3).view(8, 3)  # B=4, N=2 rows b*N+k
    assert torch.equal(parallel.shard_noise(u, "BN*", 4, 2, 2, 0), u[:4])


def test_dp_sharded_oracle_matches_full_batch_gloo_world2():
    r = _torchrun("cpu", 2, 29613)
    assert r.returncode == 0 and "DP_EQUIVALENCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("world,port", [(4, 29623), (8, 29633)])
def test_dp_sharded_oracle_matches_full_batch_gloo_wider_worlds(world, port):
    """The same sharding (row ranges, per-kind noise slices, 1/W loss scaling) at the world sizes the scaling bench
    runs: 4 and 8 ranks over gloo."""
    r = _torchrun("cpu", world, port)
    assert r.returncode == 0 and "DP_EQUIVALENCE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_dp_cuda_update_matches_full_batch_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    r = _torchrun("gpu", 2, 29614)
    assert r.returncode == 0 and "DP_EQUIVALENCE_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
