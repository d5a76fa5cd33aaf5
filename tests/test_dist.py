"""N>1 path: world_size-2 gloo test on CPU (host logic + DP identity), torchrun NCCL test on >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _torchrun(mode, nproc, port, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "dist_worker.py"), mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))


def test_data_parallel_identity_gloo_world2():
    r = _torchrun("gloo", 2, 29611)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
@pytest.mark.parametrize("exchange,env", [("peer-memory", {}), ("nccl", {"TCR_P2P": "0"})])
def test_gradient_exchange_matches_oracle(exchange, env):
    """Both exchanges of the data-parallel step: the update kernel summing the ranks' gradients from peer memory (default on
    one node) and the ncclAllReduce fallback; 4 steps, averaged gradient vs the oracle's shard-wise mean, identical replicas."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    r = _torchrun("nccl", min(torch.cuda.device_count(), 8), 29612 if exchange == "nccl" else 29613, env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "replicas bit-identical" in r.stdout and f"exchange={exchange}" in r.stdout, r.stdout[-500:]


@pytest.mark.gpu
def test_sync_bn_flag_equals_single_device_global_batch():
    """tcr_comm_set_sync_bn (parity-test flag): with BatchNorm statistics all-reduced over the ranks, every rank's
    step equals the oracle's single-device step on the concatenated global batch."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    r = _torchrun("syncbn", 2, 29614)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "syncbn world=2" in r.stdout, r.stdout[-500:]
