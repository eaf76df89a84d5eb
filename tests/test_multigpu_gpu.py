"""Needs >= 2 GPUs: runs tests/multigpu_check.py under torchrun (NCCL).  Skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_rank_sharded_update_matches_oracle(cuda, exchange):
    """exchange="p2p": the fused peer-memory reduce-scatter + sharded Adam + all-gather kernel (csrc/comm.cuh);
    "nccl": all-reduce between the phases.  Both must reproduce the unsharded fp64 oracle."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multigpu_check.py")]
    env = dict(os.environ, MG_EXCHANGE=exchange)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0 and "multigpu_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
