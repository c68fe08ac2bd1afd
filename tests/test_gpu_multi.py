"""-m gpu tests that need TWO GPUs on the box (skipped on a one-GPU box; run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_allreduce_stats_over_nccl_inside_the_library():
    """ptg_allreduce_stats (ncclAllGather of the 64-byte record + fixed-order combine kernel, no host sync) on 2 ranks:
    identical on every rank and equal to the host combine of the per-rank records."""
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tools", "check_allreduce_stats.py")], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert "ptg_allreduce_stats on 2 GPUs: OK" in out.stdout
