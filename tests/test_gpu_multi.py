"""Multi-GPU path (photon-id partition, frames combined on rank 0 by the
peer-reading gather kernel or by an NCCL reduce) -- needs >= 2 GPUs, skipped
on single-GPU boxes.  The rank-level logic is covered on CPU with gloo in
tests/test_multi_rank_cpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multi_gpu_matches_single_gpu(gpu):
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTI_GPU_CHECK_OK" in res.stdout
