"""Multi-GPU test (needs >= 2 GPUs on the box, e.g. `gpurun --gpus 2`): one process per GPU under
torchrun, NCCL all-reduce of the cost sums inside libkfbo.so; sharded == unsharded."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_ranks_nccl_allreduce_matches_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29731",
                          os.path.join(ROOT, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi_gpu_check ok" in out.stdout
