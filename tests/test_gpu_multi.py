"""Multi-process NCCL path (one process per GPU).  Needs >= 2 GPUs; skipped on a single-GPU box, where the
same slab logic is covered by tests/test_gpu_slabs.py with virtual ranks."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_slabs_match_single_gpu(gpu_lib):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(here, "multi_gpu_worker.py")]
    env = dict(os.environ)
    env["NDSM_SLAB_MIN_PLANES"] = "8"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("MULTI_GPU_OK") == 2
