"""NCCL row-sharded evaluation on >= 2 GPUs of one box: multi-GPU == single-GPU ==
oracle.  Skipped on a single-GPU box (the gloo world_size-2 test in
test_sharded_cpu.py covers the host logic there)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_nccl_sharded_matches_single_gpu(gpu):
    n = gpu.runtime.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    p = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
         f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
         "29533", os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
        capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = "\n".join((p.stdout + p.stderr).splitlines()[-30:])
    assert p.returncode == 0, tail
    assert "MULTI_GPU_OK" in p.stdout, tail
