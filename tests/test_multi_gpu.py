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


@pytest.mark.gpu
@pytest.mark.parametrize("N,K,C", [(20011, 96, 32), (3001, 37, 5), (1, 3, 2)])
def test_device_resident_categorical_matches_sync_call(gpu, N, K, C):
    """One GPU: smc_categorical_logit_glm_device (device parameters in, packed device
    result out, no host synchronisation -- what every rank of the sharded driver runs)
    gives bit for bit what the synchronous call gives, and the oracle's values."""
    import numpy as np
    import torch

    from math_b200 import _lib
    from math_b200.sharded import ShardedCategoricalGlm
    from oracle import pyoracle as po
    from tests.multi_gpu_worker import _NoDist
    from tests.util import assert_grad, assert_logp, make_inputs

    mb = gpu
    d = make_inputs("categorical", N, K, seed=21, C=C)
    x = mb.to_matrix_cuda(np.asfortranarray(d["x"]))
    y = mb.to_matrix_cuda(np.ascontiguousarray(d["y"]))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        mb.runtime.set_stream(stream.cuda_stream)
        try:
            glm = ShardedCategoricalGlm(y, x, K, C, flags=_lib.VAR_ALPHA | _lib.VAR_BETA,
                                        device="cuda:0", dist=_NoDist())
            out = glm.evaluate(ShardedCategoricalGlm.pack_params(d["alpha"], d["beta"]))
            stream.synchronize()
            res = glm.unpack(out.cpu().numpy())
        finally:
            mb.runtime.set_stream(None)
    s = mb.categorical_logit_glm_lpmf(y, x, d["alpha"], d["beta"])
    assert res["logp"] == s.logp and res["nonfinite"] == 0.0
    assert np.array_equal(res["d_alpha"], s.d_alpha)
    assert np.array_equal(res["d_beta"], s.d_beta)
    o = po.categorical_logit_glm(d["y"], d["x"], d["alpha"], d["beta"])
    assert_logp(res["logp"], o["logp"])
    assert_grad(res["d_alpha"], o["d_alpha"], "d_alpha")
    assert_grad(res["d_beta"].ravel(order="F"), o["d_beta"].ravel(order="F"), "d_beta")
