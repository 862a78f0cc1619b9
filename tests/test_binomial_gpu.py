"""GPU parity of binomial_logit_glm_lpmf (the seventh GLM) against
the CPU oracle through the C ABI -- the cases of the reference's device test
(test/unit/math/opencl/rev/binomial_logit_glm_lpmf_test.cpp: error_checking,
small_simple, broadcast_n / broadcast_N, zero_instances / zero_attributes,
vector alpha, big) plus the fused-kernel tile shapes."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import assert_grad, assert_logp, make_inputs

pytestmark = pytest.mark.gpu

SHAPES = [(3, 2), (153, 71), (1000, 1), (4097, 32), (777, 33), (5000, 128),
          (3001, 200), (20000, 256), (2000, 300)]
ALL = po.VAR_X | po.VAR_ALPHA | po.VAR_BETA


@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("propto", [False, True])
def test_binomial(gpu, N, K, propto):
    d = make_inputs("binomial", N, K, seed=N * 13 + K)
    x = gpu.to_matrix_cuda(d["x"])
    n = gpu.to_matrix_cuda(d["y"])
    t = gpu.to_matrix_cuda(d["trials"])
    r = gpu.binomial_logit_glm_lpmf(n, t, x, d["alpha"], d["beta"], propto=propto,
                                    var=("x", "alpha", "beta"))
    o = po.binomial_logit_glm(d["y"], d["trials"], d["x"], d["alpha"], d["beta"],
                              flags=ALL | (po.PROPTO if propto else 0))
    assert o["rc"] == 0
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_alpha, o["d_alpha"][0], "d_alpha", scale=np.abs(o["d_beta"]).max())
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_x.to_host(), o["d_x"], "d_x")


def test_binomial_vector_alpha_and_broadcasts(gpu):
    N, K = 1531, 71
    d = make_inputs("binomial", N, K, seed=3, vec_alpha=True)
    x = gpu.to_matrix_cuda(d["x"])
    a = gpu.to_matrix_cuda(d["alpha"])
    n, t = gpu.to_matrix_cuda(d["y"]), gpu.to_matrix_cuda(d["trials"])
    r = gpu.binomial_logit_glm_lpmf(n, t, x, a, d["beta"])
    o = po.binomial_logit_glm(d["y"], d["trials"], d["x"], d["alpha"], d["beta"])
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_alpha.to_host().ravel(), o["d_alpha"], "d_alpha_vec")
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    # broadcast_N: one population size for every instance; then the same call
    # again (cached pair statistics) and after re-uploading n (cache invalidated)
    nn = np.minimum(d["y"], 25).astype(np.int32)
    n2 = gpu.to_matrix_cuda(nn)
    for _ in range(2):
        r = gpu.binomial_logit_glm_lpmf(n2, 25, x, 0.2, d["beta"])
        o = po.binomial_logit_glm(nn, [25], d["x"], 0.2, d["beta"])
        assert_logp(r.logp, o["logp"])
        assert_grad(r.d_beta, o["d_beta"], "d_beta")
    nn3 = np.minimum(d["y"], 7).astype(np.int32)
    n2.upload_rows(0, nn3)
    r = gpu.binomial_logit_glm_lpmf(n2, 25, x, 0.2, d["beta"])
    o = po.binomial_logit_glm(nn3, [25], d["x"], 0.2, d["beta"])
    assert_logp(r.logp, o["logp"])
    # broadcast_n: one success count; both broadcast (coefficient term counted N times)
    r = gpu.binomial_logit_glm_lpmf(0, t, x, 0.2, d["beta"])
    o = po.binomial_logit_glm([0], d["trials"], d["x"], 0.2, d["beta"])
    assert_logp(r.logp, o["logp"])
    r = gpu.binomial_logit_glm_lpmf(7, 19, x, 0.2, d["beta"])
    o = po.binomial_logit_glm([7], [19], d["x"], 0.2, d["beta"])
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_alpha, o["d_alpha"][0], "d_alpha", scale=np.abs(o["d_beta"]).max())


def test_binomial_saturated_tails(gpu):
    """|theta| up to ~1e2: log_inv_logit / log1m_inv_logit on their stable
    branches; n = 0 and n = N rows."""
    x = np.array([[-12, 46], [-42, 24], [25, 27], [0.0, 0.0]], float)
    n = np.array([3, 0, 500, 2], np.int32)
    t = np.array([3, 8, 1000, 4], np.int32)
    r = gpu.binomial_logit_glm_lpmf(gpu.to_matrix_cuda(n), gpu.to_matrix_cuda(t),
                                    gpu.to_matrix_cuda(x), 0.0, [0.3, 2.0],
                                    var=("x", "alpha", "beta"))
    o = po.binomial_logit_glm(n, t, x, 0.0, [0.3, 2.0], flags=ALL)
    assert o["rc"] == 0
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_x.to_host(), o["d_x"], "d_x")


def test_binomial_zero_sizes_and_errors(gpu):
    x = gpu.to_matrix_cuda(np.array([[-12, 46], [-42, 24], [25, 27]], float) / 100)
    n = gpu.to_matrix_cuda(np.array([0, 1, 0], np.int32))
    t = gpu.to_matrix_cuda(np.array([1, 2, 3], np.int32))
    beta = [0.3, 2.0]
    # zero_instances / zero_attributes: size_zero() wins over every other check
    x0 = gpu.MatrixCuda(0, 2)
    e0 = gpu.MatrixCuda(0, 1, np.int32)
    assert gpu.binomial_logit_glm_lpmf(e0, e0, x0, 0.3, beta).logp == 0.0
    assert gpu.binomial_logit_glm_lpmf(n, t, gpu.MatrixCuda(3, 0), 0.3, []).logp == 0.0
    # all-data + propto: nothing to compute
    assert gpu.binomial_logit_glm_lpmf(n, t, x, 0.3, beta, propto=True, var=()).logp == 0.0
    with pytest.raises(ValueError):  # sizes -> std::invalid_argument
        gpu.binomial_logit_glm_lpmf(gpu.to_matrix_cuda(np.array([0, 1], np.int32)), t, x,
                                    0.3, beta)
    with pytest.raises(ValueError):
        gpu.binomial_logit_glm_lpmf(n, gpu.to_matrix_cuda(np.array([1, 2], np.int32)), x,
                                    0.3, beta)
    with pytest.raises(ValueError):
        gpu.binomial_logit_glm_lpmf(n, t, x, 0.3, [0.3, 2.0, 1.0])
    with pytest.raises(gpu.DomainError):  # n > N
        gpu.binomial_logit_glm_lpmf(gpu.to_matrix_cuda(np.array([0, 3, 0], np.int32)), t, x,
                                    0.3, beta)
    with pytest.raises(gpu.DomainError):  # n < 0
        gpu.binomial_logit_glm_lpmf(gpu.to_matrix_cuda(np.array([0, -1, 0], np.int32)), t,
                                    x, 0.3, beta)
    with pytest.raises(gpu.DomainError):  # N < 0
        gpu.binomial_logit_glm_lpmf(n, -2, x, 0.3, beta)
    with pytest.raises(gpu.DomainError):
        gpu.binomial_logit_glm_lpmf(5, 4, x, 0.3, beta)
    for bad in (np.inf, -np.inf, np.nan):  # lazy finiteness checks, L118-122
        with pytest.raises(gpu.DomainError):
            gpu.binomial_logit_glm_lpmf(n, t, x, 0.3, [0.3, bad])
        with pytest.raises(gpu.DomainError):
            gpu.binomial_logit_glm_lpmf(n, t, x, bad, beta)
        xb = np.array([[-12, 46], [-42, bad], [25, 27]], float)
        with pytest.raises(gpu.DomainError):
            gpu.binomial_logit_glm_lpmf(n, t, gpu.to_matrix_cuda(xb), 0.3, beta)


def test_binomial_device_entry_matches_sync(gpu):
    """smc_glm_eval_device(family 5): the asynchronous, device-parameter entry the
    row-sharded driver uses returns what the synchronous call returns."""
    import torch
    from math_b200 import _lib
    from math_b200.sharded import ShardedGlm
    N, K = 4099, 64
    d = make_inputs("binomial", N, K, seed=21)
    x, n, t = (gpu.to_matrix_cuda(d[k]) for k in ("x", "y", "trials"))
    glm = ShardedGlm("binomial_logit", n, x, K, alpha=d["alpha"], aux=t,
                     flags=_lib.VAR_ALPHA | _lib.VAR_BETA, device="cuda:0")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        gpu.runtime.set_stream(stream.cuda_stream)
        out = glm.evaluate(d["beta"])
        stream.synchronize()
    gpu.runtime.set_stream(None)
    res = glm.unpack(out.cpu().numpy())
    r = gpu.binomial_logit_glm_lpmf(n, t, x, d["alpha"], d["beta"])
    assert_logp(res["logp"], r.logp)
    assert_grad(res["d_beta"], r.d_beta, "d_beta")
