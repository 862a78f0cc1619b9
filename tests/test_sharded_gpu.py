"""Row-sharded matrices behind the plain entry points (include/stanmath_cuda.h,
"row-sharded matrices"): every GLM family on a sharded x against the same call on a
plain x and against the oracle.  On a box with one GPU the shard set puts several
shards on that GPU (results reduced on the host); with two or more GPUs
test_real_gpus_nccl also runs one shard per GPU with the NCCL all-reduce."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import assert_grad, assert_logp, make_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[([0, 0, 0], None), (None, None), ([0, 0, 0], "host"), (None, "nccl")],
                ids=["3-shards-on-gpu0", "one-shard-per-gpu", "3-shards-on-gpu0-host-copies",
                     "one-shard-per-gpu-nccl-only"])
def shard_set(gpu, request, monkeypatch):
    """Default: small results through the direct slots (every shard's kernel stores its packed
    result and a completion flag straight into pinned host memory, the host adds the slots in
    shard order), large ones through NCCL / the host; SMC_SHARD_REDUCE=host / nccl force the
    read-back copies / the NCCL all-reduce for everything."""
    devices, mode = request.param
    if devices is None and gpu.runtime.device_count() < 2:
        pytest.skip("one shard per GPU needs at least two GPUs")
    if mode:
        monkeypatch.setenv("SMC_SHARD_REDUCE", mode)
    else:
        monkeypatch.delenv("SMC_SHARD_REDUCE", raising=False)
    n = gpu.runtime.shard_init(devices=devices) if devices else gpu.runtime.shard_init(0)
    got = gpu.runtime.shard_reduce_mode()
    assert got == mode if mode else got.startswith("direct+")
    yield n
    gpu.runtime.shard_shutdown()


def _close(a, b, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    np.testing.assert_allclose(a, b, rtol=2e-13, atol=2e-13 * max(1.0, float(np.abs(b).max())),
                               err_msg=what)


@pytest.mark.parametrize("N,K", [(2, 3), (1000, 7), (30011, 100), (8191, 300)])
def test_bernoulli_sharded(gpu, shard_set, N, K):
    d = make_inputs("bernoulli", N, K, seed=N + K)
    xs, ys = gpu.MatrixCuda.from_host_sharded(d["x"]), gpu.MatrixCuda.from_host_sharded(d["y"])
    assert xs.shard_count == shard_set
    np.testing.assert_array_equal(xs.to_host(), d["x"])
    x, y = gpu.to_matrix_cuda(d["x"]), gpu.to_matrix_cuda(d["y"])
    var = ("x", "alpha", "beta")
    rs = gpu.bernoulli_logit_glm_lpmf(ys, xs, d["alpha"], d["beta"], var=var)
    r1 = gpu.bernoulli_logit_glm_lpmf(y, x, d["alpha"], d["beta"], var=var)
    o = po.bernoulli_logit_glm(d["y"], d["x"], d["alpha"], d["beta"],
                               po.VAR_X | po.VAR_ALPHA | po.VAR_BETA)
    _close(rs.logp, r1.logp, "logp vs one GPU")
    _close(rs.d_beta, r1.d_beta, "d_beta vs one GPU")
    _close(rs.d_alpha, r1.d_alpha, "d_alpha vs one GPU")
    assert rs.d_x.shard_count == shard_set
    # (to the last bit or two: a one-row shard takes the general kernel, whose dot
    # product associates differently from the fused one)
    _close(rs.d_x.to_host(), r1.d_x.to_host(), "d_x vs one GPU")
    assert_logp(rs.logp, o["logp"])
    assert_grad(rs.d_beta, o["d_beta"], "d_beta")
    assert_grad(rs.d_x.to_host(), o["d_x"], "d_x")
    # the factored partial and its reverse sweep into a sharded, lazily zero adjoint
    rf = gpu.bernoulli_logit_glm_lpmf(ys, xs, d["alpha"], d["beta"],
                                      var=("x_factored", "alpha", "beta"))
    adj = gpu.MatrixCuda.like(xs)
    adj.zero_lazy()
    adj.rank1_update(1.0, rf.d_x, d["beta"])
    np.testing.assert_array_equal(adj.to_host(), rs.d_x.to_host())
    # propto with nothing autodiff: 0, and the y check still runs on every shard
    assert gpu.bernoulli_logit_glm_lpmf(ys, xs, d["alpha"], d["beta"], propto=True,
                                        var=()).logp == 0.0
    bad = d["y"].copy()
    bad[-1] = 2
    with pytest.raises(gpu.DomainError):
        gpu.bernoulli_logit_glm_lpmf(gpu.MatrixCuda.from_host_sharded(bad), xs, d["alpha"],
                                     d["beta"])


def test_poisson_vector_alpha_sharded(gpu, shard_set):
    N, K = 20011, 64
    d = make_inputs("poisson", N, K, seed=5, vec_alpha=True)
    xs, ys = gpu.MatrixCuda.from_host_sharded(d["x"]), gpu.MatrixCuda.from_host_sharded(d["y"])
    als = gpu.MatrixCuda.from_host_sharded(d["alpha"])
    rs = gpu.poisson_log_glm_lpmf(ys, xs, als, d["beta"])
    o = po.poisson_log_glm(d["y"], d["x"], d["alpha"], d["beta"])
    assert_logp(rs.logp, o["logp"])
    assert_grad(rs.d_beta, o["d_beta"], "d_beta")
    assert rs.d_alpha.shard_count == shard_set
    assert_grad(rs.d_alpha.to_host().ravel(), o["d_alpha"], "d_alpha")
    # a broadcast scalar y: the reference adds lgamma(y + 1) once per call, not per shard
    r2 = gpu.poisson_log_glm_lpmf(3, xs, als, d["beta"])
    r1 = gpu.poisson_log_glm_lpmf(3, gpu.to_matrix_cuda(d["x"]), gpu.to_matrix_cuda(d["alpha"]),
                                  d["beta"])
    _close(r2.logp, r1.logp, "scalar y")
    # a per-row operand that is not sharded like x is refused
    with pytest.raises(ValueError):
        gpu.poisson_log_glm_lpmf(gpu.to_matrix_cuda(d["y"]), xs, als, d["beta"])
    with pytest.raises(ValueError):
        gpu.poisson_log_glm_lpmf(ys, gpu.to_matrix_cuda(d["x"]), 0.1, d["beta"])


def test_normal_negbin_ordered_binomial_sharded(gpu, shard_set):
    N, K = 5003, 33
    S = gpu.MatrixCuda.from_host_sharded
    d = make_inputs("normal", N, K, seed=8, vec_alpha=True, vec_aux=True)
    r = gpu.normal_id_glm_lpdf(S(d["y"]), S(d["x"]), S(d["alpha"]), d["beta"], S(d["sigma"]),
                               var=("alpha", "beta", "sigma", "y"))
    o = po.normal_id_glm(d["y"], d["x"], d["alpha"], d["beta"], d["sigma"],
                         po.VAR_ALPHA | po.VAR_BETA | po.VAR_AUX | po.VAR_Y)
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "normal d_beta")
    assert_grad(r.d_aux.to_host().ravel(), o["d_sigma"], "normal d_sigma")
    assert_grad(r.d_y.to_host().ravel(), o["d_y"], "normal d_y")
    d = make_inputs("normal", N, K, seed=9)
    r = gpu.normal_id_glm_lpdf(S(d["y"]), S(d["x"]), d["alpha"], d["beta"], d["sigma"],
                               var=("alpha", "beta", "sigma"))
    o = po.normal_id_glm(d["y"], d["x"], d["alpha"], d["beta"], d["sigma"],
                         po.VAR_ALPHA | po.VAR_BETA | po.VAR_AUX)
    assert_logp(r.logp, o["logp"])
    assert_grad(np.atleast_1d(r.d_aux), np.atleast_1d(o["d_sigma"])[:1], "scalar d_sigma",
                scale=N * 1e-2)

    d = make_inputs("neg_binomial", N, K, seed=10)
    r = gpu.neg_binomial_2_log_glm_lpmf(S(d["y"]), S(d["x"]), d["alpha"], d["beta"], d["phi"],
                                        var=("alpha", "beta", "phi"))
    o = po.neg_binomial_2_log_glm(d["y"], d["x"], d["alpha"], d["beta"], d["phi"],
                                  po.VAR_ALPHA | po.VAR_BETA | po.VAR_AUX)
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "negbin d_beta")
    assert_grad(np.atleast_1d(r.d_aux), np.atleast_1d(o["d_phi"])[:1], "d_phi", scale=N * 1e-2)

    d = make_inputs("ordered", N, K, seed=11, C=6)
    r = gpu.ordered_logistic_glm_lpmf(S(d["y"]), S(d["x"]), d["beta"], d["cuts"])
    o = po.ordered_logistic_glm(d["y"], d["x"], d["beta"], d["cuts"])
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "ordered d_beta")
    assert_grad(r.d_aux, o["d_cuts"], "d_cuts", scale=np.abs(o["d_beta"]).max() * 1e-2)

    d = make_inputs("binomial", N, K, seed=12)
    r = gpu.binomial_logit_glm_lpmf(S(d["y"]), S(d["trials"]), S(d["x"]), d["alpha"], d["beta"])
    o = po.binomial_logit_glm(d["y"], d["trials"], d["x"], d["alpha"], d["beta"])
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "binomial d_beta")


@pytest.mark.parametrize("N,K,C", [(4099, 64, 5), (3000, 100, 32), (5, 8, 3)])
def test_categorical_sharded(gpu, shard_set, N, K, C):
    d = make_inputs("categorical", N, K, seed=N + C, C=C)
    S = gpu.MatrixCuda.from_host_sharded
    r = gpu.categorical_logit_glm_lpmf(S(d["y"]), S(d["x"]), d["alpha"], d["beta"],
                                       var=("x", "alpha", "beta"))
    o = po.categorical_logit_glm(d["y"], d["x"], d["alpha"], d["beta"],
                                 flags=po.VAR_X | po.VAR_ALPHA | po.VAR_BETA)
    assert_logp(r.logp, o["logp"])
    sc = max(np.abs(o["d_beta"]).max(), 1e-3)
    assert_grad(r.d_alpha, o["d_alpha"], "d_alpha", scale=sc)
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_x.to_host(), o["d_x"], "d_x")


def test_sharded_matrix_ops(gpu, shard_set):
    rng = np.random.default_rng(3)
    N, K = 1001, 5
    a = np.asfortranarray(rng.standard_normal((N, K)))
    m = gpu.MatrixCuda.from_host_sharded(a)
    np.testing.assert_array_equal(m.rows_to_host(300, 450), a[300:750])
    blk = np.asfortranarray(rng.standard_normal((400, K)))
    m.upload_rows(500, blk)  # straddles shard boundaries
    a[500:900] = blk
    np.testing.assert_array_equal(m.to_host(), a)
    z = gpu.MatrixCuda.like(m)
    z.zero_lazy()
    z.axpy(2.0, m)
    np.testing.assert_array_equal(z.to_host(), 2.0 * a)
    assert z.all_finite()
    y = gpu.MatrixCuda.sharded(N, 1, np.int32)
    y.fill_synthetic(7, kind=1, lo=2, hi=9)
    y1 = gpu.MatrixCuda(N, 1, np.int32)
    y1.fill_synthetic(7, kind=1, lo=2, hi=9)
    np.testing.assert_array_equal(y.to_host(), y1.to_host())  # keyed by the global row
    assert y.int_range() == y1.int_range()
    # no sharded form: refused, not undefined
    y01 = gpu.MatrixCuda.sharded(N, 1, np.int32)
    y01.fill_synthetic(8, kind=1, lo=0, hi=1)
    with pytest.raises((NotImplementedError, ValueError)):
        gpu.lpmf.bernoulli_logit_lpmf(y01, gpu.MatrixCuda.sharded(N, 1))
