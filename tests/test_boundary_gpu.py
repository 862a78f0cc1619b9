"""Boundary behaviour of the C ABI added in round 2: row-vector device operands,
cache identity, lazily zeroed adjoints, the factored d_x and its rank-one reverse
sweep, wrapped-buffer invalidation, frees from foreign threads."""
import threading

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import assert_grad, assert_logp, make_inputs

pytestmark = pytest.mark.gpu


def test_row_vector_device_operands(gpu):
    """1 x N device vectors (matrix_cuda<double>(Eigen::RowVectorXd), numpy (1, N))
    are contiguous and read like N x 1 ones (they used to be padded to a
    16-element leading dimension and read as garbage)."""
    N, K = 1237, 19
    d = make_inputs("normal", N, K, seed=3, vec_alpha=True, vec_aux=True)
    x = gpu.to_matrix_cuda(d["x"])
    row = lambda v: gpu.to_matrix_cuda(np.asarray(v).reshape(1, -1))  # noqa: E731
    y, al, sg = row(d["y"]), row(d["alpha"]), row(d["sigma"])
    assert (y.rows, y.cols, y.ld) == (1, N, 1)
    r = gpu.normal_id_glm_lpdf(y, x, al, d["beta"], sg, var=("alpha", "beta", "sigma", "y"))
    o = po.normal_id_glm(d["y"], d["x"], d["alpha"], d["beta"], d["sigma"],
                         po.VAR_ALPHA | po.VAR_BETA | po.VAR_AUX | po.VAR_Y)
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_alpha.to_host().ravel(), o["d_alpha"], "d_alpha")
    assert_grad(r.d_aux.to_host().ravel(), o["d_sigma"], "d_sigma")
    np.testing.assert_array_equal(y.to_host().ravel(), d["y"])
    # integer row vector
    db = make_inputs("bernoulli", N, K, seed=4)
    rb = gpu.bernoulli_logit_glm_lpmf(row(db["y"]), gpu.to_matrix_cuda(db["x"]), db["alpha"],
                                      db["beta"])
    ob = po.bernoulli_logit_glm(db["y"], db["x"], db["alpha"], db["beta"])
    assert_logp(rb.logp, ob["logp"])
    assert_grad(rb.d_beta, ob["d_beta"], "d_beta")


def test_strided_row_operand_is_refused(gpu):
    """A wrapped 1 x N row with a stride is not a per-row operand the kernels can
    index: std::invalid_argument, not silent garbage."""
    N, K = 64, 4
    d = make_inputs("normal", N, K, seed=5)
    x = gpu.to_matrix_cuda(d["x"])
    pool = gpu.MatrixCuda(8, N)  # 8 x N, ld = 16
    strided = gpu.MatrixCuda.wrap(pool.data_ptr, 1, N, pool.ld, np.float64, keep=pool)
    with pytest.raises(ValueError):
        gpu.normal_id_glm_lpdf(strided, x, 0.1, d["beta"], 1.3)


def test_binomial_statistics_follow_the_matrix_not_the_address(gpu):
    """the cached binomial pair statistics were keyed by the partner's data
    pointer + version; a freed trials vector is recycled at the same address with the
    same version, so new values hit the stale cache."""
    N, K = 4001, 7
    d = make_inputs("binomial", N, K, seed=11)
    x = gpu.to_matrix_cuda(d["x"])
    n = gpu.to_matrix_cuda(d["y"])
    t1 = gpu.to_matrix_cuda(d["trials"])
    p1 = t1.data_ptr
    r1 = gpu.binomial_logit_glm_lpmf(n, t1, x, d["alpha"], d["beta"])
    o1 = po.binomial_logit_glm(d["y"], d["trials"], d["x"], d["alpha"], d["beta"])
    assert_logp(r1.logp, o1["logp"])
    del t1
    trials2 = (d["trials"] + 3).astype(np.int32)
    t2 = gpu.to_matrix_cuda(trials2)
    assert t2.data_ptr == p1, "the block cache should recycle the freed block"
    r2 = gpu.binomial_logit_glm_lpmf(n, t2, x, d["alpha"], d["beta"])
    o2 = po.binomial_logit_glm(d["y"], trials2, d["x"], d["alpha"], d["beta"])
    assert_logp(r2.logp, o2["logp"])
    assert abs(r2.logp - r1.logp) > 1.0


def test_lazy_zero(gpu):
    rng = np.random.default_rng(1)
    N, K = 1003, 6
    m = gpu.MatrixCuda(N, K)
    m.fill_synthetic(3, kind=0)
    m.zero_lazy()
    np.testing.assert_array_equal(m.to_host(), np.zeros((N, K)))  # realised by the read
    # a whole-matrix accumulate stores instead of adding to zeros
    xs = np.asfortranarray(rng.standard_normal((N, K)))
    xd = gpu.to_matrix_cuda(xs)
    m.fill_synthetic(4, kind=0)
    m.zero_lazy()
    m.axpy(-2.5, xd)
    np.testing.assert_array_equal(m.to_host(), -2.5 * xs)
    m.axpy(0.5, xd)
    np.testing.assert_allclose(m.to_host(), -2.5 * xs + 0.5 * xs, rtol=0, atol=1e-15)
    # a partial write realises the rest
    m.fill_synthetic(5, kind=0)
    m.zero_lazy()
    blk = np.asfortranarray(rng.standard_normal((10, K)))
    m.upload_rows(5, blk)
    want = np.zeros((N, K))
    want[5:15] = blk
    np.testing.assert_array_equal(m.to_host(), want)
    # the raw pointer escaping realises too
    m.fill_synthetic(6, kind=0)
    m.zero_lazy()
    v = gpu.MatrixCuda.wrap(m.data_ptr, N, K, m.ld, np.float64, keep=m)
    np.testing.assert_array_equal(v.to_host(), np.zeros((N, K)))


@pytest.mark.parametrize("fam", ["bernoulli", "poisson", "normal", "neg_binomial", "ordered",
                                 "binomial"])
@pytest.mark.parametrize("N,K", [(1, 1), (33, 3), (2049, 100), (5000, 256), (3001, 300)])
def test_factored_dx_is_the_same_partial(gpu, fam, N, K):
    """SMC_DX_FACTORED returns d; d beta^T is bit for bit the d_x the full variant
    writes, every other result is unchanged, and rank1_update reproduces
    x.adj() += lp.adj() * d_x (rev/functor/operands_and_partials.hpp L28-38)."""
    d = make_inputs(fam, N, K, seed=N + K, vec_alpha=(fam == "poisson"))
    x = gpu.to_matrix_cuda(d["x"])
    y = gpu.to_matrix_cuda(d["y"])

    def call(xname):
        if fam == "bernoulli":
            return gpu.bernoulli_logit_glm_lpmf(y, x, d["alpha"], d["beta"],
                                                var=(xname, "alpha", "beta"))
        if fam == "poisson":
            return gpu.poisson_log_glm_lpmf(y, x, gpu.to_matrix_cuda(d["alpha"]), d["beta"],
                                            var=(xname, "alpha", "beta"))
        if fam == "normal":
            return gpu.normal_id_glm_lpdf(y, x, d["alpha"], d["beta"], d["sigma"],
                                          var=(xname, "alpha", "beta", "sigma"))
        if fam == "neg_binomial":
            return gpu.neg_binomial_2_log_glm_lpmf(y, x, d["alpha"], d["beta"], d["phi"],
                                                   var=(xname, "alpha", "beta", "phi"))
        if fam == "binomial":
            return gpu.binomial_logit_glm_lpmf(y, gpu.to_matrix_cuda(d["trials"]), x,
                                               d["alpha"], d["beta"],
                                               var=(xname, "alpha", "beta"))
        return gpu.ordered_logistic_glm_lpmf(y, x, d["beta"], d["cuts"],
                                             var=(xname, "beta", "cuts"))

    full, fact = call("x"), call("x_factored")
    assert full.logp == fact.logp
    np.testing.assert_array_equal(full.d_beta, fact.d_beta)
    assert (fact.d_x.rows, fact.d_x.cols) == (N, 1)
    dvec = fact.d_x.to_host().ravel()
    dx = full.d_x.to_host()
    np.testing.assert_array_equal(dx, np.outer(dvec, d["beta"]))
    if fam == "poisson":  # the vector alpha's partial is the same d
        np.testing.assert_array_equal(fact.d_alpha.to_host().ravel(), dvec)
    adj = gpu.MatrixCuda(N, K)
    adj.zero_lazy()
    adj.rank1_update(1.0, fact.d_x, d["beta"])
    np.testing.assert_array_equal(adj.to_host(), dx)  # lp.adj() = 1: the partial itself
    # accumulation: adj + a * partial, to the last bit or two (the device contracts the
    # multiply-add, Eigen may or may not)
    adj.rank1_update(-0.3, fact.d_x, d["beta"])
    np.testing.assert_allclose(adj.to_host(), dx + (-0.3) * dx, rtol=4e-16, atol=0)
    base = np.asfortranarray(np.random.default_rng(9).standard_normal((N, K)))
    adj2 = gpu.to_matrix_cuda(base)
    adj2.rank1_update(0.7, fact.d_x, d["beta"])
    np.testing.assert_allclose(adj2.to_host(), base + 0.7 * dx, rtol=0,
                               atol=4e-16 * float(np.abs(base).max() + np.abs(dx).max()))


def test_invalidate_wrapped_buffer(gpu):
    """The owner of wrapped memory tells the library its contents changed."""
    N = 5000
    owner = gpu.to_matrix_cuda(np.full(N, 3, dtype=np.int32))
    view = gpu.MatrixCuda.wrap(owner.data_ptr, N, 1, N, np.int32, keep=owner)
    assert view.int_range() == (3, 3)
    owner.upload(np.arange(N, dtype=np.int32).reshape(-1, 1))
    assert view.int_range() == (3, 3)  # the view's cache cannot know
    view.invalidate()
    assert view.int_range() == (0, N - 1)


def test_free_from_another_thread(gpu):
    """A matrix created (and used) on one thread and freed on another -- a finalizer
    thread, a TBB worker -- is handed back to the driver, never into the other
    thread's recycling cache while work may still be queued on the creator's stream."""
    d = make_inputs("poisson", 20000, 64, seed=2)
    x = gpu.to_matrix_cuda(d["x"])
    y = gpu.to_matrix_cuda(d["y"])
    o = po.poisson_log_glm(d["y"], d["x"], d["alpha"], d["beta"])
    box = {}
    for _ in range(20):
        r = gpu.poisson_log_glm_lpmf(y, x, d["alpha"], d["beta"], var=("x", "alpha", "beta"))
        box["m"] = r.d_x
        del r

        def drop():
            gpu.runtime.set_device(0)
            box.pop("m")  # last reference dies here, on the foreign thread
            tmp = gpu.MatrixCuda(20000, 64)  # same size: must not be the freed block in flight
            tmp.fill_synthetic(1, kind=0)
        t = threading.Thread(target=drop)
        t.start()
        t.join()
    r = gpu.poisson_log_glm_lpmf(y, x, d["alpha"], d["beta"])
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
