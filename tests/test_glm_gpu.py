"""GPU parity of the fused CUDA path against the CPU oracle (through the C ABI).

Mirrors the reference's device-vs-CPU protocol
(test/unit/math/opencl/rev/*_glm_*_test.cpp + test/unit/math/opencl/util.hpp
compare_cpu_opencl_prim_rev): same inputs on both sides, values and every
adjoint compared; cases small_simple, broadcast_y, zero_instances,
small_vector_alpha, big (N=153, K=71 -- deliberately ragged), error_checking,
both propto settings -- with tolerances tighter than the reference's rel 1e-8."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import assert_grad, assert_logp, make_inputs

pytestmark = pytest.mark.gpu

SHAPES = [(3, 2), (153, 71), (1000, 1), (4097, 32), (777, 33), (2500, 100),
          (5000, 128), (3001, 200), (20000, 256)]


def _flags(propto, names):
    f = po.PROPTO if propto else 0
    m = {"x": po.VAR_X, "alpha": po.VAR_ALPHA, "beta": po.VAR_BETA, "aux": po.VAR_AUX,
         "y": po.VAR_Y}
    for n in names:
        f |= m[n]
    return f


@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("propto", [False, True])
@pytest.mark.parametrize("fam", ["bernoulli", "poisson"])
def test_bernoulli_poisson(gpu, fam, N, K, propto):
    d = make_inputs(fam, N, K, seed=N * 31 + K)
    fn_gpu = gpu.bernoulli_logit_glm_lpmf if fam == "bernoulli" else gpu.poisson_log_glm_lpmf
    fn_cpu = po.bernoulli_logit_glm if fam == "bernoulli" else po.poisson_log_glm
    x = gpu.to_matrix_cuda(d["x"])
    y = gpu.to_matrix_cuda(d["y"])
    r = fn_gpu(y, x, d["alpha"], d["beta"], propto=propto, var=("x", "alpha", "beta"))
    o = fn_cpu(d["y"], d["x"], d["alpha"], d["beta"],
               flags=_flags(propto, ["x", "alpha", "beta"]))
    assert o["rc"] == 0
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_alpha, o["d_alpha"][0], "d_alpha", scale=np.abs(o["d_beta"]).max())
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_x.to_host(), o["d_x"], "d_x")


@pytest.mark.parametrize("fam", ["bernoulli", "poisson"])
def test_vector_alpha_and_broadcast_y(gpu, fam):
    N, K = 1531, 71
    d = make_inputs(fam, N, K, seed=7, vec_alpha=True)
    fn_gpu = gpu.bernoulli_logit_glm_lpmf if fam == "bernoulli" else gpu.poisson_log_glm_lpmf
    fn_cpu = po.bernoulli_logit_glm if fam == "bernoulli" else po.poisson_log_glm
    x = gpu.to_matrix_cuda(d["x"])
    a = gpu.to_matrix_cuda(d["alpha"])
    r = fn_gpu(gpu.to_matrix_cuda(d["y"]), x, a, d["beta"])
    o = fn_cpu(d["y"], d["x"], d["alpha"], d["beta"])
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_alpha.to_host().ravel(), o["d_alpha"], "d_alpha_vec")
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    # broadcast_y: a scalar y used for every instance
    r = fn_gpu(1, x, 0.25, d["beta"])
    o = fn_cpu([1], d["x"], 0.25, d["beta"])
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_alpha, o["d_alpha"][0], "d_alpha", scale=np.abs(o["d_beta"]).max())


def test_bernoulli_cutoff_branches(gpu):
    """|theta| > 20 takes the Taylor branches; the > 20 derivative branch is
    -exp(-ytheta) whatever the sign (reference quirk).
    Inputs: rev/prob/bernoulli_logit_glm_lpmf_test.cpp L15-20."""
    x = np.array([[-12, 46], [-42, 24], [25, 27]], float)
    r = gpu.bernoulli_logit_glm_lpmf(gpu.to_matrix_cuda(np.array([1, 0, 1], np.int32)),
                                     gpu.to_matrix_cuda(x), 0.3, [0.3, 2.0],
                                     var=("x", "alpha", "beta"))
    assert_logp(r.logp, -35.699999999999996)
    assert_grad(r.d_alpha, -1.0)
    assert_grad(r.d_beta, [42.0, -24.0])
    want = [-9.0198789646914991e-40, -0.3, -4.3423244813222865e-28,
            -6.0132526431276658e-39, -2.0, -2.8948829875481913e-27]
    got = r.d_x.to_host().ravel(order="F")
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=0)


@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("vec", [False, True])
def test_normal(gpu, N, K, vec):
    d = make_inputs("normal", N, K, seed=N + K, vec_alpha=vec, vec_aux=vec)
    x = gpu.to_matrix_cuda(d["x"])
    y = gpu.to_matrix_cuda(d["y"])
    al = gpu.to_matrix_cuda(d["alpha"]) if vec else d["alpha"]
    sg = gpu.to_matrix_cuda(d["sigma"]) if vec else d["sigma"]
    for propto in (False, True):
        r = gpu.normal_id_glm_lpdf(y, x, al, d["beta"], sg, propto=propto,
                                   var=("x", "y", "alpha", "beta", "sigma"))
        o = po.normal_id_glm(d["y"], d["x"], d["alpha"], d["beta"], d["sigma"],
                             flags=_flags(propto, ["x", "y", "alpha", "beta", "aux"]))
        assert o["rc"] == 0
        assert_logp(r.logp, o["logp"])
        sc = np.abs(o["d_beta"]).max()
        if vec:
            assert_grad(r.d_alpha.to_host().ravel(), o["d_alpha"], "d_alpha")
            assert_grad(r.d_aux.to_host().ravel(), o["d_sigma"], "d_sigma")
        else:
            assert_grad(r.d_alpha, o["d_alpha"][0], "d_alpha", scale=sc)
            assert_grad(r.d_aux, o["d_sigma"][0], "d_sigma", scale=N)
        assert_grad(r.d_y.to_host().ravel(), o["d_y"], "d_y")
        assert_grad(r.d_beta, o["d_beta"], "d_beta")
        assert_grad(r.d_x.to_host(), o["d_x"], "d_x")


@pytest.mark.parametrize("N,K", SHAPES)
@pytest.mark.parametrize("vec", [False, True])
def test_neg_binomial(gpu, N, K, vec):
    d = make_inputs("neg_binomial", N, K, seed=N + 3 * K, vec_alpha=vec, vec_aux=vec)
    x = gpu.to_matrix_cuda(d["x"])
    y = gpu.to_matrix_cuda(d["y"])
    al = gpu.to_matrix_cuda(d["alpha"]) if vec else d["alpha"]
    ph = gpu.to_matrix_cuda(d["phi"]) if vec else d["phi"]
    for propto in (False, True):
        r = gpu.neg_binomial_2_log_glm_lpmf(y, x, al, d["beta"], ph, propto=propto,
                                            var=("x", "alpha", "beta", "phi"))
        o = po.neg_binomial_2_log_glm(d["y"], d["x"], d["alpha"], d["beta"], d["phi"],
                                      flags=_flags(propto, ["x", "alpha", "beta", "aux"]))
        assert o["rc"] == 0
        assert_logp(r.logp, o["logp"])
        sc = np.abs(o["d_beta"]).max()
        if vec:
            assert_grad(r.d_alpha.to_host().ravel(), o["d_alpha"], "d_alpha")
            assert_grad(r.d_aux.to_host().ravel(), o["d_phi"], "d_phi")
        else:
            assert_grad(r.d_alpha, o["d_alpha"][0], "d_alpha", scale=sc)
            assert_grad(r.d_aux, o["d_phi"][0], "d_phi", scale=N * 1e-2)
        assert_grad(r.d_beta, o["d_beta"], "d_beta")
        assert_grad(r.d_x.to_host(), o["d_x"], "d_x")


@pytest.mark.parametrize("N,K,C", [(5, 2, 4), (153, 71, 43), (5000, 64, 9),
                                   (20000, 256, 3), (3000, 3, 2), (999, 40, 150)])
def test_ordered(gpu, N, K, C):
    d = make_inputs("ordered", N, K, seed=N + C, C=C)
    x = gpu.to_matrix_cuda(d["x"])
    y = gpu.to_matrix_cuda(d["y"])
    for propto in (False, True):
        r = gpu.ordered_logistic_glm_lpmf(y, x, d["beta"], d["cuts"], propto=propto,
                                          var=("x", "beta", "cuts"))
        o = po.ordered_logistic_glm(d["y"], d["x"], d["beta"], d["cuts"],
                                    flags=_flags(propto, ["x", "beta", "aux"]))
        assert o["rc"] == 0
        assert_logp(r.logp, o["logp"])
        assert_grad(r.d_beta, o["d_beta"], "d_beta")
        assert_grad(r.d_aux, o["d_cuts"], "d_cuts", scale=np.abs(o["d_beta"]).max() * 1e-2)
        assert_grad(r.d_x.to_host(), o["d_x"], "d_x")


@pytest.mark.parametrize("cuts,scale", [
    # gaps of 1e-3: lim = 1e3 gap - |c1| is <= 0.5, so most rows of those classes take
    # the per-row route (closer cut points make prim's own formula ill-conditioned:
    # the term moves by ulp(loc) / gap with the summation order of x * beta)
    ([-1.0, -1.0 + 1e-3, 0.5, 0.5 + 2e-3], 1.0),
    ([-0.8, -0.79, 0.3, 0.31, 1.0], 8.0),           # gap 1e-2: |loc| up to ~30 crosses it
    ([-2.0, -0.5, 0.4, 1.7], 40.0),                 # wide gaps, saturated predictors
])
def test_ordered_interior_class_term(gpu, cuts, scale):
    """log1m_exp(cut1 - cut2) of an interior class (L160) comes from a per-class table
    only while the rounding of (loc - c1) - (loc - c2) cannot matter; rows beyond the
    table's limit (close cut points, large |loc|) evaluate it per row as prim does.
    Both routes against the oracle, which follows prim per row."""
    N, K = 20011, 64
    cuts = np.asarray(cuts)
    d = make_inputs("ordered", N, K, seed=7, C=len(cuts) + 1)
    beta = d["beta"] * scale
    x = gpu.to_matrix_cuda(d["x"])
    y = gpu.to_matrix_cuda(d["y"])
    r = gpu.ordered_logistic_glm_lpmf(y, x, beta, cuts, var=("beta", "cuts"))
    o = po.ordered_logistic_glm(d["y"], d["x"], beta, cuts, flags=_flags(False, ["beta", "aux"]))
    assert o["rc"] == 0 and np.isfinite(o["logp"])
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_aux, o["d_cuts"], "d_cuts", scale=np.abs(o["d_beta"]).max() * 1e-2)


def test_zero_instances_and_errors(gpu):
    """zero_instances + error_checking cases of the reference's device tests."""
    x0 = gpu.MatrixCuda(0, 2)
    y0 = gpu.MatrixCuda(0, 1, np.int32)
    assert gpu.bernoulli_logit_glm_lpmf(y0, x0, 0.3, [0.1, 0.2]).logp == 0.0
    assert gpu.poisson_log_glm_lpmf(y0, x0, 0.3, [0.1, 0.2]).logp == 0.0
    x = gpu.to_matrix_cuda(np.array([[-12, 46], [-42, 24], [25, 27]], float))
    y = gpu.to_matrix_cuda(np.array([1, 0, 1], np.int32))
    with pytest.raises(ValueError):  # size mismatch -> std::invalid_argument
        gpu.bernoulli_logit_glm_lpmf(y, x, 0.3, [0.3, 2.0, 1.0])
    with pytest.raises(ValueError):
        gpu.bernoulli_logit_glm_lpmf(gpu.to_matrix_cuda(np.array([1, 0], np.int32)), x,
                                     0.3, [0.3, 2.0])
    with pytest.raises(gpu.DomainError):  # y out of range -> std::domain_error
        gpu.bernoulli_logit_glm_lpmf(gpu.to_matrix_cuda(np.array([1, 2, 0], np.int32)),
                                     x, 0.3, [0.3, 2.0])
    with pytest.raises(gpu.DomainError):
        gpu.poisson_log_glm_lpmf(gpu.to_matrix_cuda(np.array([1, -1, 0], np.int32)),
                                 x, 0.3, [0.3, 2.0])
    for bad in (np.inf, -np.inf, np.nan):
        with pytest.raises(gpu.DomainError):  # non-finite beta / alpha / x
            gpu.bernoulli_logit_glm_lpmf(y, x, 0.3, [0.3, bad])
        with pytest.raises(gpu.DomainError):
            gpu.bernoulli_logit_glm_lpmf(y, x, bad, [0.3, 2.0])
        xb = np.array([[-12, 46], [-42, bad], [25, 27]], float)
        if bad == -np.inf:
            # the reference only looks when logp is non-finite (lazy check,
            # L128-132): -inf must sit in a y = 1 row to drive ytheta to -inf
            xb = np.array([[-12, bad], [-42, 24], [25, 27]], float)
        with pytest.raises(gpu.DomainError):
            gpu.bernoulli_logit_glm_lpmf(y, gpu.to_matrix_cuda(xb), 0.3, [0.3, 2.0])
        with pytest.raises(gpu.DomainError):
            gpu.neg_binomial_2_log_glm_lpmf(y, gpu.to_matrix_cuda(xb), 0.3, [0.3, 2.0], 2.0)
        yd = gpu.to_matrix_cuda(np.array([1.0, 0.5, 2.0]))
        with pytest.raises(gpu.DomainError):
            gpu.normal_id_glm_lpdf(yd, gpu.to_matrix_cuda(xb), 0.3, [0.3, 2.0], 1.0)
    yd = gpu.to_matrix_cuda(np.array([1.0, 0.5, 2.0]))
    for s in (0.0, -1.0, np.inf, np.nan):
        with pytest.raises(gpu.DomainError):
            gpu.normal_id_glm_lpdf(yd, x, 0.3, [0.3, 2.0], s)
        with pytest.raises(gpu.DomainError):
            gpu.neg_binomial_2_log_glm_lpmf(y, x, 0.3, [0.3, 2.0], s)
    with pytest.raises(gpu.DomainError):  # cuts not ordered
        gpu.ordered_logistic_glm_lpmf(gpu.to_matrix_cuda(np.array([1, 2, 3], np.int32)),
                                      x, [0.3, 2.0], [0.5, 0.2])
    with pytest.raises(gpu.DomainError):  # y outside 1..C
        gpu.ordered_logistic_glm_lpmf(gpu.to_matrix_cuda(np.array([1, 2, 4], np.int32)),
                                      x, [0.3, 2.0], [0.2, 0.5])


def test_generic_path_matches(gpu, monkeypatch):
    """K > 256 and K = 0 go through the general two-pass kernels."""
    d = make_inputs("poisson", 3000, 300, seed=5)
    x = gpu.to_matrix_cuda(d["x"])
    y = gpu.to_matrix_cuda(d["y"])
    r = gpu.poisson_log_glm_lpmf(y, x, 0.1, d["beta"], var=("x", "alpha", "beta"))
    o = po.poisson_log_glm(d["y"], d["x"], 0.1, d["beta"],
                           flags=_flags(False, ["x", "alpha", "beta"]))
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_x.to_host(), o["d_x"], "d_x")
    # zero attributes
    x0 = gpu.MatrixCuda(50, 0)
    y0 = gpu.to_matrix_cuda(d["y"][:50])
    r = gpu.poisson_log_glm_lpmf(y0, x0, 0.2, [])
    o = po.poisson_log_glm(d["y"][:50], np.zeros((50, 0)), 0.2, [])
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_alpha, o["d_alpha"][0])


@pytest.mark.parametrize("N,K", [(3001, 257), (2050, 513), (1203, 1000)])
def test_wide_x_column_chunks(gpu, monkeypatch, N, K):
    """K > 256: column chunks through the fused kernel (forward theta chain, the
    family on the last chunk, reverse x^T d), every family, against the oracle and
    against the general two-pass kernels (SMC_FORCE_GENERIC)."""
    mb = gpu
    # bernoulli with a vector alpha as an autodiff variable and x var
    d = make_inputs("bernoulli", N, K, seed=N + K, vec_alpha=True)
    x, y = mb.to_matrix_cuda(d["x"]), mb.to_matrix_cuda(d["y"])
    av = mb.to_matrix_cuda(d["alpha"])
    r = mb.bernoulli_logit_glm_lpmf(y, x, av, d["beta"], var=("x", "alpha", "beta"))
    o = po.bernoulli_logit_glm(d["y"], d["x"], d["alpha"], d["beta"],
                               flags=_flags(False, ["x", "alpha", "beta"]))
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_alpha.to_host().ravel(), o["d_alpha"], "d_alpha")
    assert_grad(r.d_x.to_host(), o["d_x"], "d_x")
    monkeypatch.setenv("SMC_FORCE_GENERIC", "1")
    rg = mb.bernoulli_logit_glm_lpmf(y, x, av, d["beta"], var=("x", "alpha", "beta"))
    monkeypatch.delenv("SMC_FORCE_GENERIC")
    assert_logp(r.logp, rg.logp)
    assert_grad(r.d_beta, rg.d_beta, "d_beta vs two-pass")
    # normal: scalar alpha / sigma, beta only
    d = make_inputs("normal", N, K, seed=N)
    x, y = mb.to_matrix_cuda(d["x"]), mb.to_matrix_cuda(d["y"])
    r = mb.normal_id_glm_lpdf(y, x, d["alpha"], d["beta"], d["sigma"])
    o = po.normal_id_glm(d["y"], d["x"], d["alpha"], d["beta"], d["sigma"])
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_alpha, o["d_alpha"][0], "d_alpha", scale=np.abs(o["d_beta"]).max())
    assert_grad(r.d_aux, o["d_sigma"][0], "d_sigma", scale=np.abs(o["d_beta"]).max())
    # neg-binomial with phi var, value only under propto
    d = make_inputs("neg_binomial", N, K, seed=K)
    x, y = mb.to_matrix_cuda(d["x"]), mb.to_matrix_cuda(d["y"])
    for propto in (False, True):
        r = mb.neg_binomial_2_log_glm_lpmf(y, x, d["alpha"], d["beta"], d["phi"],
                                           propto=propto)
        o = po.neg_binomial_2_log_glm(d["y"], d["x"], d["alpha"], d["beta"], d["phi"],
                                      flags=_flags(propto, ["alpha", "beta", "aux"]))
        assert_logp(r.logp, o["logp"])
        assert_grad(r.d_beta, o["d_beta"], "d_beta")
        assert_grad(r.d_aux, np.asarray(o["d_phi"]).ravel()[0], "d_phi",
                    scale=np.abs(o["d_beta"]).max())
    # ordered: the cut points ride with the last chunk
    d = make_inputs("ordered", N, K, seed=7, C=6)
    x, y = mb.to_matrix_cuda(d["x"]), mb.to_matrix_cuda(d["y"])
    r = mb.ordered_logistic_glm_lpmf(y, x, d["beta"], d["cuts"], var=("x", "beta", "cuts"))
    o = po.ordered_logistic_glm(d["y"], d["x"], d["beta"], d["cuts"],
                                flags=_flags(False, ["x", "beta", "aux"]))
    assert_logp(r.logp, o["logp"])
    assert_grad(r.d_beta, o["d_beta"], "d_beta")
    assert_grad(r.d_aux, o["d_cuts"], "d_cuts", scale=np.abs(o["d_beta"]).max() * 1e-2)
    assert_grad(r.d_x.to_host(), o["d_x"], "d_x")
    # value only (no reverse launches): the family call finishes the evaluation
    r = mb.ordered_logistic_glm_lpmf(y, x, d["beta"], d["cuts"], var=())
    assert_logp(r.logp, o["logp"])


def test_synthetic_fill_matches_host(gpu):
    m = gpu.MatrixCuda(1000, 7)
    m.fill_synthetic(12345, row0=17, kind=0, scale=1.0)
    assert np.array_equal(m.to_host(), gpu.synthetic_host(12345, 17, 1000, 7))
    yi = gpu.MatrixCuda(1000, 1, np.int32)
    yi.fill_synthetic(99, row0=5, kind=1, lo=0, hi=4)
    assert np.array_equal(yi.to_host(), gpu.synthetic_host(99, 5, 1000, 1, kind=1, lo=0, hi=4))


def test_determinism(gpu):
    d = make_inputs("bernoulli", 50000, 256, seed=11)
    x = gpu.to_matrix_cuda(d["x"])
    y = gpu.to_matrix_cuda(d["y"])
    r0 = gpu.bernoulli_logit_glm_lpmf(y, x, 0.1, d["beta"])
    for _ in range(5):
        r = gpu.bernoulli_logit_glm_lpmf(y, x, 0.1, d["beta"])
        assert r.logp == r0.logp and np.array_equal(r.d_beta, r0.d_beta)


@pytest.mark.parametrize("N,K,C", [(5, 2, 3), (153, 71, 43), (1000, 1, 2), (4099, 64, 8),
                                   (3000, 512, 32), (2000, 100, 64), (50, 600, 5),
                                   (700, 40, 17), (1531, 37, 65), (777, 21, 130),
                                   (300, 5, 200)])
def test_categorical(gpu, N, K, C):
    d = make_inputs("categorical", N, K, seed=N + 5 * C, C=C)
    x = gpu.to_matrix_cuda(d["x"])
    y = gpu.to_matrix_cuda(d["y"])
    for propto in (False, True):
        r = gpu.categorical_logit_glm_lpmf(y, x, d["alpha"], d["beta"], propto=propto,
                                           var=("x", "alpha", "beta"))
        o = po.categorical_logit_glm(d["y"], d["x"], d["alpha"], d["beta"],
                                     flags=_flags(propto, ["x", "alpha", "beta"]))
        assert o["rc"] == 0
        assert_logp(r.logp, o["logp"])
        sc = np.abs(o["d_beta"]).max()
        assert_grad(r.d_alpha, o["d_alpha"], "d_alpha", scale=sc * 1e-2)
        assert_grad(r.d_beta, o["d_beta"], "d_beta")
        assert_grad(r.d_x.to_host(), o["d_x"], "d_x")
    # C == 1 returns 0 (L73-75); y out of support -> domain_error
    assert gpu.categorical_logit_glm_lpmf(y, x, [0.1], np.zeros((K, 1))).logp == 0.0
    ybad = d["y"].copy()
    ybad[0] = C + 1
    with pytest.raises(gpu.DomainError):
        gpu.categorical_logit_glm_lpmf(gpu.to_matrix_cuda(ybad), x, d["alpha"], d["beta"])
