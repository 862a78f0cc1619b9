"""GPU parity of the step either side of the fused GLMs: the
device matrix-vector product and the un-fused densities on a device linear
predictor, against the CPU oracle.

The oracle restates the GLMs; an un-fused density on theta is the same GLM with
the predictor passed through the intercept vector (x = 0), or -- for the ordered
family, which has no intercept -- through a one-column x = theta with beta = 1,
whose d_x is d/dtheta.  The C++ suite (tests/cpp/unfused_lpmf_test.cpp) compares
the same entry points with the reference's own prim *_lpmf implementations."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import assert_grad, assert_logp, make_inputs

pytestmark = pytest.mark.gpu

SIZES = [(1, 1), (33, 7), (4099, 64), (20011, 256), (3001, 300)]


@pytest.mark.parametrize("N,K", SIZES)
def test_multiply_and_adjoint(gpu, N, K):
    mb = gpu
    d = make_inputs("bernoulli", N, K, seed=N + K, vec_alpha=True)
    x = mb.to_matrix_cuda(d["x"])
    theta = mb.lpmf.multiply(x, d["beta"], mb.to_matrix_cuda(d["alpha"]))
    want = d["x"] @ d["beta"] + d["alpha"]
    assert_grad(mb.from_matrix_cuda(theta).ravel(), want, "theta",
                scale=np.abs(d["x"]).max() * np.abs(d["beta"]).sum())
    theta0 = mb.lpmf.multiply(x, d["beta"], 0.25)
    assert_grad(mb.from_matrix_cuda(theta0).ravel(), d["x"] @ d["beta"] + 0.25, "theta0",
                scale=np.abs(d["x"]).max() * np.abs(d["beta"]).sum())
    rng = np.random.default_rng(7)
    v = rng.standard_normal(N)
    g, s = mb.lpmf.multiply_adjoint(x, mb.to_matrix_cuda(v))
    assert_grad(g, d["x"].T @ v, "x^T v", scale=np.abs(d["x"]).max() * np.abs(v).sum())
    assert_logp(s, float(np.sum(v)) if abs(np.sum(v)) > 1e-6 else s, "sum v")
    assert abs(s - np.sum(v)) <= 1e-12 * np.abs(v).sum()
    assert abs(mb.lpmf.vector_sum(mb.to_matrix_cuda(v)) - np.sum(v)) <= 1e-12 * np.abs(v).sum()


def _theta(N, seed, scale=2.0):
    return np.random.default_rng(seed).standard_normal(N) * scale


@pytest.mark.parametrize("N", [1, 257, 50021])
@pytest.mark.parametrize("family", ["bernoulli", "poisson"])
def test_theta_lpmf_matches_oracle(gpu, family, N):
    mb = gpu
    theta = _theta(N, N)
    rng = np.random.default_rng(N + 1)
    y = (rng.integers(0, 2, N) if family == "bernoulli" else rng.integers(0, 7, N)) \
        .astype(np.int32)
    zero_x = np.zeros((N, 1), order="F")
    if family == "bernoulli":
        o = po.bernoulli_logit_glm(y, zero_x, theta, [0.0])
        r = mb.lpmf.bernoulli_logit_lpmf(mb.to_matrix_cuda(y), mb.to_matrix_cuda(theta))
    else:
        o = po.poisson_log_glm(y, zero_x, theta, [0.0])
        r = mb.lpmf.poisson_log_lpmf(mb.to_matrix_cuda(y), mb.to_matrix_cuda(theta))
    assert_logp(r.logp, o["logp"])
    assert_grad(mb.from_matrix_cuda(r.d_theta).ravel(), o["d_alpha"], "d_theta")


@pytest.mark.parametrize("N", [1, 257, 50021])
def test_neg_binomial_lpmf_matches_oracle(gpu, N):
    mb = gpu
    eta = _theta(N, N)
    y = np.random.default_rng(N + 2).integers(0, 9, N).astype(np.int32)
    o = po.neg_binomial_2_log_glm(y, np.zeros((N, 1), order="F"), eta, [0.0], 2.5)
    r = mb.lpmf.neg_binomial_2_log_lpmf(mb.to_matrix_cuda(y), mb.to_matrix_cuda(eta), 2.5)
    assert_logp(r.logp, o["logp"])
    assert_grad(mb.from_matrix_cuda(r.d_theta).ravel(), o["d_alpha"], "d_eta")
    assert_grad(r.d_aux, np.asarray(o["d_phi"]).ravel()[0], "d_phi",
                scale=np.abs(o["d_alpha"]).sum())


@pytest.mark.parametrize("N", [1, 257, 50021])
def test_ordered_logistic_lpmf_matches_oracle(gpu, N):
    mb = gpu
    lam = _theta(N, N, 3.0)
    cuts = np.array([-1.5, -0.2, 0.4, 2.0])
    y = np.random.default_rng(N + 3).integers(1, 6, N).astype(np.int32)
    o = po.ordered_logistic_glm(y, np.asfortranarray(lam.reshape(N, 1)), [1.0], cuts,
                                flags=po.VAR_BETA | po.VAR_AUX | po.VAR_X)
    r = mb.lpmf.ordered_logistic_lpmf(mb.to_matrix_cuda(y), mb.to_matrix_cuda(lam), cuts)
    assert_logp(r.logp, o["logp"])
    assert_grad(mb.from_matrix_cuda(r.d_theta).ravel(), o["d_x"].ravel(), "d_lambda")
    assert_grad(r.d_aux, o["d_cuts"], "d_cuts", scale=np.abs(o["d_x"]).sum())


@pytest.mark.parametrize("N,C", [(1, 5), (257, 2), (300, 43), (20011, 6)])
def test_ordered_logistic_lpmf_one_cut_vector_per_outcome(gpu, N, C):
    """prim/prob/ordered_logistic_lpmf.hpp L72-200 with a std::vector of cut vectors
    (the second form of test/unit/math/opencl/rev/ordered_logistic_lpmf_test.cpp): the
    oracle evaluates outcome i on its own with cut vector i."""
    mb = gpu
    rng = np.random.default_rng(N * 31 + C)
    lam = rng.standard_normal(N) * 2.0
    cuts = np.cumsum(np.abs(rng.standard_normal((C - 1, N))) + 1e-5, axis=0) - 1.5
    y = rng.integers(1, C + 1, N).astype(np.int32)
    y[0] = 1
    y[-1] = C
    n_chk = min(N, 400)  # the oracle runs one call per outcome
    rows = np.unique(np.concatenate([[0, N - 1], rng.integers(0, N, n_chk)]))
    cu = mb.to_matrix_cuda(np.asfortranarray(cuts))
    r = mb.lpmf.ordered_logistic_lpmf_rows(mb.to_matrix_cuda(y), mb.to_matrix_cuda(lam), cu)
    d_lam = mb.from_matrix_cuda(r.d_theta).ravel()
    d_cuts = mb.from_matrix_cuda(r.d_aux)
    assert d_cuts.shape == cuts.shape
    logp_rows = 0.0
    for i in (rows if N > n_chk else range(N)):
        o = po.ordered_logistic_glm(y[i:i + 1], np.array([[lam[i]]], order="F"), [1.0],
                                    cuts[:, i], flags=po.VAR_BETA | po.VAR_AUX | po.VAR_X)
        logp_rows += o["logp"]
        assert_grad(d_lam[i:i + 1], o["d_x"].ravel(), "d_lambda")
        assert_grad(d_cuts[:, i], np.asarray(o["d_cuts"]).ravel(), "d_cuts", scale=1.0)
    if N <= n_chk:
        assert_logp(r.logp, logp_rows)
    else:
        # all rows: numpy restatement of L127-160 (checked against the oracle on the
        # sampled rows through the gradients above and on the small cases exactly)
        c1 = np.where(y == C, np.inf, cuts[np.minimum(y - 1, C - 2), np.arange(N)])
        c2 = np.where(y == 1, -np.inf, cuts[np.maximum(y - 2, 0), np.arange(N)])
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            A = -np.logaddexp(0.0, lam - c1)
            B = -np.logaddexp(0.0, -(lam - c2))
            mid = np.log1p(-np.exp(np.minimum((lam - c1) - (lam - c2), -1e-300)))
        tot = np.where(y == 1, A, np.where(y == C, B, A + B + mid)).sum()
        assert abs(r.logp - tot) <= 1e-10 * abs(tot)
    # a (C-1) x 1 device matrix is the one-vector form
    one = mb.lpmf.ordered_logistic_lpmf_rows(mb.to_matrix_cuda(y), mb.to_matrix_cuda(lam),
                                             mb.to_matrix_cuda(np.asfortranarray(cuts[:, :1])))
    ref = mb.lpmf.ordered_logistic_lpmf(mb.to_matrix_cuda(y), mb.to_matrix_cuda(lam), cuts[:, 0])
    assert one.logp == ref.logp
    np.testing.assert_array_equal(mb.from_matrix_cuda(one.d_aux).ravel(), ref.d_aux)
    # propto with nothing autodiff: checks only
    z = mb.lpmf.ordered_logistic_lpmf_rows(mb.to_matrix_cuda(y), mb.to_matrix_cuda(lam), cu,
                                           propto=True, theta_var=False, cuts_var=False)
    assert z.logp == 0.0


def test_ordered_logistic_lpmf_rows_errors(gpu):
    mb = gpu
    y = mb.to_matrix_cuda(np.array([1, 3, 2], dtype=np.int32))
    lam = mb.to_matrix_cuda(np.array([0.3, 2.0, -0.3]))
    good = np.asfortranarray(np.array([[-0.3, 0.8, 1.8, 3.0], [-0.6, 1.8, 2.4, 3.2],
                                       [-0.3, 0.8, 1.8, 3.0]]).T)
    mb.lpmf.ordered_logistic_lpmf_rows(y, lam, mb.to_matrix_cuda(good))
    for bad_col in ([-0.3, -0.8, 3.0, 4.0], [-0.3, 0.8, 3.0, np.inf], [-np.inf, 0.8, 1.8, 3.0],
                    [np.nan, 0.8, 1.8, 3.0]):
        bad = good.copy()
        bad[:, 1] = bad_col
        with pytest.raises(mb.DomainError):
            mb.lpmf.ordered_logistic_lpmf_rows(y, lam, mb.to_matrix_cuda(bad))
    with pytest.raises(mb.DomainError):
        mb.lpmf.ordered_logistic_lpmf_rows(mb.to_matrix_cuda(np.array([1, 2, 6], dtype=np.int32)),
                                           lam, mb.to_matrix_cuda(good))
    with pytest.raises(mb.DomainError):
        mb.lpmf.ordered_logistic_lpmf_rows(y, mb.to_matrix_cuda(np.array([0.3, 2.0, np.inf])),
                                           mb.to_matrix_cuda(good))
    with pytest.raises(ValueError):
        mb.lpmf.ordered_logistic_lpmf_rows(y, lam, mb.to_matrix_cuda(good[:, :2].copy(order="F")))


@pytest.mark.parametrize("N", [1, 257, 50021])
def test_normal_lpdf_matches_oracle(gpu, N):
    mb = gpu
    rng = np.random.default_rng(N + 4)
    mu = rng.standard_normal(N) * 2
    y = mu + 1.3 * rng.standard_normal(N)
    o = po.normal_id_glm(y, np.zeros((N, 1), order="F"), mu, [0.0], 1.3)
    logp, _, d_mu, d_sigma = mb.lpmf.normal_lpdf(mb.to_matrix_cuda(y),
                                                 mb.to_matrix_cuda(mu), 1.3)
    assert_logp(logp, o["logp"])
    assert_grad(mb.from_matrix_cuda(d_mu).ravel(), o["d_alpha"], "d_mu")
    assert_grad(d_sigma, np.asarray(o["d_sigma"]).ravel()[0], "d_sigma",
                scale=np.abs(o["d_alpha"]).sum())


@pytest.mark.parametrize("N,G", [(1, 1), (1000, 7), (50021, 300), (200003, 2048), (30011, 5000)])
def test_indexing_and_reverse(gpu, N, G):
    mb = gpu
    rng = np.random.default_rng(N + G)
    z = rng.standard_normal(G)
    idx = rng.integers(0, G, N).astype(np.int32)
    idx_d = mb.to_matrix_cuda(idx)
    out = mb.from_matrix_cuda(mb.lpmf.indexing(z, idx_d)).ravel()
    assert np.array_equal(out, z[idx])
    v = rng.standard_normal(N)
    v_d = mb.to_matrix_cuda(v)
    g1 = mb.lpmf.indexing_rev(idx_d, v_d, G)
    g2 = mb.lpmf.indexing_rev(idx_d, v_d, G)
    assert np.array_equal(g1, g2)  # deterministic
    want = np.zeros(G)
    np.add.at(want, idx, v)
    assert_grad(g1, want, "indexing_rev", scale=np.abs(v).sum() / max(G, 1))
    with pytest.raises(mb.DomainError):
        mb.lpmf.indexing(z, mb.to_matrix_cuda(np.array([0, G], dtype=np.int32)))


def test_indexing_rev_many_groups_follows_reupload(gpu):
    """More groups than the shared-memory accumulators hold: the reverse sweep runs
    from a sorted row list cached on the index vector; a new upload into the same
    device vector must invalidate it."""
    mb = gpu
    N, G = 400_003, 50_000
    rng = np.random.default_rng(17)
    v = rng.standard_normal(N)
    v_d = mb.to_matrix_cuda(v)
    idx = rng.integers(0, G, N).astype(np.int32)
    idx_d = mb.to_matrix_cuda(idx)
    for _ in range(2):
        g1 = mb.lpmf.indexing_rev(idx_d, v_d, G)
        g2 = mb.lpmf.indexing_rev(idx_d, v_d, G)
        assert np.array_equal(g1, g2)
        want = np.zeros(G)
        np.add.at(want, idx, v)
        assert_grad(g1, want, "indexing_rev", scale=np.abs(v).sum() / G)
        idx = rng.permutation(idx).astype(np.int32)  # second round: new contents
        idx_d.upload_rows(0, idx)
    # a different group count on the same vector (some groups empty)
    g3 = mb.lpmf.indexing_rev(idx_d, v_d, G + 1000)
    assert g3.shape == (G + 1000,) and np.all(g3[G:] == 0.0)


def test_unfused_pipeline_equals_fused_glm(gpu):
    """multiply -> density -> multiply_adjoint gives the fused GLM's value and
    gradient (three sweeps instead of one: what the fusion buys is in DESIGN.md)."""
    mb = gpu
    N, K = 30011, 128
    d = make_inputs("bernoulli", N, K, seed=99)
    x, y = mb.to_matrix_cuda(d["x"]), mb.to_matrix_cuda(d["y"])
    fused = mb.bernoulli_logit_glm_lpmf(y, x, d["alpha"], d["beta"])
    theta = mb.lpmf.multiply(x, d["beta"], d["alpha"])
    un = mb.lpmf.bernoulli_logit_lpmf(y, theta)
    g, s = mb.lpmf.multiply_adjoint(x, un.d_theta)
    assert_logp(un.logp, fused.logp)
    assert_grad(g, fused.d_beta, "d_beta")
    assert_grad(s, fused.d_alpha, "d_alpha", scale=np.abs(fused.d_beta).max())


def test_value_checks(gpu):
    mb = gpu
    y = mb.to_matrix_cuda(np.array([1, 0, 1], dtype=np.int32))
    with pytest.raises(mb.DomainError):
        mb.lpmf.bernoulli_logit_lpmf(y, mb.to_matrix_cuda(np.array([0.1, np.nan, 1.0])))
    r = mb.lpmf.bernoulli_logit_lpmf(y, mb.to_matrix_cuda(np.array([np.inf, -np.inf, 1.0])))
    assert np.isfinite(r.logp)
    with pytest.raises(ValueError):
        mb.lpmf.bernoulli_logit_lpmf(y, mb.to_matrix_cuda(np.zeros(4)))
    r = mb.lpmf.poisson_log_lpmf(y, mb.to_matrix_cuda(np.array([0.1, np.inf, 1.0])))
    assert r.logp == -np.inf
    assert np.all(mb.from_matrix_cuda(r.d_theta) == 0)
    with pytest.raises(mb.DomainError):
        mb.lpmf.neg_binomial_2_log_lpmf(y, mb.to_matrix_cuda(np.array([0.1, np.inf, 1.0])), 2.0)
    with pytest.raises(mb.DomainError):
        mb.lpmf.ordered_logistic_lpmf(y, mb.to_matrix_cuda(np.zeros(3)), [0.5, 0.1])
