"""CPU: pin the plain-C oracle (oracle/glm_oracle.c) against the reference.

(a) known answers of SURVEY.md 8(c) / the fixed inputs of
    test/unit/math/rev/prob/*_glm_*_test.cpp,
(b) tests/golden/glm_golden.json, produced by the unmodified reference
    (oracle/ref_driver.cpp, tests/golden/make_golden.py),
(c) the reference itself, live, when oracle/_ref/libstan_ref.so is present,
(d) the scalar kernels (digamma, lgamma, log1p_exp, log1m_exp) pointwise.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import golden_util as gu
from tests.util import assert_grad, assert_logp, make_inputs

CASES = gu.load()


def run_oracle(fam, d, flags, impl="oracle"):
    if fam == "bernoulli":
        return po.bernoulli_logit_glm(d["y"], d["x"], d["alpha"], d["beta"], flags, impl)
    if fam == "poisson":
        return po.poisson_log_glm(d["y"], d["x"], d["alpha"], d["beta"], flags, impl)
    if fam == "normal":
        return po.normal_id_glm(d["y"], d["x"], d["alpha"], d["beta"], d["sigma"], flags, impl)
    if fam == "neg_binomial":
        return po.neg_binomial_2_log_glm(d["y"], d["x"], d["alpha"], d["beta"], d["phi"],
                                         flags, impl)
    if fam == "binomial":
        return po.binomial_logit_glm(d["y"], d["trials"], d["x"], d["alpha"], d["beta"],
                                     flags, impl)
    if fam == "ordered":
        return po.ordered_logistic_glm(d["y"], d["x"], d["beta"], d["cuts"], flags, impl)
    return po.categorical_logit_glm(d["y"], d["x"], d["alpha"], d["beta"], flags, impl)


def compare(got, expect, N):
    assert_logp(got["logp"], expect["logp"])
    scale = max(float(np.max(np.abs(expect["d_beta"]))) if len(np.atleast_1d(expect["d_beta"])) else 0.0, 1e-300)
    for k, want in expect.items():
        if k == "logp" or want is None:
            continue
        g = got[k]
        if g is None:
            continue
        assert_grad(np.asarray(g).ravel(order="F"), np.atleast_1d(want), k, scale=scale * 1e-2)


@pytest.mark.parametrize("case", CASES, ids=gu.ids(CASES))
def test_oracle_matches_golden(case):
    d = gu.inputs_of(case)
    for v in case["variants"]:
        r = run_oracle(case["family"], d, v["flags"])
        assert r["rc"] == 0
        compare(r, v["expect"], case["shape"][0])


def test_known_answers_survey_8c():
    x = np.array([[-12, 46], [-42, 24], [25, 27]], float)
    r = po.bernoulli_logit_glm([1, 0, 1], x, 0.3, [0.3, 2.0],
                               flags=po.VAR_X | po.VAR_ALPHA | po.VAR_BETA)
    assert r["logp"] == pytest.approx(-35.699999999999996, rel=1e-15)
    assert r["d_alpha"][0] == pytest.approx(-1.0, rel=1e-14)
    np.testing.assert_allclose(r["d_beta"], [42.0, -24.0], rtol=1e-14)
    np.testing.assert_allclose(
        r["d_x"].ravel(order="F"),
        [-9.0198789646914991e-40, -0.3, -4.3423244813222865e-28,
         -6.0132526431276658e-39, -2.0, -2.8948829875481913e-27], rtol=1e-13)
    r = po.normal_id_glm([14, 32, 21], x, 0.3, [0.3, 2.0], 10.0)
    assert r["logp"] == pytest.approx(-45.956670878596157, rel=1e-14)
    assert r["d_sigma"][0] == pytest.approx(6.9584200000000012, rel=1e-13)
    np.testing.assert_allclose(r["d_beta"], [0.31800000000000139, -46.265999999999998],
                               rtol=1e-12, atol=1e-13)
    r = po.poisson_log_glm([14, 2, 5], x / 100, 0.3, [0.3, 2.0])
    assert r["logp"] == pytest.approx(-15.900271464537942, rel=1e-14)
    np.testing.assert_allclose(r["d_beta"], [-0.69435197905709523, 5.631286107135665],
                               rtol=1e-13)
    r = po.neg_binomial_2_log_glm([14, 2, 5], x / 100, 0.3, [0.3, 2.0], 2.0)
    assert r["logp"] == pytest.approx(-10.359512037713642, rel=1e-14)
    assert r["d_phi"][0] == pytest.approx(-0.46459289780560997, rel=1e-12)
    xo = np.array([[1, 2], [3, 4], [5, 6], [7, 8], [9, 0]], float)
    r = po.ordered_logistic_glm([1, 1, 2, 4, 4], xo, [1.1, 0.4], [0.9, 1.1, 7.0])
    assert r["logp"] == pytest.approx(-13.914810581699157, rel=1e-14)
    np.testing.assert_allclose(
        r["d_cuts"], [-2.804494248653481, 5.5155430300941335, -0.071993868812495185],
        rtol=1e-13)
    xc = np.array([[-12, 46], [-42, 24], [25, 27], [-14, -11], [5, 18]], float)
    r = po.categorical_logit_glm([1, 3, 1, 2, 2], xc, [0.5, -2.0, 4.0],
                                 np.array([[0.3, 2, 0.4], [-0.1, -1.3, 1]]))
    assert r["logp"] == pytest.approx(-141.10004744408855, rel=1e-14)
    np.testing.assert_allclose(
        r["d_beta"].ravel(order="F"),
        [26.999335799326794, 83.999478127000543, -8.9999713681453422,
         7.0000224964526279, -17.999364431181448, -90.999500623453187], rtol=1e-13)


def test_oracle_error_codes():
    """Exception kinds of the reference tests: size -> invalid_argument (1),
    value -> domain_error (2)."""
    x = np.array([[-12, 46], [-42, 24], [25, 27]], float)
    assert po.bernoulli_logit_glm([1, 2, 0], x, 0.3, [0.3, 2.0])["rc"] == 2
    assert po.bernoulli_logit_glm([1, 0], x, 0.3, [0.3, 2.0])["rc"] == 1
    assert po.bernoulli_logit_glm([1, 0, 1], x, np.inf, [0.3, 2.0])["rc"] == 2
    assert po.poisson_log_glm([1, -1, 0], x, 0.3, [0.3, 2.0])["rc"] == 2
    assert po.normal_id_glm([1.0, 2, 3], x, 0.3, [0.3, 2.0], 0.0)["rc"] == 2
    assert po.normal_id_glm([1.0, np.nan, 3], x, 0.3, [0.3, 2.0], 1.0)["rc"] == 2
    assert po.neg_binomial_2_log_glm([1, 1, 0], x, 0.3, [0.3, np.nan], 2.0)["rc"] == 2
    assert po.neg_binomial_2_log_glm([1, 1, 0], x, 0.3, [0.3, 2.0], -1.0)["rc"] == 2
    assert po.ordered_logistic_glm([1, 2, 3], x, [0.3, 2.0], [0.5, 0.2])["rc"] == 2
    assert po.ordered_logistic_glm([1, 2, 4], x, [0.3, 2.0], [0.2, 0.5])["rc"] == 2
    assert po.categorical_logit_glm([1, 4, 1], x, [0.1, 0.2, 0.3],
                                    np.ones((2, 3)))["rc"] == 2


def test_zero_sizes():
    x0 = np.zeros((0, 2))
    assert po.bernoulli_logit_glm([], x0, 0.3, [0.1, 0.2])["logp"] == 0.0
    assert po.poisson_log_glm([], x0, 0.3, [0.1, 0.2])["logp"] == 0.0
    xk0 = np.zeros((4, 0))
    r = po.poisson_log_glm([1, 2, 0, 1], xk0, 0.2, [])
    assert np.isfinite(r["logp"]) and r["d_beta"].size == 0


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built here")
@pytest.mark.parametrize("fam", ["bernoulli", "poisson", "normal", "neg_binomial",
                                 "ordered", "categorical", "binomial"])
def test_oracle_matches_reference_live(fam):
    for seed, (N, K) in enumerate([(11, 3), (200, 17), (1000, 40)]):
        d = make_inputs(fam, N, K, seed=seed, C=5)
        for propto in (0, po.PROPTO):
            fl = po.ALL_PARAMS | propto
            a = run_oracle(fam, d, fl, "oracle")
            b = run_oracle(fam, d, fl, "ref")
            assert a["rc"] == b["rc"] == 0
            assert_logp(a["logp"], b["logp"])
            for k in b:
                if k.startswith("d_") and b[k] is not None and a[k] is not None:
                    assert_grad(a[k], b[k], k,
                                scale=float(np.max(np.abs(b["d_beta"]))) * 1e-2)


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built here")
def test_scalar_kernels_match_reference():
    lib, ref = po.oracle_lib(), po.ref_lib()
    xs = np.concatenate([np.linspace(0.01, 30, 400), np.linspace(30, 1e4, 50),
                         -np.linspace(0.3, 9.7, 48)])
    for x in xs:
        a, b = lib.oracle_digamma(x), ref.ref_digamma(x)
        assert abs(a - b) <= 1e-13 * max(1.0, abs(b)), (x, a, b)
        if x > 0:
            a, b = lib.oracle_lgamma(x), ref.ref_lgamma(x)
            assert abs(a - b) <= 1e-14 * max(1.0, abs(b))
    for x in np.linspace(-50, 50, 201):
        assert lib.oracle_log1p_exp(x) == ref.ref_log1p_exp(x)
        if x < 0:
            assert lib.oracle_log1m_exp(x) == ref.ref_log1m_exp(x)


@pytest.mark.skipif(not po.ref_available(), reason="oracle/_ref not built here")
def test_binomial_scalar_kernels_match_reference():
    """log_inv_logit / log1m_inv_logit bit for bit; binomial_coefficient_log over
    every branch of lbeta (both small, one large, both large) and the symmetric
    k > n/2 fold."""
    lib, ref = po.oracle_lib(), po.ref_lib()
    for u in np.concatenate([np.linspace(-60, 60, 241), [0.0, -0.0, 1e-300, -745.0, 710.0]]):
        assert lib.oracle_log_inv_logit(u) == ref.ref_log_inv_logit(u)
        assert lib.oracle_log1m_inv_logit(u) == ref.ref_log1m_inv_logit(u)
    for n in [0, 1, 2, 8, 9, 10, 11, 18, 19, 20, 25, 100, 1000, 123456, 2_000_000_000]:
        ks = sorted({0, 1, 2, n // 3, n // 2, n // 2 + 1, max(n - 9, 0), max(n - 1, 0), n})
        for k in ks:
            if k > n:
                continue
            a = lib.oracle_binomial_coefficient_log(float(n), float(k))
            b = ref.ref_binomial_coefficient_log(float(n), float(k))
            assert abs(a - b) <= 1e-14 * max(1.0, abs(b)), (n, k, a, b)


def test_binomial_known_answers_and_errors():
    """Exact small cases: log C(N, n) + n log p + (N - n) log(1 - p), and the
    reference's error kinds (binomial_logit_glm_lpmf.hpp L88-99)."""
    from math import comb, log
    x = np.array([[0.5, -1.0], [0.25, 2.0], [-1.5, 0.75]])
    beta, alpha = np.array([0.3, -0.2]), 0.1
    n, trials = np.array([0, 3, 7]), np.array([4, 3, 12])
    th = x @ beta + alpha
    p = 1 / (1 + np.exp(-th))
    want = sum(log(comb(int(N), int(k))) + k * log(pi) + (N - k) * log(1 - pi)
               for k, N, pi in zip(n, trials, p))
    r = po.binomial_logit_glm(n, trials, x, alpha, beta)
    assert r["rc"] == 0 and r["logp"] == pytest.approx(want, rel=1e-13)
    np.testing.assert_allclose(r["d_beta"], x.T @ (n - trials * p), rtol=1e-12)
    assert r["d_alpha"][0] == pytest.approx(float(np.sum(n - trials * p)), rel=1e-12)
    # propto drops the binomial coefficient only
    r2 = po.binomial_logit_glm(n, trials, x, alpha, beta, flags=po.PROPTO | po.VAR_BETA)
    coef = sum(log(comb(int(N), int(k))) for k, N in zip(n, trials))
    assert r2["logp"] == pytest.approx(want - coef, rel=1e-13)
    assert po.binomial_logit_glm([0, 5, 1], [4, 3, 12], x, alpha, beta)["rc"] == 2  # n > N
    assert po.binomial_logit_glm([0, -1, 1], [4, 3, 12], x, alpha, beta)["rc"] == 2
    assert po.binomial_logit_glm([0, 1], [4, 3, 12], x, alpha, beta)["rc"] == 1
    assert po.binomial_logit_glm([0, 1, 1], [4, 3], x, alpha, beta)["rc"] == 1
    assert po.binomial_logit_glm(n, trials, x, np.inf, beta)["rc"] == 2
    assert po.binomial_logit_glm(n, trials, np.zeros((3, 0)), alpha, [])["logp"] == 0.0
