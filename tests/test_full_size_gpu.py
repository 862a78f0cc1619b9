"""Parity at BASELINE.json's FULL sizes, through properties that do not need a
CPU pass over the whole matrix (SURVEY.md 8(d) "Parity at scale"):

  additivity   every family is a sum over independent rows, so the value and the
               reduced partials over all N rows must equal the sum of four
               separate evaluations over contiguous row blocks (a checksum of
               checksums; the two sides use different tile -> CTA schedules)
  sample       the CPU oracle on a contiguous row block downloaded from the
               device against the GPU on a view of the same block
  d_x          (config 4) a downloaded block of the N x K gradient written by the
               full-size run against the oracle's d_x for those rows
  determinism  two full-size runs return the same bits

Inputs are the counter-based synthetic data of bench.py, generated on the device.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import assert_grad, assert_logp

pytestmark = pytest.mark.gpu

SAMPLE = 4096


def _synth(mb, N, K, seed=12345):
    x = mb.MatrixCuda(N, K)
    x.fill_synthetic(seed, kind=0)
    return x


def _ints(mb, N, lo, hi, seed=777):
    y = mb.MatrixCuda(N, 1, np.int32)
    y.fill_synthetic(seed, kind=1, lo=lo, hi=hi)
    return y


def _blocks(N, parts=4):
    # block edges on multiples of 32 rows (a view must keep x 16-byte aligned)
    edges = [0] + [(N * i // parts) // 32 * 32 for i in range(1, parts)] + [N]
    return list(zip(edges[:-1], edges[1:]))


def _close(a, b, what, rel=1e-11):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), 1e-300)
    assert np.all(np.abs(a - b) <= rel * scale), \
        f"{what}: max |diff| {float(np.max(np.abs(a - b))):.3e} vs scale {scale:.3e}"


def _beta(K, seed=12345):
    return np.random.default_rng(seed).standard_normal(K) / np.sqrt(K)


def test_config2_bernoulli_N1e7_K256(gpu):
    mb = gpu
    N, K = 10_000_000, 256
    x, y, beta = _synth(mb, N, K), _ints(mb, N, 0, 1), _beta(K)
    full = mb.bernoulli_logit_glm_lpmf(y, x, 0.1, beta)
    again = mb.bernoulli_logit_glm_lpmf(y, x, 0.1, beta)
    assert full.logp == again.logp and np.array_equal(full.d_beta, again.d_beta)
    lp, da, db = 0.0, 0.0, np.zeros(K)
    for r0, r1 in _blocks(N):
        r = mb.bernoulli_logit_glm_lpmf(y.view_rows(r0, r1 - r0), x.view_rows(r0, r1 - r0),
                                        0.1, beta)
        lp, da, db = lp + r.logp, da + r.d_alpha, db + r.d_beta
    _close(full.logp, lp, "logp additivity")
    _close(full.d_alpha, da, "d_alpha additivity")
    _close(full.d_beta, db, "d_beta additivity")
    r0 = 7_000_000 - 64
    xs, ys = x.rows_to_host(r0, SAMPLE), y.rows_to_host(r0, SAMPLE).ravel()
    g = mb.bernoulli_logit_glm_lpmf(y.view_rows(r0, SAMPLE), x.view_rows(r0, SAMPLE), 0.1, beta)
    o = po.bernoulli_logit_glm(ys, xs, 0.1, beta)
    assert_logp(g.logp, o["logp"])
    assert_grad(g.d_beta, o["d_beta"], "d_beta (sample)")


def test_config3_poisson_N1e8_K64(gpu):
    mb = gpu
    N, K = 100_000_000, 64
    x, y, beta = _synth(mb, N, K), _ints(mb, N, 0, 4), _beta(K)
    full = mb.poisson_log_glm_lpmf(y, x, 0.1, beta)
    lp, db = 0.0, np.zeros(K)
    for r0, r1 in _blocks(N):
        r = mb.poisson_log_glm_lpmf(y.view_rows(r0, r1 - r0), x.view_rows(r0, r1 - r0), 0.1, beta)
        lp, db = lp + r.logp, db + r.d_beta
    _close(full.logp, lp, "logp additivity")
    _close(full.d_beta, db, "d_beta additivity")
    r0 = 99_000_000
    xs, ys = x.rows_to_host(r0, SAMPLE), y.rows_to_host(r0, SAMPLE).ravel()
    g = mb.poisson_log_glm_lpmf(y.view_rows(r0, SAMPLE), x.view_rows(r0, SAMPLE), 0.1, beta)
    o = po.poisson_log_glm(ys, xs, 0.1, beta)
    assert_logp(g.logp, o["logp"])
    assert_grad(g.d_beta, o["d_beta"], "d_beta (sample)")


def test_config4_neg_binomial_N1e7_K128_x_var(gpu):
    mb = gpu
    N, K = 10_000_000, 128
    x, y, beta = _synth(mb, N, K), _ints(mb, N, 0, 4), _beta(K)
    var = ("x", "alpha", "beta", "phi")
    full = mb.neg_binomial_2_log_glm_lpmf(y, x, 0.1, beta, 2.5, var=var)
    lp, db, dphi = 0.0, np.zeros(K), 0.0
    for r0, r1 in _blocks(N):
        r = mb.neg_binomial_2_log_glm_lpmf(y.view_rows(r0, r1 - r0), x.view_rows(r0, r1 - r0),
                                           0.1, beta, 2.5, var=("alpha", "beta", "phi"))
        lp, db, dphi = lp + r.logp, db + r.d_beta, dphi + r.d_aux
    _close(full.logp, lp, "logp additivity")
    _close(full.d_beta, db, "d_beta additivity")
    _close(full.d_aux, dphi, "d_phi additivity", rel=1e-10)
    for r0 in (0, 5_000_000 + 32, N - SAMPLE):  # first, middle and last (ragged) tiles
        xs, ys = x.rows_to_host(r0, SAMPLE), y.rows_to_host(r0, SAMPLE).ravel()
        o = po.neg_binomial_2_log_glm(ys, xs, 0.1, beta, 2.5,
                                      flags=po.VAR_X | po.VAR_ALPHA | po.VAR_BETA | po.VAR_AUX)
        assert_grad(full.d_x.rows_to_host(r0, SAMPLE), o["d_x"], f"d_x rows {r0}..")


def test_config5b_ordered_N1e7_K64(gpu):
    mb = gpu
    N, K = 10_000_000, 64
    x, y, beta = _synth(mb, N, K), _ints(mb, N, 1, 9), _beta(K)
    cuts = np.linspace(-2, 2, 8)
    full = mb.ordered_logistic_glm_lpmf(y, x, beta, cuts)
    lp, db, dc = 0.0, np.zeros(K), np.zeros(8)
    for r0, r1 in _blocks(N):
        r = mb.ordered_logistic_glm_lpmf(y.view_rows(r0, r1 - r0), x.view_rows(r0, r1 - r0),
                                         beta, cuts)
        lp, db, dc = lp + r.logp, db + r.d_beta, dc + r.d_aux
    _close(full.logp, lp, "logp additivity")
    _close(full.d_beta, db, "d_beta additivity")
    _close(full.d_aux, dc, "d_cuts additivity")
    r0 = 3_333_312
    xs, ys = x.rows_to_host(r0, SAMPLE), y.rows_to_host(r0, SAMPLE).ravel()
    g = mb.ordered_logistic_glm_lpmf(y.view_rows(r0, SAMPLE), x.view_rows(r0, SAMPLE), beta, cuts)
    o = po.ordered_logistic_glm(ys, xs, beta, cuts)
    assert_logp(g.logp, o["logp"])
    assert_grad(g.d_beta, o["d_beta"], "d_beta (sample)")
    assert_grad(g.d_aux, o["d_cuts"], "d_cuts (sample)", scale=np.abs(o["d_beta"]).max() * 1e-2)


def test_config5a_categorical_N2e6_K512_C32(gpu):
    mb = gpu
    N, K, C = 2_000_000, 512, 32
    rng = np.random.default_rng(12345)
    x, y = _synth(mb, N, K), _ints(mb, N, 1, C)
    beta = np.asfortranarray(rng.standard_normal((K, C)) / np.sqrt(K))
    alpha = 0.1 * rng.standard_normal(C)
    full = mb.categorical_logit_glm_lpmf(y, x, alpha, beta)
    lp, da, db = 0.0, np.zeros(C), np.zeros((K, C))
    for r0, r1 in _blocks(N):
        r = mb.categorical_logit_glm_lpmf(y.view_rows(r0, r1 - r0), x.view_rows(r0, r1 - r0),
                                          alpha, beta)
        lp, da, db = lp + r.logp, da + r.d_alpha, db + r.d_beta
    _close(full.logp, lp, "logp additivity")
    _close(full.d_alpha, da, "d_alpha additivity")
    _close(full.d_beta, db, "d_beta additivity")
    r0, n = 1_000_000 - 32, 2048
    xs, ys = x.rows_to_host(r0, n), y.rows_to_host(r0, n).ravel()
    g = mb.categorical_logit_glm_lpmf(y.view_rows(r0, n), x.view_rows(r0, n), alpha, beta)
    o = po.categorical_logit_glm(ys, xs, alpha, beta)
    assert_logp(g.logp, o["logp"])
    assert_grad(g.d_alpha, o["d_alpha"], "d_alpha (sample)", scale=np.abs(o["d_beta"]).max())
    assert_grad(g.d_beta, o["d_beta"], "d_beta (sample)")
