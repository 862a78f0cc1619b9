"""Shared helpers of the parity tests: tolerances and seeded inputs.

Tolerances are BASELINE.json's: relative 1e-10 on the log density, relative 1e-9
on gradients, plus an absolute floor near zero (sums that cancel: the floor is
1e-12 times the largest magnitude in the compared vector, never below 1e-13)."""
import numpy as np

RTOL_LOGP = 1e-10
RTOL_GRAD = 1e-9


def assert_logp(got, want, what="logp"):
    tol = RTOL_LOGP * max(abs(want), abs(got)) + 1e-13
    assert abs(got - want) <= tol, f"{what}: got {got!r} want {want!r} diff {got-want:.3e}"


def assert_grad(got, want, what="grad", scale=None):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} vs {want.shape}"
    if want.size == 0:
        return
    big = max(float(np.max(np.abs(want))), float(scale or 0.0))
    floor = max(1e-12 * big, 1e-13)
    err = np.abs(got - want)
    tol = RTOL_GRAD * np.maximum(np.abs(got), np.abs(want)) + floor
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{want.size} entries differ; "
                           f"worst {float(err.max()):.3e} at "
                           f"{np.unravel_index(int(np.argmax(err - tol)), want.shape)}")


def make_inputs(family, N, K, seed=12345, C=None, vec_alpha=False, vec_aux=False):
    """Seeded synthetic inputs of SURVEY.md 8(d): x ~ N(0,1), beta ~ N(0,1)/sqrt(K)."""
    rng = np.random.default_rng(seed)
    x = np.asfortranarray(rng.standard_normal((N, K)))
    beta = rng.standard_normal(K) / np.sqrt(max(K, 1))
    alpha = 0.1 + 0.3 * rng.standard_normal(N) if vec_alpha else 0.1
    theta = x @ beta + alpha
    d = dict(x=x, beta=beta, alpha=alpha)
    if family == "bernoulli":
        d["y"] = (rng.random(N) < 1 / (1 + np.exp(-theta))).astype(np.int32)
    elif family == "binomial":
        # trials span both sides of binomial_coefficient_log's N + 1 < 10 switch
        # and its lbeta branches (small/small, small/large, large/large)
        trials = np.where(rng.random(N) < 0.5, rng.integers(0, 12, N),
                          rng.integers(12, 400, N)).astype(np.int32)
        d["trials"] = trials
        d["y"] = rng.binomial(trials, 1 / (1 + np.exp(-theta))).astype(np.int32)
    elif family in ("poisson", "neg_binomial"):
        d["y"] = rng.integers(0, 5, N).astype(np.int32)
        if family == "neg_binomial":
            d["phi"] = 0.5 + 4 * rng.random(N) if vec_aux else 2.5
    elif family == "normal":
        sigma = 0.5 + 2 * rng.random(N) if vec_aux else 1.3
        d["sigma"] = sigma
        d["y"] = theta + sigma * rng.standard_normal(N)
    elif family == "ordered":
        C = C or 9
        d["cuts"] = np.sort(rng.uniform(-2, 2, C - 1)) + np.arange(C - 1) * 1e-3
        d["y"] = rng.integers(1, C + 1, N).astype(np.int32)
        d.pop("alpha")
    elif family == "categorical":
        C = C or 4
        d["beta"] = np.asfortranarray(rng.standard_normal((K, C)) / np.sqrt(max(K, 1)))
        d["alpha"] = 0.1 * rng.standard_normal(C)
        d["y"] = rng.integers(1, C + 1, N).astype(np.int32)
    return d
