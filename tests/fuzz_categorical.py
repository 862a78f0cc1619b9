"""Random-shape sweep of the categorical family against the oracle: the fused GLM (incl. d_x and
more than 64 classes), the device matrix product + reverse sweep, and the row-wise
categorical_logit_lpmf.  tests/test_fuzz_gpu.py runs a seeded sweep; a longer one on a GPU box:
    python tests/fuzz_categorical.py [n_cases] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import math_b200 as mb  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.util import assert_grad, assert_logp  # noqa: E402


def one_case(rng):
    N = int(rng.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 33, 63, 129, 255, 257, 1000, 4099]))
    K = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 31, 32, 33, 64, 100, 257, 300]))
    C = int(rng.choice([2, 3, 7, 8, 9, 16, 17, 31, 32, 33, 41, 48, 56, 63, 64, 65, 72, 129]))
    x = np.asfortranarray(rng.standard_normal((N, K)))
    beta = np.asfortranarray(rng.standard_normal((K, C)) / np.sqrt(K))
    alpha = rng.standard_normal(C) * 0.3
    y = rng.integers(1, C + 1, N).astype(np.int32)
    tag = f"categorical N={N} K={K} C={C}"
    try:
        x_d, y_d = mb.to_matrix_cuda(x), mb.to_matrix_cuda(y)
        o = po.categorical_logit_glm(y, x, alpha, beta,
                                     flags=po.VAR_X | po.VAR_ALPHA | po.VAR_BETA)
        r = mb.categorical_logit_glm_lpmf(y_d, x_d, alpha, beta, var=("x", "alpha", "beta"))
        assert_logp(r.logp, o["logp"])
        sc = max(np.abs(o["d_beta"]).max(), 1e-3)
        assert_grad(r.d_alpha, o["d_alpha"], "d_alpha", scale=sc)
        assert_grad(r.d_beta, o["d_beta"], "d_beta")
        assert_grad(r.d_x.to_host(), o["d_x"], "d_x")
        lin_d = mb.lpmf.multiply_matrix(x_d, beta, alpha)
        lin = x @ beta + alpha
        assert_grad(mb.from_matrix_cuda(lin_d).ravel(), lin.ravel(), "lin",
                    scale=(np.abs(x) @ np.abs(beta)).max())
        u = mb.lpmf.categorical_logit_lpmf(y_d, lin_d)
        assert_logp(u.logp, o["logp"])
        g, cs = mb.lpmf.multiply_matrix_adjoint(x_d, u.d_theta)
        assert_grad(g, o["d_beta"], "x^T T")
        assert_grad(cs, o["d_alpha"], "colsum T", scale=sc)
    except Exception as e:  # noqa: BLE001
        return tag, f"{type(e).__name__}: {str(e)[:200]}"
    return tag, None


def run(n_cases=120, seed=2024):
    """Returns the failing cases as (tag, error) pairs (empty = all green)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_cases):
        tag, err = one_case(rng)
        if err:
            out.append((tag, err))
    return out


if __name__ == "__main__":
    mb.runtime.set_device(0)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    bad = run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
    for tag, err in bad:
        print(f"FAIL {tag}: {err}", flush=True)
    print(f"fuzz: {n - len(bad)}/{n} cases ok")
    sys.exit(1 if bad else 0)
