"""Random-shape sweep of the memory-bound GLM families against the oracle: random N, K (incl.
K > 256: column chunks), scalar / vector alpha and sigma / phi, x as data or autodiff (the full
d_x, or the factor d of d_x = d beta^T applied by rank1_update), propto.
tests/test_fuzz_gpu.py runs a seeded sweep; a longer one on a GPU box:
    python tests/fuzz_glm.py [n_cases] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import math_b200 as mb  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tests.util import assert_grad, assert_logp, make_inputs  # noqa: E402

FLAG = {"x": po.VAR_X, "alpha": po.VAR_ALPHA, "beta": po.VAR_BETA, "aux": po.VAR_AUX}
AUX_NAME = {"normal": "sigma", "neg_binomial": "phi", "ordered": "cuts"}
AUX_KEY = {"normal": "d_sigma", "neg_binomial": "d_phi", "ordered": "d_cuts"}


def host(v):
    return v.to_host().ravel() if hasattr(v, "to_host") else np.atleast_1d(v)


def one_case(rng):
    """One random case; returns (tag, error text or None)."""
    fam = str(rng.choice(["bernoulli", "poisson", "normal", "neg_binomial", "ordered"]))
    N = int(rng.choice([1, 2, 31, 32, 33, 127, 129, 1000, 4099, 20011]))
    K = int(rng.choice([1, 2, 3, 31, 32, 33, 64, 100, 128, 255, 256, 257, 300, 520]))
    vec = bool(rng.integers(0, 2)) and fam != "ordered"
    xvar = int(rng.integers(0, 3))  # 0: data, 1: full d_x, 2: factored d_x
    propto = bool(rng.integers(0, 2))
    d = make_inputs(fam, N, K, seed=int(rng.integers(1, 1 << 30)), C=int(rng.integers(2, 12)),
                    vec_alpha=vec, vec_aux=vec)
    names = ["beta"] + (["x"] if xvar else []) + ([] if fam == "ordered" else ["alpha"]) \
        + (["aux"] if fam in AUX_NAME else [])
    oflags = po.PROPTO if propto else 0
    for n in names:
        oflags |= FLAG[n]
    tag = f"{fam} N={N} K={K} vec={vec} xvar={xvar} propto={propto}"
    var = [AUX_NAME[fam] if n == "aux" else ("x_factored" if n == "x" and xvar == 2 else n)
           for n in names]
    try:
        x = mb.to_matrix_cuda(d["x"])
        y = mb.to_matrix_cuda(d["y"])
        al = mb.to_matrix_cuda(d["alpha"]) if vec else d.get("alpha")
        if fam == "bernoulli":
            r = mb.bernoulli_logit_glm_lpmf(y, x, al, d["beta"], propto=propto, var=var)
            o = po.bernoulli_logit_glm(d["y"], d["x"], d["alpha"], d["beta"], oflags)
        elif fam == "poisson":
            r = mb.poisson_log_glm_lpmf(y, x, al, d["beta"], propto=propto, var=var)
            o = po.poisson_log_glm(d["y"], d["x"], d["alpha"], d["beta"], oflags)
        elif fam == "normal":
            sg = mb.to_matrix_cuda(d["sigma"]) if vec else d["sigma"]
            r = mb.normal_id_glm_lpdf(y, x, al, d["beta"], sg, propto=propto, var=var)
            o = po.normal_id_glm(d["y"], d["x"], d["alpha"], d["beta"], d["sigma"], oflags)
        elif fam == "neg_binomial":
            ph = mb.to_matrix_cuda(d["phi"]) if vec else d["phi"]
            r = mb.neg_binomial_2_log_glm_lpmf(y, x, al, d["beta"], ph, propto=propto, var=var)
            o = po.neg_binomial_2_log_glm(d["y"], d["x"], d["alpha"], d["beta"], d["phi"], oflags)
        else:
            r = mb.ordered_logistic_glm_lpmf(y, x, d["beta"], d["cuts"], propto=propto, var=var)
            o = po.ordered_logistic_glm(d["y"], d["x"], d["beta"], d["cuts"], oflags)
        assert o["rc"] == 0, o["rc"]
        assert_logp(r.logp, o["logp"])
        sc = max(np.abs(o["d_beta"]).max(), 1e-6)
        assert_grad(r.d_beta, o["d_beta"], "d_beta")
        if "alpha" in names:
            if vec:
                assert_grad(host(r.d_alpha), o["d_alpha"], "d_alpha")
            else:
                assert_grad(host(r.d_alpha), np.atleast_1d(o["d_alpha"])[:1], "d_alpha", scale=sc)
        if "aux" in names:
            key = AUX_KEY[fam]
            want = np.atleast_1d(o[key])
            got = host(r.d_aux)
            if fam == "ordered" or vec:
                assert_grad(got, want, key, scale=sc * 1e-2 if fam == "ordered" else None)
            else:
                assert_grad(got, want[:1], key, scale=max(N * 1e-2, sc))
        if xvar == 1:
            assert_grad(r.d_x.to_host(), o["d_x"], "d_x")
        elif xvar == 2:
            # the reverse sweep from the factor: x.adj = a * d beta^T into a lazily zero
            # adjoint (pure store), then accumulated once more (read-modify-write)
            a = float(rng.choice([1.0, -0.75]))
            adj = mb.MatrixCuda(N, K)
            adj.zero_lazy()
            adj.rank1_update(a, r.d_x, d["beta"])
            assert_grad(adj.to_host(), a * o["d_x"], "x.adj (store)")
            adj.rank1_update(a, r.d_x, d["beta"])
            assert_grad(adj.to_host(), 2 * a * o["d_x"], "x.adj (accumulate)")
    except Exception as e:  # noqa: BLE001
        return tag, f"{type(e).__name__}: {str(e)[:200]}"
    return tag, None


def run(n_cases=200, seed=7):
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    bad = run(n, int(sys.argv[2]) if len(sys.argv) > 2 else 7)
    for tag, err in bad:
        print(f"FAIL {tag}: {err}", flush=True)
    print(f"fuzz: {n - len(bad)}/{n} cases ok")
    sys.exit(1 if bad else 0)
