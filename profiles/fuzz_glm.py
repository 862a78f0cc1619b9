"""Random-shape sweep of the memory-bound GLM families against the oracle: random N, K (incl.
K > 256: column chunks), scalar / vector alpha and sigma / phi, x as data or autodiff, propto.
Not a test (tests/ holds the fixed cases); run on a GPU box:
    python profiles/fuzz_glm.py [n_cases] [seed]"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import math_b200 as mb
from oracle import pyoracle as po
from tests.util import assert_grad, assert_logp, make_inputs

mb.runtime.set_device(0)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
FLAG = {"x": po.VAR_X, "alpha": po.VAR_ALPHA, "beta": po.VAR_BETA, "aux": po.VAR_AUX}
bad = 0


def host(v):
    return v.to_host().ravel() if hasattr(v, "to_host") else np.atleast_1d(v)


for case in range(n_cases):
    fam = str(rng.choice(["bernoulli", "poisson", "normal", "neg_binomial", "ordered"]))
    N = int(rng.choice([1, 2, 31, 32, 33, 127, 129, 1000, 4099, 20011]))
    K = int(rng.choice([1, 2, 3, 31, 32, 33, 64, 100, 128, 255, 256, 257, 300, 520]))
    vec = bool(rng.integers(0, 2)) and fam != "ordered"
    xvar = bool(rng.integers(0, 2))
    propto = bool(rng.integers(0, 2))
    d = make_inputs(fam, N, K, seed=int(rng.integers(1, 1 << 30)), C=int(rng.integers(2, 12)),
                    vec_alpha=vec, vec_aux=vec)
    names = ["beta"] + (["x"] if xvar else []) + ([] if fam == "ordered" else ["alpha"]) \
        + (["aux"] if fam in ("normal", "neg_binomial", "ordered") else [])
    oflags = (po.PROPTO if propto else 0)
    for n in names:
        oflags |= FLAG[n]
    tag = f"{fam} N={N} K={K} vec={vec} xvar={xvar} propto={propto}"
    try:
        x = mb.to_matrix_cuda(d["x"])
        y = mb.to_matrix_cuda(d["y"])
        al = mb.to_matrix_cuda(d["alpha"]) if vec else d.get("alpha")
        if fam == "bernoulli":
            var = [n for n in names]
            r = mb.bernoulli_logit_glm_lpmf(y, x, al, d["beta"], propto=propto, var=var)
            o = po.bernoulli_logit_glm(d["y"], d["x"], d["alpha"], d["beta"], oflags)
        elif fam == "poisson":
            r = mb.poisson_log_glm_lpmf(y, x, al, d["beta"], propto=propto, var=names)
            o = po.poisson_log_glm(d["y"], d["x"], d["alpha"], d["beta"], oflags)
        elif fam == "normal":
            sg = mb.to_matrix_cuda(d["sigma"]) if vec else d["sigma"]
            var = [("sigma" if n == "aux" else n) for n in names]
            r = mb.normal_id_glm_lpdf(y, x, al, d["beta"], sg, propto=propto, var=var)
            o = po.normal_id_glm(d["y"], d["x"], d["alpha"], d["beta"], d["sigma"], oflags)
        elif fam == "neg_binomial":
            ph = mb.to_matrix_cuda(d["phi"]) if vec else d["phi"]
            var = [("phi" if n == "aux" else n) for n in names]
            r = mb.neg_binomial_2_log_glm_lpmf(y, x, al, d["beta"], ph, propto=propto, var=var)
            o = po.neg_binomial_2_log_glm(d["y"], d["x"], d["alpha"], d["beta"], d["phi"], oflags)
        else:
            var = [("cuts" if n == "aux" else n) for n in names]
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
            key = {"normal": "d_sigma", "neg_binomial": "d_phi", "ordered": "d_cuts"}[fam]
            want = np.atleast_1d(o[key])
            got = host(r.d_aux)
            if fam == "ordered" or vec:
                assert_grad(got, want, key, scale=sc * 1e-2 if fam == "ordered" else None)
            else:
                assert_grad(got, want[:1], key, scale=max(N * 1e-2, sc))
        if xvar:
            assert_grad(r.d_x.to_host(), o["d_x"], "d_x")
    except Exception as e:  # noqa: BLE001
        bad += 1
        print(f"FAIL {tag}: {type(e).__name__}: {str(e)[:160]}", flush=True)
print(f"fuzz: {n_cases - bad}/{n_cases} cases ok")
sys.exit(1 if bad else 0)
