#!/usr/bin/env python
"""Generate tests/golden/glm_golden.json from the UNMODIFIED reference.

Runs only in the build container (needs oracle/_ref/libstan_ref.so, i.e.
/root/reference compiled by oracle/Makefile).  Inputs are the fixed inputs of the
reference's own tests (test/unit/math/rev/prob/*_glm_*_test.cpp) plus seeded
random cases in the shapes of test/unit/math/opencl/rev/*_glm_*_test.cpp
(small_simple 3x2, big 153x71, C = 43), for both propto settings and for x as
data and as var.  Floats are written with repr() so they round-trip exactly.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from tests.util import make_inputs  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "glm_golden.json")


def tolist(v):
    if v is None:
        return None
    a = np.asarray(v)
    return a.ravel(order="F").tolist() if a.ndim else a.item()


def flags_for(fam, propto, xvar):
    f = po.ALL_PARAMS | (po.PROPTO if propto else 0)
    if xvar:
        f |= po.VAR_X | (po.VAR_Y if fam == "normal" else 0)
    return f


def run(fam, d, flags):
    if fam == "bernoulli":
        return po.bernoulli_logit_glm(d["y"], d["x"], d["alpha"], d["beta"], flags, "ref")
    if fam == "poisson":
        return po.poisson_log_glm(d["y"], d["x"], d["alpha"], d["beta"], flags, "ref")
    if fam == "normal":
        return po.normal_id_glm(d["y"], d["x"], d["alpha"], d["beta"], d["sigma"], flags, "ref")
    if fam == "neg_binomial":
        return po.neg_binomial_2_log_glm(d["y"], d["x"], d["alpha"], d["beta"], d["phi"], flags, "ref")
    if fam == "binomial":
        return po.binomial_logit_glm(d["y"], d["trials"], d["x"], d["alpha"], d["beta"],
                                     flags, "ref")
    if fam == "ordered":
        return po.ordered_logistic_glm(d["y"], d["x"], d["beta"], d["cuts"], flags, "ref")
    return po.categorical_logit_glm(d["y"], d["x"], d["alpha"], d["beta"], flags, "ref")


def fixed_cases():
    x = np.array([[-12, 46], [-42, 24], [25, 27]], float)
    xo = np.array([[1, 2], [3, 4], [5, 6], [7, 8], [9, 0]], float)
    xc = np.array([[-12, 46], [-42, 24], [25, 27], [-14, -11], [5, 18]], float)
    return [
        ("bernoulli", "ref_test_fixed", dict(y=[1, 0, 1], x=x, alpha=0.3, beta=[0.3, 2.0])),
        ("normal", "ref_test_fixed", dict(y=[14.0, 32.0, 21.0], x=x, alpha=0.3,
                                          beta=[0.3, 2.0], sigma=10.0)),
        ("poisson", "ref_test_fixed", dict(y=[14, 2, 5], x=x / 100, alpha=0.3, beta=[0.3, 2.0])),
        ("neg_binomial", "ref_test_fixed", dict(y=[14, 2, 5], x=x / 100, alpha=0.3,
                                                beta=[0.3, 2.0], phi=2.0)),
        ("ordered", "ref_test_fixed", dict(y=[1, 1, 2, 4, 4], x=xo, beta=[1.1, 0.4],
                                           cuts=[0.9, 1.1, 7.0])),
        ("categorical", "ref_test_fixed", dict(y=[1, 3, 1, 2, 2], x=xc, alpha=[0.5, -2.0, 4.0],
                                               beta=np.array([[0.3, 2, 0.4], [-0.1, -1.3, 1]]))),
    ]


def binomial_fixed_cases():
    """Inputs of test/unit/math/opencl/rev/binomial_logit_glm_lpmf_test.cpp
    (small_simple: n = {0, 1, 0}, N = {1, 2, 3}) and scalar broadcasts."""
    x = np.array([[-12, 46], [-42, 24], [25, 27]], float) / 100
    return [
        ("binomial", "ref_test_fixed", dict(y=[0, 1, 0], trials=[1, 2, 3], x=x, alpha=0.3,
                                            beta=[0.3, 2.0])),
        ("binomial", "broadcast_trials", dict(y=[0, 11, 30], trials=[30], x=x, alpha=0.3,
                                              beta=[0.3, 2.0])),
        ("binomial", "broadcast_both", dict(y=[7], trials=[19], x=x, alpha=-0.2,
                                            beta=[0.3, 2.0])),
        ("binomial", "saturated_tails", dict(y=[3, 0, 500], trials=[3, 8, 1000],
                                             x=x * 100, alpha=0.3, beta=[0.3, 2.0])),
    ]


def random_cases(families=("bernoulli", "poisson", "normal", "neg_binomial", "ordered",
                           "categorical")):
    cases = []
    for fam in families:
        big = (153, 71, 43, "big") if fam in ("bernoulli", "neg_binomial", "binomial",
                                               "categorical") else (64, 17, 9, "big")
        for (N, K, C, tag) in ((3, 2, 3, "small_simple"), big,
                               (40, 5, 4, "mid"), (17, 1, 2, "one_attribute")):
            for vec in ((False, True) if fam in ("bernoulli", "poisson", "normal",
                                                 "neg_binomial", "binomial") else (False,)):
                if tag == "big" and vec:
                    continue
                d = make_inputs(fam, N, K, seed=1000 + N * 7 + K, C=C, vec_alpha=vec,
                                vec_aux=vec)
                cases.append((fam, tag + ("_vec" if vec else ""), d))
        # broadcast_y: one scalar response for every instance
        d = make_inputs(fam, 9, 3, seed=77, C=4)
        d["y"] = [float(d["y"][0])] if fam == "normal" else [int(d["y"][0])]
        if fam == "binomial":
            d["trials"] = np.maximum(d["trials"], d["y"][0])
        cases.append((fam, "broadcast_y", d))
    return cases


def main():
    if not po.ref_available():
        raise SystemExit("oracle/_ref/libstan_ref.so missing: run `make -C oracle ref` "
                         "in the container that has /root/reference")
    if len(sys.argv) > 1 and sys.argv[1] == "binomial":
        # the seventh GLM has its own fixture: python tests/golden/make_golden.py binomial
        write(binomial_fixed_cases() + random_cases(("binomial",)),
              os.path.join(os.path.dirname(OUT), "binomial_golden.json"))
    else:
        write(fixed_cases() + random_cases(), OUT)


def write(all_cases, path):
    out = []
    for fam, tag, d in all_cases:
        variants = []
        for propto in (False, True):
            for xvar in (False, True):
                if xvar and np.asarray(d["x"]).size > 2000:
                    continue  # keep the fixture small: no N x K gradient for "big"
                fl = flags_for(fam, propto, xvar)
                r = run(fam, d, fl)
                assert r["rc"] == 0, (fam, tag, r["rc"])
                variants.append({"propto": propto, "x_var": xvar, "flags": fl,
                                 "expect": {k: tolist(v) for k, v in r.items()
                                            if k != "rc" and v is not None}})
        out.append({"family": fam, "case": tag,
                    "shape": list(np.asarray(d["x"]).shape),
                    "beta_shape": list(np.asarray(d["beta"]).shape),
                    "inputs": {k: tolist(v) for k, v in d.items()},
                    "variants": variants})
    with open(path, "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py",
                   "reference": "stan-dev/math 5.0.x at /root/reference, prim GLMs under "
                                "reverse-mode var (oracle/ref_driver.cpp)",
                   "cases": out}, f)
    print(f"wrote {len(out)} cases x variants to {path} ({os.path.getsize(path)/1e6:.2f} MB)")


if __name__ == "__main__":
    main()
