#!/usr/bin/env python
"""Generate tests/golden/categorical_lpmf_golden.json from the UNMODIFIED reference.

Runs only in the build container (needs oracle/_ref/libstan_ref.so): every row of
an N x C matrix of log odds goes through the reference's own
categorical_logit_lpmf(int, column vector) with `var` log odds and grad()
(oracle/ref_driver.cpp: ref_categorical_logit_lpmf).  Inputs: the log odds the
reference's test uses (test/unit/math/prim/prob/categorical_logit_test.cpp:
theta = (-1, 2, -10)), rows with saturated / tied entries, and seeded random
matrices for C = 1, 2, 8, 9, 32, 43.  Floats are written with repr().

    python tests/golden/make_golden_categorical_lpmf.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                   "categorical_lpmf_golden.json")


def cases():
    out = [("ref_test_theta", [1, 2, 3], np.tile([-1.0, 2.0, -10.0], (3, 1))),
           ("tails_and_ties", [1, 3, 1, 2, 2],
            np.array([[0.5, -2, 4], [1.3, 0.2, -0.7], [-30, 2, 55], [0, 0, 0],
                      [700, -700, 1e-3]], float)),
           ("broadcast_y", [2], np.array([[0.1, 0.2, 0.3], [3.0, -1.0, 0.5]]))]
    for C in (1, 2, 8, 9, 32, 43):
        rng = np.random.default_rng(500 + C)
        N = 37 if C < 32 else 19
        out.append((f"random_C{C}", rng.integers(1, C + 1, N).tolist(),
                    rng.standard_normal((N, C)) * 4.0))
    return out


def main():
    if not po.ref_available():
        raise SystemExit("oracle/_ref/libstan_ref.so missing: run `make -C oracle ref`")
    doc = []
    for tag, y, lin in cases():
        for propto in (False, True):
            r = po.categorical_logit_lpmf(y, lin, po.VAR_ALPHA | (po.PROPTO if propto else 0),
                                          impl="ref")
            assert r["rc"] == 0, (tag, r["rc"])
            doc.append({"case": tag, "propto": propto, "shape": list(lin.shape), "y": list(y),
                        "lin": lin.ravel(order="F").tolist(), "logp": r["logp"],
                        "d_lin": r["d_lin"].ravel(order="F").tolist()})
    with open(OUT, "w") as f:
        json.dump({"generator": "tests/golden/make_golden_categorical_lpmf.py",
                   "reference": "stan-dev/math prim/prob/categorical_logit_lpmf.hpp, per row",
                   "cases": doc}, f)
    print(f"wrote {len(doc)} cases to {OUT}")


if __name__ == "__main__":
    main()
