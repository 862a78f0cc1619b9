"""Loader for tests/golden/glm_golden.json (reference outputs, see make_golden.py)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "glm_golden.json")
# the seventh GLM has its own fixture, same generator
PATH_BINOMIAL = os.path.join(HERE, "golden", "binomial_golden.json")


def load():
    cases = []
    for p in (PATH, PATH_BINOMIAL):
        with open(p) as f:
            cases += json.load(f)["cases"]
    return cases


def inputs_of(case):
    N, K = case["shape"]
    d = {}
    for k, v in case["inputs"].items():
        a = np.asarray(v, dtype=np.float64)
        if k == "x":
            a = a.reshape((N, K), order="F")
        elif k == "beta" and len(case["beta_shape"]) == 2:
            a = a.reshape(tuple(case["beta_shape"]), order="F")
        elif k in ("y", "trials") and case["family"] != "normal":
            a = a.astype(np.int32)
        d[k] = a
    return d


def ids(cases):
    return [f'{c["family"]}-{c["case"]}' for c in cases]
