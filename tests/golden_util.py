"""Loader for tests/golden/glm_golden.json (reference outputs, see make_golden.py)."""
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glm_golden.json")


def load():
    with open(PATH) as f:
        return json.load(f)["cases"]


def inputs_of(case):
    N, K = case["shape"]
    d = {}
    for k, v in case["inputs"].items():
        a = np.asarray(v, dtype=np.float64)
        if k == "x":
            a = a.reshape((N, K), order="F")
        elif k == "beta" and len(case["beta_shape"]) == 2:
            a = a.reshape(tuple(case["beta_shape"]), order="F")
        elif k == "y" and case["family"] != "normal":
            a = a.astype(np.int32)
        d[k] = a
    return d


def ids(cases):
    return [f'{c["family"]}-{c["case"]}' for c in cases]
