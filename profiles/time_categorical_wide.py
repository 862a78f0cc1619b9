"""Fused categorical_logit_glm_lpmf for 33-64 classes (NT = 5 ... 8 of cat_lin_tma_kernel):
wall-clock over synchronous C-ABI calls, alpha + beta var.  usage: [N K C]..."""
import json
import sys
import time

sys.path.insert(0, '/root/repo')
import numpy as np
import math_b200 as mb

mb.runtime.set_device(0)
SHAPES = [(1_000_000, 256, 64), (2_000_000, 512, 40), (1_000_000, 512, 56), (2_000_000, 256, 48)]
rng = np.random.default_rng(3)
for N, K, C in SHAPES:
    x = mb.MatrixCuda(N, K); x.fill_synthetic(12345, kind=0)
    y = mb.MatrixCuda(N, 1, np.int32); y.fill_synthetic(777, kind=1, lo=1, hi=C)
    beta = rng.standard_normal((K, C)) / np.sqrt(K)
    alpha = rng.standard_normal(C) * 0.1
    f = lambda: mb.categorical_logit_glm_lpmf(y, x, alpha, beta)
    r = None
    for _ in range(3):
        r = f()
    mb.runtime.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        r = f()
    mb.runtime.synchronize()
    t = (time.perf_counter() - t0) / 10
    print(json.dumps({"N": N, "K": K, "C": C, "ms": round(t * 1e3, 4),
                      "fp64_tflops": round(4.0 * N * K * C / t / 1e12, 2),
                      "logp_per_row": r.logp / N}))
    del x, y
