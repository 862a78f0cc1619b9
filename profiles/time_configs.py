#!/usr/bin/env python
"""Times every BASELINE.json config on ONE GPU through the public C-ABI call
(host parameters in, host results out, x resident) and prints one JSON line per
config with ms/eval and the fraction of the measured HBM peak (or FP64 rate for
the categorical GLM).  Not the bench contract -- a side table for DESIGN.md."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import math_b200 as mb  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps):
    fn()
    fn()
    mb.runtime.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    mb.runtime.synchronize()
    return (time.perf_counter() - t0) / reps


def synth(N, K, seed=12345):
    x = mb.MatrixCuda(N, K)
    x.fill_synthetic(seed, kind=0)
    return x


def ints(N, lo, hi, seed=777):
    y = mb.MatrixCuda(N, 1, np.int32)
    y.fill_synthetic(seed, kind=1, lo=lo, hi=hi)
    return y


def main():
    which = sys.argv[1:] or ["1", "2", "3", "4", "5a", "5b"]
    if len(which) > 1:
        # one fresh process per config: within one process the timing of a config
        # depends on what was allocated and freed before it (up to 15 % on the
        # short kernels); fresh processes repeat to 0.1 %
        import subprocess
        for cfg in which:
            subprocess.run([sys.executable, os.path.abspath(__file__), cfg], check=False)
        return
    rng = np.random.default_rng(12345)
    mb.runtime.set_device(0)
    for cfg in which:
        if cfg == "1":
            N, K = 10_000, 100
            x = synth(N, K)
            y = mb.MatrixCuda(N, 1)
            y.fill_synthetic(5, kind=0, scale=2.0)
            beta = rng.standard_normal(K) / np.sqrt(K)
            t = timeit(lambda: mb.normal_id_glm_lpdf(y, x, 0.1, beta, 1.3), 200)
            byt, name = N * K * 8, "normal_id N=1e4 K=100, alpha/beta/sigma var"
        elif cfg == "2":
            N, K = 10_000_000, 256
            x, y = synth(N, K), ints(N, 0, 1)
            beta = rng.standard_normal(K) / np.sqrt(K)
            t = timeit(lambda: mb.bernoulli_logit_glm_lpmf(y, x, 0.1, beta), 20)
            byt, name = N * K * 8, "bernoulli_logit N=1e7 K=256, beta var"
        elif cfg == "3":
            N, K = 100_000_000, 64
            x, y = synth(N, K), ints(N, 0, 4)
            beta = rng.standard_normal(K) / np.sqrt(K)
            t = timeit(lambda: mb.poisson_log_glm_lpmf(y, x, 0.1, beta), 5)
            byt, name = N * K * 8, "poisson_log N=1e8 K=64 on ONE GPU (51.2 GB)"
        elif cfg == "4":
            N, K = 10_000_000, 128
            x, y = synth(N, K), ints(N, 0, 4)
            beta = rng.standard_normal(K) / np.sqrt(K)
            dx = [None]

            def f():
                r = mb.neg_binomial_2_log_glm_lpmf(y, x, 0.1, beta, 2.5,
                                                   var=("x", "alpha", "beta", "phi"))
                dx[0] = r.d_x
            t = timeit(f, 10)
            byt, name = 2 * N * K * 8, "neg_binomial_2_log N=1e7 K=128, phi var + x var (writes N x K)"
        elif cfg == "4b":
            N, K = 10_000_000, 128
            x, y = synth(N, K), ints(N, 0, 4)
            beta = rng.standard_normal(K) / np.sqrt(K)
            t = timeit(lambda: mb.neg_binomial_2_log_glm_lpmf(y, x, 0.1, beta, 2.5), 10)
            byt, name = N * K * 8, "neg_binomial_2_log N=1e7 K=128, phi var, x data"
        elif cfg == "5a":
            N, K, C = 2_000_000, 512, 32
            x, y = synth(N, K), ints(N, 1, C)
            beta = np.asfortranarray(rng.standard_normal((K, C)) / np.sqrt(K))
            alpha = 0.1 * rng.standard_normal(C)
            t = timeit(lambda: mb.categorical_logit_glm_lpmf(y, x, alpha, beta), 5)
            byt, name = N * K * 8, "categorical_logit N=2e6 K=512 C=32"
            fl = 4.0 * N * K * C
            print(json.dumps({"config": cfg, "name": name, "ms_per_eval": t * 1e3,
                              "fp64_tflops": fl / t / 1e12,
                              "x_bytes_GBps_single_read": byt / t / 1e9}), flush=True)
            del x
            continue
        elif cfg == "5b":
            N, K = 10_000_000, 64
            x, y = synth(N, K), ints(N, 1, 9)
            beta = rng.standard_normal(K) / np.sqrt(K)
            cuts = np.linspace(-2, 2, 8)
            t = timeit(lambda: mb.ordered_logistic_glm_lpmf(y, x, beta, cuts), 20)
            byt, name = N * K * 8, "ordered_logistic N=1e7 K=64, 8 cuts"
        elif cfg == "u2":
            # the same model as config 2 built step by step on the
            # device -- theta = x beta + alpha, bernoulli_logit_lpmf(y | theta),
            # d_beta = x^T d_theta -- two sweeps over x instead of one
            N, K = 10_000_000, 256
            x, y = synth(N, K), ints(N, 0, 1)
            beta = rng.standard_normal(K) / np.sqrt(K)

            def unfused():
                theta = mb.lpmf.multiply(x, beta, 0.1)
                r = mb.lpmf.bernoulli_logit_lpmf(y, theta)
                return mb.lpmf.multiply_adjoint(x, r.d_theta)
            t = timeit(unfused, 20)
            t_mul = timeit(lambda: mb.lpmf.multiply(x, beta, 0.1), 20)
            byt, name = 2 * N * K * 8, ("un-fused bernoulli_logit N=1e7 K=256: multiply + "
                                        "lpmf + multiply_adjoint (2 sweeps over x); "
                                        f"multiply alone {t_mul*1e3:.3f} ms = "
                                        f"{N*K*8/t_mul/1e9:.0f} GB/s")
        else:
            continue
        print(json.dumps({"config": cfg, "name": name, "ms_per_eval": t * 1e3,
                          "evals_per_s": 1 / t, "GBps": byt / t / 1e9,
                          "frac_of_measured_hbm_peak": byt / t / 1e9 / PEAK}), flush=True)
        del x


if __name__ == "__main__":
    main()
