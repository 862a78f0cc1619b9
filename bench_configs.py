#!/usr/bin/env python
"""One BASELINE.json config per process, measured three ways (bench.py's `configs`):

  ms_per_eval   device time (CUDA events on the launching stream) of the
                device-resident evaluation: parameters and the packed result stay in
                HBM, no host synchronisation inside the timed region
  e2e           wall-clock of the public synchronous C-ABI call the Stan header binds
                (host parameters in, host value + gradient out, x resident)
  roofline      algorithmic bytes (N K 8; 2 N K 8 when x is an autodiff variable) over
                ms_per_eval against the measured HBM peak -- or, for the categorical
                GLM, 4 N K C flop against the FP64 tensor rate measured on this GPU

A fresh process per config: inside one process the time of a short kernel depends on
what was allocated and freed before it (up to 15 %); fresh processes repeat to 0.1 %.
No torch here (ctypes over libstanmath_cuda.so only), so a process starts in a second.

    python bench_configs.py 4            -> one JSON line
    python bench_configs.py              -> every config, one line each
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALL = ["1", "2", "3", "4", "4b", "5a", "5b"]
SEED = 12345


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_one(cfg, reps=None):
    import math_b200 as mb
    from math_b200 import _lib
    from math_b200.sharded import cuda_local_eval, cuda_local_eval_categorical
    rt = mb.runtime
    rt.set_device(0)
    rng = np.random.default_rng(SEED)

    def synth(N, K):
        x = mb.MatrixCuda(N, K)
        x.fill_synthetic(SEED, kind=0)
        return x

    def ints(N, lo, hi):
        y = mb.MatrixCuda(N, 1, np.int32)
        y.fill_synthetic(777, kind=1, lo=lo, hi=hi)
        return y

    def dev_vec(a):
        return mb.to_matrix_cuda(np.asarray(a, dtype=np.float64))

    def wall(fn, n):
        fn()
        fn()
        rt.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        rt.synchronize()
        return (time.perf_counter() - t0) / n * 1e3

    def device(fn, n):
        fn()
        fn()
        rt.synchronize()
        rt.timer_start()
        for _ in range(n):
            fn()
        return rt.timer_stop() / n

    rec = {"id": cfg}
    flops = None
    if cfg == "1":
        N, K = 10_000, 100
        x = synth(N, K)
        y = mb.MatrixCuda(N, 1)
        y.fill_synthetic(5, kind=0, scale=2.0)
        beta = rng.standard_normal(K) / np.sqrt(K)
        flags = _lib.VAR_ALPHA | _lib.VAR_BETA | _lib.VAR_AUX
        params, out = dev_vec(beta), mb.MatrixCuda(_lib.OUT_HEADER + K, 1)
        n = reps or 500
        # the launches must come faster than the ~10 us kernel or the figure is the host's
        # call rate: the C entry point with pre-converted arguments, not the Python wrapper
        from math_b200.sharded import FAMILY, _ptr
        fn = _lib.lib().smc_glm_eval_device
        args = (FAMILY["normal_id"], y.handle, 0.0, x.handle, None, 0.1, None, 1.3,
                _ptr(params), 0, int(flags), _ptr(out), None, None, None, None)
        ms = device(lambda: fn(*args), n)
        e2e = wall(lambda: mb.normal_id_glm_lpdf(y, x, 0.1, beta, 1.3), n)
        byt = N * K * 8
        rec["workload"] = ("normal_id_glm_lpdf N=1e4 K=100, alpha/beta/sigma var (8 MB: "
                           "L2-resident, latency-bound)")
    elif cfg in ("2", "3", "4b", "5b"):
        fam, N, K, lo, hi, name = {
            "2": ("bernoulli_logit", 10_000_000, 256, 0, 1,
                  "bernoulli_logit_glm_lpmf N=1e7 K=256, alpha+beta var"),
            "3": ("poisson_log", 100_000_000, 64, 0, 4,
                  "poisson_log_glm_lpmf N=1e8 K=64 on ONE GPU (51.2 GB), alpha+beta var"),
            "4b": ("neg_binomial_2_log", 10_000_000, 128, 0, 4,
                   "neg_binomial_2_log_glm_lpmf N=1e7 K=128, alpha+beta+phi var, x data"),
            "5b": ("ordered_logistic", 10_000_000, 64, 1, 9,
                   "ordered_logistic_glm_lpmf N=1e7 K=64, 8 cuts, beta+cuts var"),
        }[cfg]
        x, y = synth(N, K), ints(N, lo, hi)
        beta = rng.standard_normal(K) / np.sqrt(K)
        cuts = np.linspace(-2, 2, 8)
        ncuts = 8 if cfg == "5b" else 0
        flags = _lib.VAR_BETA | (_lib.VAR_AUX if cfg in ("4b", "5b") else 0) \
            | (0 if cfg == "5b" else _lib.VAR_ALPHA)
        params = dev_vec(np.concatenate([beta, cuts]) if ncuts else beta)
        out = mb.MatrixCuda(_lib.OUT_HEADER + K + ncuts, 1)
        aux = 2.5 if cfg == "4b" else None
        n = reps or {"2": 20, "3": 8, "4b": 20, "5b": 30}[cfg]
        ms = device(lambda: cuda_local_eval(fam, y, x, 0.0 if cfg == "5b" else 0.1, aux, params, ncuts, flags, out), n)
        pub = {"2": lambda: mb.bernoulli_logit_glm_lpmf(y, x, 0.1, beta),
               "3": lambda: mb.poisson_log_glm_lpmf(y, x, 0.1, beta),
               "4b": lambda: mb.neg_binomial_2_log_glm_lpmf(y, x, 0.1, beta, 2.5),
               "5b": lambda: mb.ordered_logistic_glm_lpmf(y, x, beta, cuts)}[cfg]
        e2e = wall(pub, n)
        byt = N * K * 8
        rec["workload"] = name
    elif cfg == "4":
        # x is an autodiff variable: forward sweep leaves the factor d of d_x = d beta^T,
        # the reverse sweep stores lp.adj * d beta^T into the (lazily zero) N x K adjoint
        N, K = 10_000_000, 128
        x, y = synth(N, K), ints(N, 0, 4)
        beta = rng.standard_normal(K) / np.sqrt(K)
        flags = _lib.VAR_X | _lib.DX_FACTORED | _lib.VAR_ALPHA | _lib.VAR_BETA | _lib.VAR_AUX
        params, out = dev_vec(beta), mb.MatrixCuda(_lib.OUT_HEADER + K, 1)
        dvec, adj = mb.MatrixCuda(N, 1), mb.MatrixCuda(N, K)

        def step():
            cuda_local_eval("neg_binomial_2_log", y, x, 0.1, 2.5, params, 0, flags, out, d_x=dvec)
            adj.zero_lazy()
            adj.rank1_update(1.0, dvec, beta)
        n = reps or 10
        ms = device(step, n)

        def pub():
            r = mb.neg_binomial_2_log_glm_lpmf(y, x, 0.1, beta, 2.5,
                                               var=("x_factored", "alpha", "beta", "phi"))
            adj.zero_lazy()
            adj.rank1_update(1.0, r.d_x, beta)
        e2e = wall(pub, n)
        byt = 2 * N * K * 8
        rec["workload"] = ("neg_binomial_2_log_glm_lpmf N=1e7 K=128, phi var + x var: forward "
                           "sweep + reverse sweep writing the N x K adjoint of x")
    elif cfg == "5a":
        N, K, Cc = 2_000_000, 512, 32
        x, y = synth(N, K), ints(N, 1, Cc)
        beta = np.asfortranarray(rng.standard_normal((K, Cc)) / np.sqrt(K))
        alpha = 0.1 * rng.standard_normal(Cc)
        flags = _lib.VAR_ALPHA | _lib.VAR_BETA
        params = dev_vec(np.concatenate([beta.ravel(order="F"), alpha]))
        out = mb.MatrixCuda(2 + Cc + K * Cc, 1)
        n = reps or 8
        ms = device(lambda: cuda_local_eval_categorical(y, x, params, Cc, flags, out), n)
        e2e = wall(lambda: mb.categorical_logit_glm_lpmf(y, x, alpha, beta), n)
        byt = N * K * 8
        flops = 4.0 * N * K * Cc
        rec["workload"] = "categorical_logit_glm_lpmf N=2e6 K=512 C=32, alpha+beta var"
    else:
        raise SystemExit(f"unknown config {cfg}")

    rec["ms_per_eval"] = ms
    rec["evals_per_s"] = 1e3 / ms
    rec["e2e"] = {"ms_per_eval": e2e, "value": 1e3 / e2e, "unit": "evals/s"}
    if flops is None:
        peak, src = hbm_peak()
        ach = byt / (ms * 1e-3) / 1e9
        rec["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                           "frac": ach / peak, "algorithmic_bytes_per_eval": byt,
                           "peak_source": src}
    else:
        peak = rt.measure_dmma_peak()
        ach = flops / (ms * 1e-3) / 1e12
        rec["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                           "frac": ach / peak, "algorithmic_flop_per_eval": flops,
                           "peak_source": "FP64 DMMA rate measured from registers on this "
                                          "GPU in this process (smc_measure_dmma_peak)",
                           "x_sweeps": 2, "hbm_GBps_two_sweeps": 2 * byt / (ms * 1e-3) / 1e9}
    return rec


def cpp_drop_in(mode, rows, cols, steps, warmup):
    """The same config through the C++ overload on the reference's own tape
    (tests/cpp/_build/glm_bench: call + grad() + recover_memory())."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "glm_bench")
    if not os.path.exists(exe):
        return {"ms_per_eval": None, "note": "tests/cpp/_build/glm_bench not built"}
    try:
        p = subprocess.run([exe, str(rows), str(cols), str(steps), str(warmup), mode],
                           capture_output=True, text=True, timeout=600)
        r = json.loads(p.stdout.strip().splitlines()[-1])
        return {"ms_per_eval": r["ms_per_eval"], "value": r["evals_per_s"], "unit": "evals/s",
                "call": r["call"]}
    except Exception as e:  # noqa: BLE001
        return {"ms_per_eval": None, "note": f"failed: {e}"}


def run_all(which=None, timeout=300):
    """Every config in its own process; returns the list of records."""
    out = []
    for cfg in which or ALL:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), cfg],
                               capture_output=True, text=True, timeout=timeout)
            rec = json.loads(p.stdout.strip().splitlines()[-1])
        except Exception as e:  # noqa: BLE001
            rec = {"id": cfg, "ms_per_eval": None, "note": f"failed: {e}"}
        if cfg == "1":
            rec["cpp_drop_in"] = cpp_drop_in("normal", 10_000, 100, 3000, 200)
        elif cfg == "4":
            rec["cpp_drop_in"] = cpp_drop_in("negbin_xvar", 10_000_000, 128, 10, 3)
        out.append(rec)
    return out


if __name__ == "__main__":
    if len(sys.argv) == 2:
        print(json.dumps(run_one(sys.argv[1])), flush=True)
    else:
        for r in run_all(sys.argv[1:] or None):
            print(json.dumps(r), flush=True)
