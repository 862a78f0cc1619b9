"""The N x K reverse sweep of an autodiff design matrix (x.adj += lp.adj * d_x, smc_matrix_axpy)
and the lazy finiteness scan (smc_matrix_all_finite) at the shape of BASELINE config 4
(N=1e7, K=128: 10.24 GB per matrix).  Wall-clock, run alone."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, math_b200 as mb
mb.runtime.set_device(0)
N, K = 10_000_000, 128
a = mb.MatrixCuda(N, K); a.fill_synthetic(1, kind=0)
b = mb.MatrixCuda(N, K); b.fill_synthetic(2, kind=0)
def timeit(f, reps=10):
    for _ in range(3): f()
    mb.runtime.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    mb.runtime.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
t_axpy = timeit(lambda: a.axpy(0.5, b))
t_fin = timeit(lambda: a.all_finite())
print(json.dumps({"N": N, "K": K, "axpy_ms": t_axpy, "axpy_GBps": 3 * N * K * 8 / t_axpy / 1e6,
                  "all_finite_ms": t_fin, "all_finite_GBps": N * K * 8 / t_fin / 1e6}))
