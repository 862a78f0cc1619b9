"""Config 4 (neg_binomial_2_log_glm, N=1e7, K=128, phi var, x autodiff): the fused kernel that
writes d_x = beta (x) d in the same sweep, against the same evaluation split into a read-only
sweep (which leaves d in an N-vector) plus a write-only outer-product kernel."""
import sys, time, json, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, math_b200 as mb
from math_b200._lib import lib, check
mb.runtime.set_device(0)
N, K = 10_000_000, 128
rng = np.random.default_rng(12345)
x = mb.MatrixCuda(N, K); x.fill_synthetic(12345, kind=0)
y = mb.MatrixCuda(N, 1, np.int32); y.fill_synthetic(777, kind=1, lo=0, hi=4)
beta = rng.standard_normal(K) / np.sqrt(K)
alpha_vec = mb.MatrixCuda(N, 1); alpha_vec.zero(); check(lib().smc_matrix_add_scalar(alpha_vec.handle, 0.1))
dx = mb.MatrixCuda(N, K)
bp = beta.ctypes.data_as(C.POINTER(C.c_double))
def timeit(f, reps=10):
    for _ in range(3): f()
    mb.runtime.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    mb.runtime.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
st = {}
def fused(): st["f"] = mb.neg_binomial_2_log_glm_lpmf(y, x, 0.1, beta, 2.5, var=("x", "beta", "phi"))
def read_only(): st["r"] = mb.neg_binomial_2_log_glm_lpmf(y, x, alpha_vec, beta, 2.5, var=("alpha", "beta", "phi"))
def outer(): check(lib().smc_matrix_outer(dx.handle, st["r"].d_alpha.handle, bp))
def split(): read_only(); outer()
out = {"N": N, "K": K, "fused_ms": timeit(fused), "read_only_ms": timeit(read_only)}
out["outer_ms"] = timeit(outer)
out["split_ms"] = timeit(split)
out["outer_write_GBps"] = N * K * 8 / out["outer_ms"] / 1e6
a = st["f"].d_x.rows_to_host(5_000_000, 1000); b = dx.rows_to_host(5_000_000, 1000)
out["max_abs_diff_dx_block"] = float(np.abs(a - b).max())
print(json.dumps(out))
