"""Config 4 (neg_binomial_2_log_glm_lpmf N=1e7 K=128, phi var + x var) in parts, same process:
4b (beta, phi var: the read-only sweep), the sweep that also leaves d in an N-vector (x var,
factored), the store stream alone, and the whole evaluation.  Device-timed (smc_timer)."""
import sys, json, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, math_b200 as mb
from math_b200._lib import lib, check
mb.runtime.set_device(0)
N, K = 10_000_000, 128
rng = np.random.default_rng(12345)
x = mb.MatrixCuda(N, K); x.fill_synthetic(12345, kind=0)
y = mb.MatrixCuda(N, 1, np.int32); y.fill_synthetic(777, kind=1, lo=0, hi=4)
beta = rng.standard_normal(K) / np.sqrt(K)
bp = beta.ctypes.data_as(C.POINTER(C.c_double))
dx = mb.MatrixCuda(N, K)
st = {}
def sweep_4b(): mb.neg_binomial_2_log_glm_lpmf(y, x, 0.1, beta, 2.5, var=("beta", "phi"))
def sweep_d(): st["r"] = mb.neg_binomial_2_log_glm_lpmf(y, x, 0.1, beta, 2.5, var=("x_factored", "beta", "phi"))
def outer(): check(lib().smc_matrix_outer(dx.handle, st["r"].d_x.handle, bp))
def both(): sweep_d(); outer()
def whole(): mb.neg_binomial_2_log_glm_lpmf(y, x, 0.1, beta, 2.5, var=("x", "beta", "phi"))
def timeit(f, reps=20):
    for _ in range(4): f()
    mb.runtime.synchronize()
    ms = C.c_double()
    check(lib().smc_timer_start())
    for _ in range(reps): f()
    check(lib().smc_timer_stop(C.byref(ms)))
    return ms.value / reps
out = {"N": N, "K": K}
for name, f in (("sweep_4b_ms", sweep_4b), ("sweep_leaving_d_ms", sweep_d), ("outer_ms", outer),
                ("sweep_then_outer_ms", both), ("whole_call_ms", whole), ("sweep_4b_again_ms", sweep_4b)):
    out[name] = round(timeit(f), 4)
print(json.dumps(out))
