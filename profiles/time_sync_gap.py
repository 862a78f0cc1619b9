"""Where the gap between the device-timed evaluation and the synchronous call goes (headline:
bernoulli N=1e7 K=256).  (a) 20 asynchronous evaluations back to back, CUDA events around the loop;
(b) 20 synchronous C-ABI calls, wall clock; (c) the same calls with CUDA events around EACH one
(the kernel alone, started on an idle GPU); (d) SMC_SPIN_US variants come from the environment."""
import sys, time, json, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, math_b200 as mb
from math_b200._lib import lib, check
mb.runtime.set_device(0)
N, K = 10_000_000, 256
x = mb.MatrixCuda(N, K); x.fill_synthetic(12345, kind=0)
y = mb.MatrixCuda(N, 1, np.int32); y.fill_synthetic(12346, kind=1, lo=0, hi=1)
beta = np.random.default_rng(12345).standard_normal(K) / np.sqrt(K)
f = lambda: mb.bernoulli_logit_glm_lpmf(y, x, 0.1, beta)
for _ in range(5): f()
ms = C.c_double()
out = {}
# (b) wall clock over synchronous calls
t0 = time.perf_counter()
for _ in range(20): f()
out["sync_wall_ms"] = (time.perf_counter() - t0) / 20 * 1e3
# (c) events around each synchronous call
per = []
for _ in range(20):
    check(lib().smc_timer_start()); f(); check(lib().smc_timer_stop(C.byref(ms))); per.append(ms.value)
out["sync_kernel_event_ms_median"] = float(np.median(per))
out["sync_kernel_event_ms_min"] = float(np.min(per))
# pure Python + C overhead of the call path with a tiny problem (kernel ~8 us)
xs = mb.MatrixCuda(1000, K); xs.fill_synthetic(1, kind=0)
ys = mb.MatrixCuda(1000, 1, np.int32); ys.fill_synthetic(2, kind=1, lo=0, hi=1)
g = lambda: mb.bernoulli_logit_glm_lpmf(ys, xs, 0.1, beta)
for _ in range(50): g()
t0 = time.perf_counter()
for _ in range(2000): g()
out["tiny_call_wall_us"] = (time.perf_counter() - t0) / 2000 * 1e6
# (a) events around 20 synchronous calls in a row (device busy + gaps)
check(lib().smc_timer_start())
for _ in range(20): f()
check(lib().smc_timer_stop(C.byref(ms)))
out["sync_loop_event_ms"] = ms.value / 20
print(json.dumps(out))
