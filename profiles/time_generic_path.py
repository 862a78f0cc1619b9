import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, math_b200 as mb
mb.runtime.set_device(0)
rng = np.random.default_rng(1)
for N, K in [(4_000_000, 512), (6_000_000, 300), (4_000_000, 256)]:
    x = mb.MatrixCuda(N, K); x.fill_synthetic(12345, kind=0)
    y = mb.MatrixCuda(N, 1, np.int32); y.fill_synthetic(777, kind=1, lo=0, hi=1)
    beta = rng.standard_normal(K) / np.sqrt(K)
    f = lambda: mb.bernoulli_logit_glm_lpmf(y, x, 0.1, beta)
    f(); f(); mb.runtime.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): f()
    mb.runtime.synchronize()
    t = (time.perf_counter() - t0) / 10
    print(json.dumps({"N": N, "K": K, "ms": t * 1e3, "GBps_one_read": N * K * 8 / t / 1e9}))
    del x, y
