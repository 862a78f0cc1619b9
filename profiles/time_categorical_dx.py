"""categorical_logit_glm_lpmf with x as an autodiff variable (writes the N x K d_x) at the
shape of BASELINE config 5a, next to the data-x evaluation.  Wall-clock, run alone."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, math_b200 as mb
mb.runtime.set_device(0)
N, K, C = 2_000_000, 512, 32
if len(sys.argv) > 3:
    N, K, C = (int(v) for v in sys.argv[1:4])
rng = np.random.default_rng(5)
x = mb.MatrixCuda(N, K); x.fill_synthetic(12345, kind=0)
y = mb.MatrixCuda(N, 1, np.int32); y.fill_synthetic(777, kind=1, lo=1, hi=C)
beta = np.asfortranarray(rng.standard_normal((K, C)) / np.sqrt(K))
alpha = rng.standard_normal(C) * 0.1
res = {"N": N, "K": K, "C": C}
for var in (("alpha", "beta"), ("x", "alpha", "beta")):
    f = lambda: mb.categorical_logit_glm_lpmf(y, x, alpha, beta, var=var)
    r = None
    for _ in range(3): r = f()
    mb.runtime.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): r = f()
    mb.runtime.synchronize()
    res["ms_" + "_".join(var)] = (time.perf_counter() - t0) / 10 * 1e3
res["d_x_extra_ms"] = res["ms_x_alpha_beta"] - res["ms_alpha_beta"]
res["d_x_write_GBps"] = N * K * 8 / res["d_x_extra_ms"] / 1e6
print(json.dumps(res))
