"""Un-fused categorical pipeline at BASELINE config 5a (N=2e6, K=512, C=32) next to the fused
GLM: multiply_matrix (DMMA sweep 1) -> categorical_logit_lpmf -> multiply_matrix_adjoint (DMMA
sweep 2 + column sums).  Wall-clock over synchronous calls, run alone ."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, math_b200 as mb
mb.runtime.set_device(0)
N, K, C = 2_000_000, 512, 32
rng = np.random.default_rng(5)
x = mb.MatrixCuda(N, K); x.fill_synthetic(12345, kind=0)
y = mb.MatrixCuda(N, 1, np.int32); y.fill_synthetic(777, kind=1, lo=1, hi=C)
beta = np.asfortranarray(rng.standard_normal((K, C)) / np.sqrt(K))
alpha = rng.standard_normal(C) * 0.1


def timeit(f, reps=10):
    for _ in range(3): f()
    mb.runtime.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    mb.runtime.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


state = {}
def prod(): state["lin"] = mb.lpmf.multiply_matrix(x, beta, alpha)
def dens(): state["r"] = mb.lpmf.categorical_logit_lpmf(y, state["lin"])
def adj(): state["g"] = mb.lpmf.multiply_matrix_adjoint(x, state["r"].d_theta)
def pipeline(): prod(); dens(); adj()
def fused(): state["f"] = mb.categorical_logit_glm_lpmf(y, x, alpha, beta)

out = {"N": N, "K": K, "C": C}
out["multiply_matrix_ms"] = timeit(prod)
out["categorical_logit_lpmf_ms"] = timeit(dens)
out["multiply_matrix_adjoint_ms"] = timeit(adj)
out["unfused_pipeline_ms"] = timeit(pipeline)
out["fused_glm_ms"] = timeit(fused)
out["multiply_matrix_tflops"] = 2.0 * N * K * C / out["multiply_matrix_ms"] / 1e9
out["adjoint_tflops"] = 2.0 * N * K * C / out["multiply_matrix_adjoint_ms"] / 1e9
f, r, (g, cs) = state["f"], state["r"], state["g"]
out["logp_rel_diff"] = abs(r.logp - f.logp) / abs(f.logp)
out["d_beta_max_abs_diff"] = float(np.abs(g - f.d_beta).max())
out["d_alpha_max_abs_diff"] = float(np.abs(cs - f.d_alpha).max())
print(json.dumps(out))
