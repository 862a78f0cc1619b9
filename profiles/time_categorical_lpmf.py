"""Timing of the un-fused categorical_logit_lpmf on a device N x C matrix of log odds
(value + d_lin: reads N*C*8, writes N*C*8 bytes).  Wall-clock over synchronous C-ABI
calls; run alone on the GPU ."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, math_b200 as mb
mb.runtime.set_device(0)
SHAPES = [(2_000_000, 32), (10_000_000, 8), (10_000_000, 32), (4_000_000, 64), (2_000_000, 128)]
if len(sys.argv) > 2:
    SHAPES = [(int(sys.argv[1]), int(sys.argv[2]))]
for N, C in SHAPES:
    lin = mb.MatrixCuda(N, C); lin.fill_synthetic(12345, kind=0)
    y = mb.MatrixCuda(N, 1, np.int32); y.fill_synthetic(777, kind=1, lo=1, hi=C)
    for lin_var in (True, False):
        f = lambda: mb.lpmf.categorical_logit_lpmf(y, lin, lin_var=lin_var)
        r = None
        for _ in range(4): r = f()
        mb.runtime.synchronize()
        t0 = time.perf_counter()
        for _ in range(20): r = f()
        mb.runtime.synchronize()
        t = (time.perf_counter() - t0) / 20
        nbytes = N * C * 8 * (2 if lin_var else 1)
        print(json.dumps({"N": N, "C": C, "lin_var": lin_var, "ms": round(t * 1e3, 4),
                          "GBps": round(nbytes / t / 1e9, 1), "logp_per_row": r.logp / N}))
    del lin, y
