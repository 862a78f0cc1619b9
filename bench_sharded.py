#!/usr/bin/env python
"""ONE host process driving all GPUs through the C ABI's row-sharded matrices
(include/stanmath_cuda.h, "row-sharded matrices"): the form in which a Stan model uses
more than one B200 -- smc_shard_init, sharded x / y scattered once, then the ordinary
smc_<family>_glm call per evaluation (kernel on every GPU, parameters as kernel
arguments, NCCL all-reduce of the packed partials, one read-back).  bench.py runs this
in its own process (no torch, no torch.distributed) and reports it as
`single_process_sharded`.

    python bench_sharded.py N_GPUS STEPS WARMUP   -> one JSON line

  weak    BASELINE configs[1] per GPU (bernoulli N=1e7 K=256 rows per shard)
  strong  BASELINE configs[2] in total (poisson N=1e8 K=64 over the shards)
Each is timed twice over the same loop of public synchronous calls (host parameters
in, host value + gradient out): wall-clock (e2e) and CUDA events on every shard's
stream (device; the maximum over the shards).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
SEED = 12345


def measure(mb, G, steps, warmup):
    rt = mb.runtime
    n = rt.shard_init(G)
    out = {"n_shards": n, "reduce": rt.shard_reduce_mode()}

    def timed(fn):
        for _ in range(max(warmup, 2)):
            fn()
        rt.synchronize()
        rt.timer_start()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = fn()
        wall = (time.perf_counter() - t0) / steps * 1e3
        dev = rt.timer_stop() / steps
        return wall, dev, r

    rng = np.random.default_rng(SEED)
    # small problem first: sharded against one GPU (parity inside the same process)
    ns, K = 1_000_003, 256
    beta = rng.standard_normal(K) / np.sqrt(K)
    xs, ys = mb.MatrixCuda.sharded(ns, K), mb.MatrixCuda.sharded(ns, 1, np.int32)
    xs.fill_synthetic(SEED + 7, kind=0)
    ys.fill_synthetic(SEED + 8, kind=1, lo=0, hi=1)
    x1, y1 = mb.MatrixCuda(ns, K), mb.MatrixCuda(ns, 1, np.int32)
    x1.fill_synthetic(SEED + 7, kind=0)
    y1.fill_synthetic(SEED + 8, kind=1, lo=0, hi=1)
    rs = mb.bernoulli_logit_glm_lpmf(ys, xs, 0.1, beta)
    r1 = mb.bernoulli_logit_glm_lpmf(y1, x1, 0.1, beta)
    got = np.concatenate([[rs.logp, rs.d_alpha], rs.d_beta])
    want = np.concatenate([[r1.logp, r1.d_alpha], r1.d_beta])
    scale = np.maximum(np.abs(want), np.abs(r1.d_beta).max() * 1e-3)
    out["vs_single_gpu_rel"] = float(np.max(np.abs(got - want) / scale))
    del xs, ys, x1, y1

    # weak: config 2 per GPU
    N = 10_000_000
    xs, ys = mb.MatrixCuda.sharded(N * n, K), mb.MatrixCuda.sharded(N * n, 1, np.int32)
    xs.fill_synthetic(SEED, kind=0)
    ys.fill_synthetic(SEED + 1, kind=1, lo=0, hi=1)
    wall, dev, r = timed(lambda: mb.bernoulli_logit_glm_lpmf(ys, xs, 0.1, beta))
    out["weak"] = {"workload": f"bernoulli_logit_glm_lpmf N={N} K={K} per GPU x {n} GPUs, "
                               "one process, smc_bernoulli_logit_glm on sharded handles",
                   "value": n * 1e3 / wall, "unit": "evals/s (N=1e7-row units)",
                   "e2e_ms_per_eval": wall, "device_ms_per_eval": dev,
                   "aggregate_GBps": n * N * K * 8 / (dev * 1e-3) / 1e9,
                   "logp_per_row": r.logp / (N * n)}
    del xs, ys
    rt.synchronize()

    # strong: config 3 in total
    NS, KS = 100_000_000, 64
    beta_s = np.random.default_rng(SEED).standard_normal(KS) / np.sqrt(KS)
    xs, ys = mb.MatrixCuda.sharded(NS, KS), mb.MatrixCuda.sharded(NS, 1, np.int32)
    xs.fill_synthetic(SEED, kind=0)
    ys.fill_synthetic(777, kind=1, lo=0, hi=4)
    wall, dev, r = timed(lambda: mb.poisson_log_glm_lpmf(ys, xs, 0.1, beta_s))
    out["strong"] = {"workload": f"poisson_log_glm_lpmf N={NS} K={KS} in total over {n} GPUs, "
                                 "one process, smc_poisson_log_glm on sharded handles",
                     "e2e_ms_per_eval": wall, "device_ms_per_eval": dev,
                     "evals_per_s": 1e3 / wall,
                     "aggregate_GBps": NS * KS * 8 / (dev * 1e-3) / 1e9,
                     "logp_per_row": r.logp / NS}
    del xs, ys
    rt.shard_shutdown()
    return out


def main():
    G = int(sys.argv[1])
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    child = len(sys.argv) > 4 and sys.argv[4] == "--only"
    import math_b200 as mb
    mb.runtime.set_device(0)
    if child:  # one mode (whatever SMC_SHARD_REDUCE says), nothing else
        print(json.dumps(measure(mb, G, steps, warmup)), flush=True)
        return
    # the default reduction (small results: every shard's kernel stores its packed result and
    # a completion flag straight into pinned host memory, the host adds the G slots in shard
    # order -- no collective on the evaluation path) ...
    os.environ.pop("SMC_SHARD_REDUCE", None)
    out = measure(mb, G, steps, warmup)
    # ... then the same evaluations in a fresh process with SMC_SHARD_REDUCE=nccl: the packed
    # result all-reduced in place over NCCL on the compute streams, one read-back
    import subprocess
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), str(G), str(steps),
                            str(warmup), "--only"], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, SMC_SHARD_REDUCE="nccl"))
        lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
        out["nccl_all_reduce"] = json.loads(lines[-1]) if lines else {
            "error": f"rc={p.returncode} " + (p.stderr or "")[-300:]}
    except Exception as e:  # noqa: BLE001 -- the default mode's record stands on its own
        out["nccl_all_reduce"] = {"error": repr(e)[:300]}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
