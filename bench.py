#!/usr/bin/env python
"""bench.py -- GLM lpdf+grad evaluations/sec on B200 (BASELINE.json's metric).

Workload (config.workload): BASELINE.json configs[1], the roofline headline --
bernoulli_logit_glm_lpmf, N = 1e7 rows x K = 256 columns of FP64 per GPU, alpha
and beta autodiff variables, x resident in HBM (20.48 GB).  A "step" is one
lpdf+grad evaluation: one pass of the fused kernel over x.

  value     evaluations/s with every input resident in HBM (device-side call,
            CUDA-event timed on the launching stream, max over ranks)
  e2e       the same evaluation through the public C-ABI call the Stan header
            binds (smc_bernoulli_logit_glm): HOST beta/alpha in, HOST logp +
            gradient out, every step; x stays resident -- that is the path's
            contract (uploaded once per model, reused by every HMC evaluation)
  roofline  algorithmic bytes N*K*8 per launch / measured kernel time, against
            MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the reference itself (oracle/_ref, reduce_sum over all host
            cores) timed on a bounded row sample of the same workload

N > 1 (torchrun, one rank per GPU): weak scaling -- every rank holds its own
N-row shard; one evaluation broadcasts the parameters, runs the fused kernel and
all-reduces the K+8 packed partials over NCCL.  value counts N-row units/s.

--impl reference times the reference's own CPU implementation on host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_ROWS = 10_000_000
K_COLS = 256
SEED = 12345
METRIC = "GLM lpdf+grad evals/sec (bernoulli_logit_glm_lpmf, N=1e7 K=256 per GPU)"
UNIT = "evals/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rows", type=int, default=N_ROWS)
    p.add_argument("--cols", type=int, default=K_COLS)
    p.add_argument("--cpu-rows", type=int, default=200_000,
                   help="row sample for the CPU baseline")
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


def make_params(K):
    rng = np.random.default_rng(SEED)
    beta = rng.standard_normal(K) / np.sqrt(K)
    return 0.1, beta


# --------------------------------------------------------------- clock sampling
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """Per-launch DRAM bytes of the fused kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("bernoulli_N1e7_K256_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------ CPU baseline
def cpu_baseline(rows_full, K, sample_rows, reps=3):
    """The reference (oracle/_ref) on the host: reduce_sum over TBB on all cores
    when the threaded build is there, else the single-call path.  Bounded sample
    of the same workload; evaluations/s scaled linearly in N to the full size."""
    from oracle import pyoracle as po
    from math_b200.matrix_cuda import synthetic_host
    n = min(sample_rows, rows_full)
    x = synthetic_host(SEED, 0, n, K)
    y = synthetic_host(SEED + 1, 0, n, 1, kind=1, lo=0, hi=1).ravel()
    alpha, beta = make_params(K)
    cores = os.cpu_count() or 1
    if po.ref_available(mt=True) and cores > 1:
        kind, threads = "reference", cores
        sec, _, _ = po.ref_time("bernoulli", y, x, alpha, beta, reps=reps,
                                threads=threads)
        how = f"reduce_sum over TBB, {threads} threads"
    elif po.ref_available():
        kind, threads = "reference", 1
        sec, _, _ = po.ref_time("bernoulli", y, x, alpha, beta, reps=reps)
        how = "single call, 1 thread"
    else:
        kind, threads = "port", 1
        t0 = time.perf_counter()
        po.bernoulli_logit_glm(y, x, alpha, beta)
        sec = time.perf_counter() - t0
        how = "C oracle port, 1 thread"
    evals = (n / rows_full) / sec
    out = {"value": evals, "unit": UNIT, "cores": threads, "kind": kind,
           "sample": f"{n} of {rows_full} rows x K={K} ({how}; best of {reps}; "
                     f"{sec*1e3:.1f} ms per sample eval, scaled linearly in N)",
           "sec_per_sample_eval": sec}
    if kind == "reference" and threads > 1:
        # the reference's plain single call on one core, same sample
        try:
            sec1, _, _ = po.ref_time("bernoulli", y, x, alpha, beta, reps=2,
                                     single_call_in_mt_lib=True)
            out["single_call"] = {"value": (n / rows_full) / sec1, "unit": UNIT,
                                  "cores": 1, "sec_per_sample_eval": sec1}
        except Exception as e:
            out["single_call"] = {"value": None, "note": f"failed: {e}"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    from math_b200.matrix_cuda import synthetic_host
    n = min(args.cpu_rows, args.rows)
    K = args.cols
    x = synthetic_host(SEED, 0, n, K)
    y = synthetic_host(SEED + 1, 0, n, 1, kind=1, lo=0, hi=1).ravel()
    alpha, beta = make_params(K)
    cores = os.cpu_count() or 1
    mt = po.ref_available(mt=True) and cores > 1
    if not po.ref_available():
        # the oracle always exists: fall back to the C port, single thread
        def step():
            po.bernoulli_logit_glm(y, x, alpha, beta)
        kind, threads = "port", 1
    else:
        threads = cores if mt else 1

        def step():
            po.ref_time("bernoulli", y, x, alpha, beta, reps=1, threads=threads)
        kind = "reference"
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = args.gpus * args.steps * (n / args.rows) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, extra={"timed_on": "host CPU"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"each step = one lpdf+grad on {n} of {args.rows} "
                                   f"rows x K={K}, scaled linearly in N"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, extra=None):
    c = {"workload": f"bernoulli_logit_glm_lpmf N={args.rows} K={args.cols} per GPU, "
                     "alpha+beta var, x resident (BASELINE.json configs[1])",
         "rows_per_gpu": args.rows, "cols": args.cols,
         "l2": "inputs (N*K*8 bytes per GPU) far exceed the 126 MB L2; no flush needed",
         "parallelism": f"row-sharded x{args.gpus}, params broadcast + packed all-reduce"
                        if args.gpus > 1 else "single GPU"}
    if extra:
        c.update(extra)
    return c


def cpp_drop_in(rows, cols, steps, warmup):
    """The same evaluation through the C++ drop-in a Stan model calls
    (stan::math::bernoulli_logit_glm_lpmf on matrix_cuda with var alpha / beta,
    grad(), recover_memory()): tests/cpp/_build/glm_bench, built against the
    reference's own headers where /root/reference exists."""
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "glm_bench")
    if not os.path.exists(exe):
        return {"value": None, "note": "tests/cpp/_build/glm_bench not built"}
    try:
        p = subprocess.run([exe, str(rows), str(cols), str(steps), str(warmup)],
                           capture_output=True, text=True, timeout=600)
        r = json.loads(p.stdout.strip().splitlines()[-1])
        return {"value": r["evals_per_s"], "unit": UNIT, "ms_per_eval": r["ms_per_eval"],
                "call": r["call"], "logp_per_row": r["logp_per_row"]}
    except Exception as e:
        return {"value": None, "note": f"failed: {e}"}


# ----------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    import math_b200 as mb
    from math_b200 import _lib
    from math_b200.sharded import ShardedGlm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- there is no CPU fallback")
    torch.cuda.set_device(local)
    mb.runtime.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a real (non-legacy-default) stream shared by the library's launches, the
    # NCCL collectives and the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    mb.runtime.set_stream(stream.cuda_stream)

    N, K = args.rows, args.cols
    row0 = rank * N  # weak scaling: rank r owns global rows [r*N, (r+1)*N)
    x = mb.MatrixCuda(N, K)
    x.fill_synthetic(SEED, row0=row0, kind=0, scale=1.0)
    y = mb.MatrixCuda(N, 1, np.int32)
    y.fill_synthetic(SEED + 1, row0=row0, kind=1, lo=0, hi=1)
    alpha, beta = make_params(K)
    flags = _lib.VAR_ALPHA | _lib.VAR_BETA
    glm = ShardedGlm("bernoulli_logit", y, x, K, alpha=alpha, flags=flags,
                     device=f"cuda:{local}")
    mb.runtime.synchronize()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident evaluation --------------------------------
    glm.evaluate(beta)  # parameters -> device once; resident from here on
    for _ in range(args.warmup):
        glm.evaluate()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    mb.runtime.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        glm.evaluate()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = mb.runtime.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    out_host = glm.out.cpu().numpy()

    # ---- kernel-only time for the roofline (same stream, same inputs) ------
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record(stream)
    for _ in range(args.steps):
        from math_b200.sharded import cuda_local_eval
        cuda_local_eval("bernoulli_logit", y, x, alpha, None, glm.params, 0, flags,
                        glm.out)
    k1.record(stream)
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / args.steps

    # ---- e2e: public C-ABI call, host params in / host results out ---------
    for _ in range(2):
        r = mb.bernoulli_logit_glm_lpmf(y, x, alpha, beta)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = mb.bernoulli_logit_glm_lpmf(y, x, alpha, beta)
    e2e_s = time.perf_counter() - t0
    barrier()

    t = torch.tensor([ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64,
                     device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kernel_ms = (float(v) for v in t.cpu())

    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_per_launch = N * K * 8
        achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": world * args.steps / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (counter-based hash, seed 12345; random-init beta)",
            "config": workload_config(args),
            "e2e": {"value": world * args.steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": (K + 1) * 8,
                    "d2h_bytes_per_step": (_lib.OUT_HEADER + K) * 8,
                    "note": "smc_bernoulli_logit_glm per step: host beta/alpha -> "
                            "kernel parameters, packed result -> pinned host memory, "
                            "stream sync; x resident by contract; wall-clock timed"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(),
                         "kernel": "glm_fused_kernel<bernoulli>",
                         "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "peak_source": peak_src},
            "check": {"logp_per_row": float(out_host[0]) / (N * world),
                      "nonfinite_rows": float(out_host[3])},
        }
        if world == 1:
            # frees this process's 20 GB first: the C++ binary allocates its own x
            del glm, x, y
            mb.runtime.synchronize()
            _lib.lib().smc_trim_cache()
            line["e2e"]["cpp_drop_in"] = cpp_drop_in(N, K, args.steps, args.warmup)
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(N, K, args.cpu_rows)
            except Exception as e:  # the baseline must never sink the bench line
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0,
                                        "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
