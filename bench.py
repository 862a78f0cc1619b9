#!/usr/bin/env python
"""bench.py -- GLM lpdf+grad evaluations/sec on B200 (BASELINE.json's metric).

Headline workload (config.workload): BASELINE.json configs[1] -- bernoulli_logit_glm_lpmf,
N = 1e7 rows x K = 256 columns of FP64 per GPU, alpha and beta autodiff variables, x
resident in HBM (20.48 GB).  A "step" is one lpdf+grad evaluation: one pass of the fused
kernel over x.

  value     evaluations/s with every input resident in HBM (device-side call,
            CUDA-event timed on the launching stream, max over ranks)
  e2e       N=1: the same evaluation through the public C-ABI call the Stan header
            binds (smc_bernoulli_logit_glm): HOST beta/alpha in, HOST logp + gradient
            out, every step; x stays resident -- that is the path's contract (uploaded
            once per model, reused by every HMC evaluation).
            N>1: the sharded evaluation end to end: host parameters on rank 0 ->
            broadcast -> fused kernel on every rank -> NCCL all-reduce of the packed
            K+8 partials -> host result on rank 0, every step.
  roofline  algorithmic bytes N*K*8 per launch / measured kernel time, against
            MEASURED_PEAKS.json's HBM copy bandwidth
  configs   every other BASELINE.json config (1, 3 on one GPU, 4, 4b, 5a, 5b), each
            measured in a fresh process by bench_configs.py (N=1 only)
  strong    N>1: BASELINE configs[2], poisson_log_glm_lpmf N=1e8 K=64 in total,
            row-sharded over the ranks; efficiency against the same problem timed on
            rank 0's GPU alone in the same run
  single_process_sharded   N>1: the same two measurements through the C ABI's
            row-sharded matrices -- ONE host process driving all N GPUs
            (smc_shard_init / smc_sharded_matrix_create), the form a Stan model uses
  check     N>1: vs_single_gpu_rel, the sharded result of a small problem against the
            same problem on one GPU (NCCL parity inside the driver's own run)
  cpu_baseline  the reference itself (oracle/_ref): reduce_sum over all host cores and
            the single call, on the FULL workload when host RAM holds it; `opencl`:
            the reference's STAN_OPENCL path on this GPU where an ICD answers

N > 1 (torchrun, one rank per GPU): weak scaling -- every rank holds its own N-row
shard; value counts N-row units/s.

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
STRONG_ROWS, STRONG_COLS = 100_000_000, 64


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--rows", type=int, default=N_ROWS)
    p.add_argument("--cols", type=int, default=K_COLS)
    p.add_argument("--cpu-rows", type=int, default=0,
                   help="rows timed by the CPU arms (0: the full workload when host RAM "
                        "holds it, else the largest sample that fits)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-configs", action="store_true",
                   help="skip the per-config table / strong scaling / single-process legs")
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


def ncu_traffic(live=True):
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the fused
    kernel on the headline workload.  Measured in this run when ncu is on the box: after the
    timed region, config 2 is run once more in a fresh process under `ncu --metrics ...`
    (one launch captured; a number printed under the profiler is never a bench value, only
    the byte counters are read).  Otherwise the figure of the committed capture
    (profiles/traffic.json).  Returns (bytes, source)."""
    committed = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            committed = json.load(f).get("bernoulli_N1e7_K256_bytes_per_launch")
    except Exception:
        pass
    committed_src = ("profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of "
                     "this kernel from the committed ncu --set full capture of this command")
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not live or not os.path.exists(ncu):
        return committed, committed_src
    try:
        cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum",
               "--clock-control", "none", "-k", "regex:glm_fused", "-s", "3", "-c", "1",
               "--csv", sys.executable, os.path.join(ROOT, "bench_configs.py"), "2"]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
        total, unit_scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
                                  "Tbyte": 1e12}
        import csv as _csv
        rows = [r for r in _csv.reader(p.stdout.splitlines()) if len(r) > 3]
        hdr = next(r for r in rows if "Metric Name" in r)
        i_name, i_unit, i_val = (hdr.index("Metric Name"), hdr.index("Metric Unit"),
                                 hdr.index("Metric Value"))
        seen = 0
        for r in rows:
            if r is hdr or len(r) <= i_val or not r[i_name].startswith("dram__bytes_"):
                continue
            total += float(r[i_val].replace(",", "")) * unit_scale.get(r[i_unit], 1.0)
            seen += 1
        if seen == 2 and total > 0:
            return total, ("measured in this run: ncu --metrics dram__bytes_read.sum,"
                           "dram__bytes_write.sum, one launch of glm_fused_kernel<bernoulli> "
                           "on N=1e7 K=256 in a fresh process after the timed region")
    except Exception:
        pass
    return committed, committed_src


# ------------------------------------------------------------------ CPU baseline
def host_ram_gb():
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1048576.0
    except Exception:
        pass
    return 0.0


def cpu_rows(rows_full, K, requested):
    """Rows the CPU arms time: the full workload when the host holds x twice over
    (the inputs plus the reference's own temporaries), else the largest 1-2-5 sample."""
    if requested:
        return min(requested, rows_full)
    ram = host_ram_gb()
    n = rows_full
    steps = [1.0, 0.5, 0.2, 0.1, 0.05, 0.02, 0.01]
    for s in steps:
        n = int(rows_full * s)
        if n * K * 8 * 2.5 / 2**30 < ram:
            break
    return max(n, 1)


def host_inputs(n, K):
    from oracle import pyoracle as po
    x = po.synthetic(SEED, 0, n, K)
    y = po.synthetic(SEED + 1, 0, n, 1, kind=1, lo=0, hi=1).ravel()
    return x, y


def opencl_baseline(rows_full, K):
    """The reference's STAN_OPENCL path on this GPU (own process, see
    oracle/opencl_baseline.py); rows capped so that rows * K stays a 32-bit index."""
    rows = min(rows_full, (2**31 - 1) // K // 1_000_000 * 1_000_000)
    exe = os.path.join(ROOT, "oracle", "opencl_baseline.py")
    try:
        p = subprocess.run([sys.executable, exe, str(rows), str(K), "3"],
                           capture_output=True, text=True, timeout=420)
        lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
        if not lines:
            tail = (p.stderr or "").strip().splitlines()[-1:] or ["no output"]
            return {"available": False, "why": f"rc {p.returncode}: {tail[0][:200]}"}
        r = json.loads(lines[-1])
        if r.get("available"):
            r["value"] = (rows / rows_full) / r["sec_per_eval"]
            r["unit"] = UNIT
            r["sample"] = (f"{rows} of {rows_full} rows x K={K}, scaled linearly in N (the "
                           "reference's OpenCL kernels index x with 32-bit integers); x, y "
                           "resident as matrix_cl, var alpha/beta, grad() per evaluation")
        return r
    except Exception as e:  # noqa: BLE001
        return {"available": False, "why": f"failed: {e}"}


def cpu_baseline(rows_full, K, requested_rows, reps=2):
    """The reference (oracle/_ref) on the host: reduce_sum over TBB on all cores when
    the threaded build is there, else the single-call path.  Evaluations/s of the
    rows_full-row workload (scaled linearly in N only when a sample had to be used)."""
    from oracle import pyoracle as po
    n = cpu_rows(rows_full, K, requested_rows)
    x, y = host_inputs(n, K)
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
    scaled = "" if n == rows_full else ", scaled linearly in N"
    out = {"value": evals, "unit": UNIT, "cores": threads, "kind": kind,
           "sample": f"{n} of {rows_full} rows x K={K} ({how}; best of {reps}; "
                     f"{sec*1e3:.1f} ms per evaluation of those rows{scaled})",
           "rows_timed": n, "sec_per_sample_eval": sec,
           "host_ram_gb_available": round(host_ram_gb(), 1)}
    if kind == "reference" and threads > 1:
        # the reference's plain single call on one core
        try:
            n1 = min(n, 1_000_000)
            sec1, _, _ = po.ref_time("bernoulli", y[:n1], x[:n1], alpha, beta, reps=1,
                                     single_call_in_mt_lib=True)
            out["single_call"] = {"value": (n1 / rows_full) / sec1, "unit": UNIT,
                                  "cores": 1, "rows_timed": n1,
                                  "sec_per_sample_eval": sec1}
        except Exception as e:  # noqa: BLE001
            out["single_call"] = {"value": None, "note": f"failed: {e}"}
    del x, y
    out["opencl"] = opencl_baseline(rows_full, K)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    K = args.cols
    n = cpu_rows(args.rows, K, args.cpu_rows)
    x, y = host_inputs(n, K)
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
    # one step = one evaluation of n rows; a unit of the metric is args.rows rows.  The
    # host is the same whatever --gpus says: no multiplication by the GPU count.
    value = args.steps * (n / args.rows) / dt
    scaled = "" if n == args.rows else ", scaled linearly in N"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"each step = one lpdf+grad on {n} of {args.rows} "
                                   f"rows x K={K} on the host's {threads} threads{scaled}",
                         "rows_timed": n,
                         "host_ram_gb_available": round(host_ram_gb(), 1)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"bernoulli_logit_glm_lpmf N={args.rows} K={args.cols} per GPU, "
                        "alpha+beta var, x resident (BASELINE.json configs[1])",
            "rows_per_gpu": args.rows, "cols": args.cols,
            "l2": "inputs (N*K*8 bytes per GPU) far exceed the 126 MB L2; no flush needed",
            "parallelism": f"row-sharded x{args.gpus}, params broadcast + packed all-reduce"
                           if args.gpus > 1 else "single GPU"}


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
    except Exception as e:  # noqa: BLE001
        return {"value": None, "note": f"failed: {e}"}


def single_process_sharded(n_gpus, steps, warmup):
    """ONE process, all GPUs, through the C ABI's row-sharded matrices
    (bench_sharded.py in its own process: no torch, no torch.distributed)."""
    exe = os.path.join(ROOT, "bench_sharded.py")
    try:
        p = subprocess.run([sys.executable, exe, str(n_gpus), str(steps), str(warmup)],
                           capture_output=True, text=True, timeout=600,
                           env={k: v for k, v in os.environ.items()
                                if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")})
        lines = [ln for ln in p.stdout.strip().splitlines() if ln.startswith("{")]
        if not lines:
            return {"value": None,
                    "note": f"rc {p.returncode}: {(p.stderr or '').strip()[-300:]}"}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"value": None, "note": f"failed: {e}"}


# ----------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    import math_b200 as mb
    from math_b200 import _lib
    from math_b200.sharded import ShardedGlm, cuda_local_eval

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- there is no CPU fallback")
    torch.cuda.set_device(local)
    mb.runtime.set_device(local)
    host_pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_pg = dist.new_group(backend="gloo")  # host-only waits (no spinning kernels)
    # a real (non-legacy-default) stream shared by the library's launches, the
    # NCCL collectives and the timing events
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    mb.runtime.set_stream(stream.cuda_stream)
    dev = f"cuda:{local}"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_pg)

    def max_over_ranks(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    def device_time(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    N, K = args.rows, args.cols
    row0 = rank * N  # weak scaling: rank r owns global rows [r*N, (r+1)*N)
    x = mb.MatrixCuda(N, K)
    x.fill_synthetic(SEED, row0=row0, kind=0, scale=1.0)
    y = mb.MatrixCuda(N, 1, np.int32)
    y.fill_synthetic(SEED + 1, row0=row0, kind=1, lo=0, hi=1)
    alpha, beta = make_params(K)
    flags = _lib.VAR_ALPHA | _lib.VAR_BETA
    glm = ShardedGlm("bernoulli_logit", y, x, K, alpha=alpha, flags=flags, device=dev)
    mb.runtime.synchronize()
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
    ms = device_time(glm.evaluate, args.steps)
    launches = mb.runtime.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    out_host = glm.out.cpu().numpy()

    # ---- kernel-only time for the roofline (same stream, same inputs) ------
    kernel_ms = device_time(
        lambda: cuda_local_eval("bernoulli_logit", y, x, alpha, None, glm.params, 0, flags,
                                glm.out), args.steps) / args.steps

    # ---- e2e ---------------------------------------------------------------
    if world == 1:
        def e2e_step():
            return mb.bernoulli_logit_glm_lpmf(y, x, alpha, beta)
        e2e_note = ("smc_bernoulli_logit_glm per step: host beta/alpha -> kernel "
                    "parameters, packed result -> pinned host memory, stream sync; x "
                    "resident by contract; wall-clock timed")
    else:
        pinned = torch.empty(glm.out.numel(), dtype=torch.float64).pin_memory()

        def e2e_step():
            glm.evaluate(beta)  # rank 0: host parameters -> device; then broadcast
            if rank == 0:
                pinned.copy_(glm.out, non_blocking=True)
            stream.synchronize()
            return pinned
        e2e_note = ("per step: host beta on rank 0 -> device -> NCCL broadcast -> fused "
                    "kernel on every rank -> NCCL all-reduce of the packed K+8 doubles -> "
                    "pinned host memory on rank 0, stream sync; x shards resident; "
                    "wall-clock timed, max over ranks")
    for _ in range(max(3, args.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    barrier()
    ms, e2e_ms, kernel_ms = max_over_ranks([ms, e2e_s * 1e3, kernel_ms])

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        bytes_per_launch = N * K * 8
        achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        # (only for the default shape on one GPU: the counter run needs its own 20 GB)
        traffic, traffic_src = ncu_traffic(
            live=world == 1 and not args.no_configs and (N, K) == (10_000_000, 256))
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
                    "note": e2e_note},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "glm_fused_kernel<bernoulli>",
                         "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": bytes_per_launch,
                         "peak_source": peak_src},
            "check": {"logp_per_row": float(out_host[0]) / (N * world),
                      "nonfinite_rows": float(out_host[3])},
        }

    # ---- N>1: NCCL parity, strong scaling of config 3, single-process form ----
    del glm, x, y
    mb.runtime.synchronize()
    _lib.lib().smc_trim_cache()
    if world > 1:
        # (a) a small sharded problem against the same problem on one GPU
        n_small = 1_000_003
        lo, hi = (n_small * rank) // world, (n_small * (rank + 1)) // world
        xs = mb.MatrixCuda(hi - lo, K)
        xs.fill_synthetic(SEED + 7, row0=lo, kind=0)
        ys = mb.MatrixCuda(hi - lo, 1, np.int32)
        ys.fill_synthetic(SEED + 8, row0=lo, kind=1, lo=0, hi=1)
        g2 = ShardedGlm("bernoulli_logit", ys, xs, K, alpha=alpha, flags=flags, device=dev)
        sharded = g2.evaluate(beta).cpu().numpy().copy()
        if rank == 0:
            xf = mb.MatrixCuda(n_small, K)
            xf.fill_synthetic(SEED + 7, row0=0, kind=0)
            yf = mb.MatrixCuda(n_small, 1, np.int32)
            yf.fill_synthetic(SEED + 8, row0=0, kind=1, lo=0, hi=1)
            r1 = mb.bernoulli_logit_glm_lpmf(yf, xf, alpha, beta)
            got = np.concatenate([[sharded[0]], sharded[_lib.OUT_HEADER:_lib.OUT_HEADER + K]])
            want = np.concatenate([[r1.logp], r1.d_beta])
            scale = np.maximum(np.abs(want), np.abs(r1.d_beta).max() * 1e-3)
            line["check"]["vs_single_gpu_rel"] = float(np.max(np.abs(got - want) / scale))
            line["check"]["vs_single_gpu_problem"] = f"bernoulli N={n_small} K={K}"
            del xf, yf
        del g2, xs, ys
        host_barrier()

    if world > 1 and not args.no_configs:
        # (b) BASELINE configs[2]: poisson N=1e8 K=64 in total, row-sharded
        NS, KS = STRONG_ROWS, STRONG_COLS
        lo, hi = (NS * rank) // world, (NS * (rank + 1)) // world
        xs = mb.MatrixCuda(hi - lo, KS)
        xs.fill_synthetic(SEED, row0=lo, kind=0)
        ys = mb.MatrixCuda(hi - lo, 1, np.int32)
        ys.fill_synthetic(777, row0=lo, kind=1, lo=0, hi=4)
        rngs = np.random.default_rng(SEED)
        beta_s = rngs.standard_normal(KS) / np.sqrt(KS)
        g3 = ShardedGlm("poisson_log", ys, xs, KS, alpha=0.1, flags=flags, device=dev)
        g3.evaluate(beta_s)
        for _ in range(3):
            g3.evaluate()
        n_s = max(args.steps, 10)
        ms_s = device_time(g3.evaluate, n_s) / n_s
        (ms_s,) = max_over_ranks([ms_s])
        out_s = g3.out.cpu().numpy().copy()
        del g3, xs, ys
        mb.runtime.synchronize()
        _lib.lib().smc_trim_cache()
        host_barrier()
        if rank == 0:
            # the same problem alone on this GPU (51.2 GB), same run, same clocks
            xf = mb.MatrixCuda(NS, KS)
            xf.fill_synthetic(SEED, row0=0, kind=0)
            yf = mb.MatrixCuda(NS, 1, np.int32)
            yf.fill_synthetic(777, row0=0, kind=1, lo=0, hi=4)
            p1 = torch.as_tensor(beta_s).to(dev)
            o1 = torch.zeros(_lib.OUT_HEADER + KS, dtype=torch.float64, device=dev)
            one = lambda: cuda_local_eval("poisson_log", yf, xf, 0.1, None, p1, 0,  # noqa: E731
                                          flags, o1)
            for _ in range(3):
                one()
            k0, k1 = (torch.cuda.Event(enable_timing=True),
                      torch.cuda.Event(enable_timing=True))
            torch.cuda.synchronize()
            k0.record(stream)
            for _ in range(n_s):
                one()
            k1.record(stream)
            torch.cuda.synchronize()
            ms_1 = k0.elapsed_time(k1) / n_s
            o1h = o1.cpu().numpy()
            line["strong"] = {
                "workload": f"poisson_log_glm_lpmf N={NS} K={KS} in total, row-sharded over "
                            f"{world} GPUs (BASELINE.json configs[2]): broadcast + fused "
                            "kernel + NCCL all-reduce per evaluation",
                "ms_per_eval": ms_s, "evals_per_s": 1e3 / ms_s,
                "ms_per_eval_one_gpu": ms_1,
                "efficiency_vs_n1": ms_1 / (world * ms_s),
                "aggregate_GBps": NS * KS * 8 / (ms_s * 1e-3) / 1e9,
                "logp_rel_vs_one_gpu": float(abs(out_s[0] - o1h[0]) / abs(o1h[0])),
                "timing": "CUDA events on the launching stream, max over ranks; the "
                          "one-GPU figure is timed on rank 0 while the other ranks wait "
                          "on a host barrier"}
            del xf, yf
            mb.runtime.synchronize()
            _lib.lib().smc_trim_cache()
        host_barrier()
        # (c) the same through ONE process driving every GPU behind the C ABI
        if rank == 0:
            torch.cuda.empty_cache()
            line["single_process_sharded"] = single_process_sharded(world, args.steps,
                                                                    args.warmup)
        host_barrier()

    if rank == 0:
        if world == 1:
            line["e2e"]["cpp_drop_in"] = cpp_drop_in(N, K, args.steps, args.warmup)
            if not args.no_configs:
                import bench_configs
                line["configs"] = bench_configs.run_all(["1", "3", "4", "4b", "5a", "5b"])
            if not args.no_cpu_baseline:
                try:
                    line["cpu_baseline"] = cpu_baseline(N, K, args.cpu_rows)
                except Exception as e:  # the baseline must never sink the bench line
                    line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0,
                                            "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries ONE JSON line: whatever libraries print there while the run is set up
    # (NCCL's version banner, for one) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
