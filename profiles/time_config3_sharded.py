#!/usr/bin/env python
"""BASELINE.json config 3: poisson_log_glm_lpmf, N = 1e8 rows x K = 64, row-sharded
over the ranks of one box (STRONG scaling: the global N is fixed), parameters
broadcast and the packed K + 8 partials all-reduced over NCCL every evaluation.
Launch with torchrun (one rank per GPU); rank 0 prints one JSON line.  Timed on
the device (CUDA events on the launch stream, max over ranks)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import math_b200 as mb
    from math_b200 import _lib
    from math_b200.sharded import ShardedGlm, shard_rows

    N, K, steps, warmup = 100_000_000, 64, 30, 5
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    mb.runtime.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    mb.runtime.set_stream(stream.cuda_stream)
    lo, hi = shard_rows(N, world, rank)
    x = mb.MatrixCuda(hi - lo, K)
    x.fill_synthetic(12345, row0=lo, kind=0)
    y = mb.MatrixCuda(hi - lo, 1, np.int32)
    y.fill_synthetic(777, row0=lo, kind=1, lo=0, hi=4)
    beta = np.random.default_rng(12345).standard_normal(K) / np.sqrt(K)
    glm = ShardedGlm("poisson_log", y, x, K, alpha=0.1,
                     flags=_lib.VAR_ALPHA | _lib.VAR_BETA, device=f"cuda:{local}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    glm.evaluate(beta)
    for _ in range(warmup):
        glm.evaluate(beta)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        glm.evaluate(beta)  # host parameters -> rank 0 -> broadcast, every evaluation
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = glm.out.cpu().numpy()
    if rank == 0:
        ms = float(t.cpu()[0])
        print(json.dumps({"config": "3", "n_gpus": world, "N_global": N, "K": K,
                          "ms_per_eval": ms, "evals_per_s": 1e3 / ms,
                          "GBps_aggregate": N * K * 8 / ms / 1e6,
                          "logp": float(out[0]), "nonfinite": float(out[3])}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
