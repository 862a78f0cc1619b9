"""One rank of the NCCL row-sharded parity check (launched by torchrun from
tests/test_multi_gpu.py, or by hand:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/multi_gpu_worker.py).

Every rank uploads its contiguous row shard of ONE global problem, the sharded
driver broadcasts the parameters from rank 0 and all-reduces the packed result;
rank 0 also evaluates the whole problem on its own GPU and checks
  * multi-GPU == single-GPU to 1e-12 relative (summation order only), and
  * multi-GPU == the CPU oracle within BASELINE.json's tolerances.
"""
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
    from math_b200.sharded import ShardedCategoricalGlm, ShardedGlm, shard_rows
    from oracle import pyoracle as po
    from tests.util import assert_grad, assert_logp, make_inputs

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    mb.runtime.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    mb.runtime.set_stream(stream.cuda_stream)
    report = {}
    N = 20_011
    cases = [("bernoulli_logit", "bernoulli", 256, {}), ("poisson_log", "poisson", 64, {}),
             ("normal_id", "normal", 100, {}), ("neg_binomial_2_log", "neg_binomial", 128, {}),
             ("ordered_logistic", "ordered", 64, {"C": 9})]
    for family, short, K, extra in cases:
        d = make_inputs(short, N, K, seed=11, **extra)
        lo, hi = shard_rows(N, world, rank)
        x = mb.to_matrix_cuda(np.asfortranarray(d["x"][lo:hi]))
        y = mb.to_matrix_cuda(np.ascontiguousarray(d["y"][lo:hi]))
        ncuts = len(d["cuts"]) if "cuts" in d else 0
        params = np.concatenate([d["beta"], d["cuts"]]) if ncuts else d["beta"]
        aux = d.get("sigma", d.get("phi"))
        flags = _lib.VAR_BETA | _lib.VAR_AUX | (0 if short == "ordered" else _lib.VAR_ALPHA)
        glm = ShardedGlm(family, y, x, K, ncuts=ncuts, alpha=d.get("alpha", 0.0), aux=aux,
                         flags=flags, device=f"cuda:{local}")
        out = glm.evaluate(params if rank == 0 else None)
        torch.cuda.synchronize()
        res = glm.unpack(out.cpu().numpy())
        if rank == 0:
            xf = mb.to_matrix_cuda(np.asfortranarray(d["x"]))
            yf = mb.to_matrix_cuda(np.ascontiguousarray(d["y"]))
            single = ShardedGlm(family, yf, xf, K, ncuts=ncuts, alpha=d.get("alpha", 0.0),
                                aux=aux, flags=flags, device=f"cuda:{local}", dist=_NoDist())
            s = single.unpack(single.evaluate(params).cpu().numpy())
            rel = abs(res["logp"] - s["logp"]) / abs(s["logp"])
            gscale = np.abs(s["d_beta"]).max()
            grel = np.abs(res["d_beta"] - s["d_beta"]).max() / gscale
            assert rel < 1e-12 and grel < 1e-12, (family, rel, grel)
            fn = getattr(po, {"bernoulli": "bernoulli_logit_glm", "poisson": "poisson_log_glm",
                              "normal": "normal_id_glm", "neg_binomial": "neg_binomial_2_log_glm",
                              "ordered": "ordered_logistic_glm"}[short])
            if short == "ordered":
                o = fn(d["y"], d["x"], d["beta"], d["cuts"])
                assert_grad(res["d_cuts"], o["d_cuts"], "d_cuts")
            elif aux is not None:
                o = fn(d["y"], d["x"], d["alpha"], d["beta"], aux)
            else:
                o = fn(d["y"], d["x"], d["alpha"], d["beta"])
            assert_logp(res["logp"], o["logp"])
            assert_grad(res["d_beta"], o["d_beta"], "d_beta")
            report[family] = {"rel_logp_vs_single": rel, "rel_dbeta_vs_single": grel}
    # the GEMM-shaped family: K*C + C parameters broadcast, 2 + C + K*C doubles reduced
    for K, Cc in ((96, 32), (37, 5)):
        d = make_inputs("categorical", N, K, seed=12, C=Cc)
        lo, hi = shard_rows(N, world, rank)
        x = mb.to_matrix_cuda(np.asfortranarray(d["x"][lo:hi]))
        y = mb.to_matrix_cuda(np.ascontiguousarray(d["y"][lo:hi]))
        flags = _lib.VAR_ALPHA | _lib.VAR_BETA
        glm = ShardedCategoricalGlm(y, x, K, Cc, flags=flags, device=f"cuda:{local}")
        params = ShardedCategoricalGlm.pack_params(d["alpha"], d["beta"])
        out = glm.evaluate(params if rank == 0 else None)
        torch.cuda.synchronize()
        res = glm.unpack(out.cpu().numpy())
        if rank == 0:
            s = mb.categorical_logit_glm_lpmf(mb.to_matrix_cuda(np.ascontiguousarray(d["y"])),
                                              mb.to_matrix_cuda(np.asfortranarray(d["x"])),
                                              d["alpha"], d["beta"])
            rel = abs(res["logp"] - s.logp) / abs(s.logp)
            grel = np.abs(res["d_beta"] - s.d_beta).max() / np.abs(s.d_beta).max()
            arel = np.abs(res["d_alpha"] - s.d_alpha).max() / np.abs(s.d_alpha).max()
            assert rel < 1e-12 and grel < 1e-12 and arel < 1e-12, (K, Cc, rel, grel, arel)
            assert res["nonfinite"] == 0.0
            o = po.categorical_logit_glm(d["y"], d["x"], d["alpha"], d["beta"])
            assert_logp(res["logp"], o["logp"])
            assert_grad(res["d_alpha"], o["d_alpha"], "d_alpha")
            assert_grad(res["d_beta"].ravel(order="F"), o["d_beta"].ravel(order="F"), "d_beta")
            report[f"categorical_logit_K{K}_C{Cc}"] = {"rel_logp_vs_single": rel,
                                                       "rel_dbeta_vs_single": grel}
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK " + json.dumps({"world": world, "cases": report}), flush=True)
    dist.destroy_process_group()


class _NoDist:
    """Stands in for torch.distributed on the single-GPU comparison run."""

    @staticmethod
    def is_initialized():
        return False


if __name__ == "__main__":
    main()
