"""CPU, world_size 2 over gloo: the host logic of the row-sharded evaluation
(partitioning, parameter broadcast, packed all-reduce, unpacking).  The local
evaluator is injected: here it is the CPU oracle writing the same packed layout
the CUDA kernel writes, so the collective plumbing is exercised without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from math_b200 import _lib
from math_b200.sharded import ShardedCategoricalGlm, ShardedGlm, packed_size, shard_rows
from oracle import pyoracle as po
from tests.util import assert_grad, assert_logp, make_inputs


def test_shard_rows_partition():
    for n in (0, 1, 7, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            blocks = [shard_rows(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _oracle_local_eval(family, y, x, alpha, aux, params_t, ncuts, flags, out_t, **_):
    K = x.shape[1]
    p = params_t.numpy()
    beta = p[:K]
    out = np.zeros(out_t.numel())
    if family == "poisson_log":
        r = po.poisson_log_glm(y, x, alpha, beta, flags)
        out[0], out[1] = r["logp"], r["d_alpha"][0]
        out[_lib.OUT_HEADER:_lib.OUT_HEADER + K] = r["d_beta"]
    elif family == "ordered_logistic":
        r = po.ordered_logistic_glm(y, x, beta, p[K:K + ncuts], flags)
        out[0] = r["logp"]
        out[_lib.OUT_HEADER:_lib.OUT_HEADER + K] = r["d_beta"]
        out[_lib.OUT_HEADER + K:_lib.OUT_HEADER + K + ncuts] = r["d_cuts"]
    else:
        raise AssertionError(family)
    out_t.copy_(torch.from_numpy(out))


def _oracle_local_eval_categorical(y, x, params_t, n_classes, flags, out_t, **_):
    K, Cc = x.shape[1], n_classes
    p = params_t.numpy()
    beta = p[:K * Cc].reshape((K, Cc), order="F")
    alpha = p[K * Cc:]
    r = po.categorical_logit_glm(y, x, alpha, beta, flags)
    out = np.concatenate([[r["logp"], 0.0], r["d_alpha"], r["d_beta"].ravel(order="F")])
    out_t.copy_(torch.from_numpy(out))


def _worker(rank, world, port, family, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, K = 1001, 13
        if family == "categorical_logit":
            d = make_inputs("categorical", N, K, seed=5, C=6)
            lo, hi = shard_rows(N, world, rank)
            glm = ShardedCategoricalGlm(d["y"][lo:hi], d["x"][lo:hi], K, 6,
                                        flags=po.VAR_ALPHA | po.VAR_BETA, device="cpu",
                                        local_eval=_oracle_local_eval_categorical)
            params = ShardedCategoricalGlm.pack_params(d["alpha"], d["beta"])
            res = glm.unpack(glm.evaluate(params if rank == 0 else None).numpy())
            q.put(res if rank == 0 else {"logp": res["logp"]})
            return
        if family == "poisson_log":
            d = make_inputs("poisson", N, K, seed=3)
            params, ncuts = d["beta"], 0
            flags = po.VAR_ALPHA | po.VAR_BETA
        else:
            d = make_inputs("ordered", N, K, seed=4, C=6)
            params, ncuts = np.concatenate([d["beta"], d["cuts"]]), 5
            flags = po.VAR_BETA | po.VAR_AUX
        lo, hi = shard_rows(N, world, rank)
        glm = ShardedGlm(family, d["y"][lo:hi], d["x"][lo:hi], K, ncuts=ncuts,
                         alpha=0.1, flags=flags, device="cpu",
                         local_eval=_oracle_local_eval)
        # only rank 0 supplies the parameters; the others get them by broadcast
        out = glm.evaluate(params if rank == 0 else None)
        res = glm.unpack(out.numpy())
        if rank == 0:
            q.put(res)
        else:
            q.put({"logp": res["logp"]})
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("family", ["poisson_log", "ordered_logistic", "categorical_logit"])
def test_two_ranks_match_single(family):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, family, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = next(g for g in got if "d_beta" in g)
    other = next(g for g in got if "d_beta" not in g)
    assert full["logp"] == other["logp"]  # every rank holds the reduced result
    N, K = 1001, 13
    if family == "categorical_logit":
        d = make_inputs("categorical", N, K, seed=5, C=6)
        ref = po.categorical_logit_glm(d["y"], d["x"], d["alpha"], d["beta"])
        assert_grad(full["d_alpha"], ref["d_alpha"], "d_alpha")
        assert full["d_beta"].shape == (K, 6)
    elif family == "poisson_log":
        d = make_inputs("poisson", N, K, seed=3)
        ref = po.poisson_log_glm(d["y"], d["x"], 0.1, d["beta"])
    else:
        d = make_inputs("ordered", N, K, seed=4, C=6)
        ref = po.ordered_logistic_glm(d["y"], d["x"], d["beta"], d["cuts"])
        assert_grad(full["d_cuts"], ref["d_cuts"], "d_cuts")
    assert_logp(full["logp"], ref["logp"])
    assert_grad(np.ravel(full["d_beta"], order="F"), np.ravel(ref["d_beta"], order="F"),
                "d_beta")
    assert packed_size(K, 5) == _lib.OUT_HEADER + K + 5
