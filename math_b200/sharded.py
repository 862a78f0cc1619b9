"""Row-sharded GLM evaluation over the GPUs of one box.

Every family is a sum over independent rows and the parameters are small, so
rows are partitioned contiguously (``rows_g = [g*N/G, (g+1)*N/G)``), each shard
of x / y is uploaded once to its GPU, and one evaluation is

    broadcast(params)  ->  fused kernel per rank  ->  all_reduce(SUM) of the packed
    vector [logp, sum d, aux, nonfinite, aux2, -, -, -, d_beta[K], d_cuts[..]]

(K + O(1) doubles, latency-bound over NVLink / NVSwitch).  N-vector partials and
the N x K x-gradient stay sharded on their GPU: no exchange.  One process per
GPU; ``torch.distributed`` is the plumbing (NCCL on GPUs; the same code runs on
``gloo`` for the CPU tests of the host logic, with the local evaluator injected).

This is the B200 analogue of the reference's "scatter static data once, broadcast
parameters per call, reduce results" MPI pattern
(stan/math/prim/functor/mpi_parallel_call.hpp L332-392, L408-449); the GLMs
themselves never call a collective in the reference.
"""
import ctypes as C

import numpy as np

from ._lib import OUT_HEADER, check, lib

FAMILY = {"normal_id": 0, "bernoulli_logit": 1, "poisson_log": 2,
          "neg_binomial_2_log": 3, "ordered_logistic": 4, "binomial_logit": 5}


def shard_rows(n_rows, world_size, rank):
    """Contiguous row block of `rank`: [lo, hi)."""
    lo = (n_rows * rank) // world_size
    hi = (n_rows * (rank + 1)) // world_size
    return lo, hi


def packed_size(K, ncuts=0):
    return OUT_HEADER + K + ncuts


def _ptr(buf):
    """Device address of a torch tensor or of a MatrixCuda."""
    p = buf.data_ptr
    return C.c_void_p(p() if callable(p) else p)


def cuda_local_eval(family, y, x, alpha, aux, params_t, ncuts, flags, out_t,
                    d_alpha_vec=None, d_aux_vec=None, d_y_vec=None, d_x=None):
    """Asynchronous evaluation of this rank's shard on the current torch stream:
    parameters read from the device tensor `params_t`, packed result left in the
    device tensor `out_t`."""
    from .matrix_cuda import MatrixCuda

    def split(v):
        return (v.handle, 0.0) if isinstance(v, MatrixCuda) else (None, float(v))

    yh, ys = split(y)
    ah, a0 = split(alpha)
    xh, x0 = split(aux if aux is not None else 0.0)
    h = lambda m: m.handle if m is not None else None  # noqa: E731
    check(lib().smc_glm_eval_device(
        FAMILY[family], yh, ys, x.handle, ah, a0, xh, x0,
        _ptr(params_t), int(ncuts), int(flags), _ptr(out_t), h(d_alpha_vec),
        h(d_aux_vec), h(d_y_vec), h(d_x)))


class ShardedGlm:
    """One rank's view of a row-sharded GLM.

    ``x``/``y`` are this rank's shard (already on its device).  ``evaluate``
    takes the parameters on rank 0 (ignored elsewhere), broadcasts them,
    evaluates the local shard and all-reduces the packed result; every rank
    returns the same packed tensor."""

    def __init__(self, family, y, x, K, ncuts=0, alpha=0.0, aux=None, flags=0,
                 device="cuda", local_eval=cuda_local_eval, dist=None):
        import torch
        self.torch = torch
        if dist is None:
            import torch.distributed as dist
        self.dist = dist
        self.family, self.y, self.x = family, y, x
        self.K, self.ncuts, self.alpha, self.aux, self.flags = K, ncuts, alpha, aux, flags
        self.local_eval = local_eval
        self.params = torch.zeros(K + ncuts, dtype=torch.float64, device=device)
        self.out = torch.zeros(packed_size(K, ncuts), dtype=torch.float64, device=device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0

    def _adopt_stream(self):
        """The library launches on its own non-blocking stream unless told otherwise:
        the parameter copy / broadcast before the kernel and the all-reduce after it
        are only ordered with it when all of them share torch's current stream."""
        if self.local_eval in (cuda_local_eval, cuda_local_eval_categorical):
            s = self.torch.cuda.current_stream().cuda_stream
            if s != getattr(self, "_stream", None):
                check(lib().smc_set_stream(C.c_void_p(s)))
                self._stream = s

    def evaluate(self, params_host=None, **row_outputs):
        t = self.torch
        self._adopt_stream()
        if self.rank == 0 and params_host is not None:
            p = t.as_tensor(np.ascontiguousarray(params_host, dtype=np.float64))
            self.params.copy_(p, non_blocking=True)
        if self.world > 1:
            self.dist.broadcast(self.params, src=0)
        self.local_eval(self.family, self.y, self.x, self.alpha, self.aux,
                        self.params, self.ncuts, self.flags, self.out, **row_outputs)
        if self.world > 1:
            self.dist.all_reduce(self.out, op=self.dist.ReduceOp.SUM)
        return self.out

    def unpack(self, out_host):
        o = np.asarray(out_host, dtype=np.float64)
        res = {"logp": float(o[0]), "sum_d": float(o[1]), "aux": float(o[2]),
               "nonfinite": float(o[3]), "aux2": float(o[4]),
               "d_beta": o[OUT_HEADER:OUT_HEADER + self.K].copy()}
        if self.ncuts:
            res["d_cuts"] = o[OUT_HEADER + self.K:OUT_HEADER + self.K + self.ncuts].copy()
        return res


def cuda_local_eval_categorical(y, x, params_t, n_classes, flags, out_t, d_x=None):
    """This rank's shard of categorical_logit_glm_lpmf, asynchronous on the current
    torch stream: params_t = [beta (K x C, column-major), alpha (C)] on the device,
    out_t = [logp, nonfinite, d_alpha (C), d_beta (K x C)] left on the device."""
    from .matrix_cuda import MatrixCuda
    yh, ys = (y.handle, 0) if isinstance(y, MatrixCuda) else (None, int(y))
    check(lib().smc_categorical_logit_glm_device(
        yh, ys, x.handle, _ptr(params_t), int(n_classes), int(flags), _ptr(out_t),
        d_x.handle if d_x is not None else None))


class ShardedCategoricalGlm(ShardedGlm):
    """Row-sharded categorical_logit_glm_lpmf (K*C + C parameters
    broadcast, 2 + C + K*C doubles all-reduced -- 16,418 at K=512, C=32).  Same
    protocol as ShardedGlm; ``evaluate`` takes [beta.ravel(order="F"), alpha]."""

    def __init__(self, y, x, K, n_classes, flags=0, device="cuda",
                 local_eval=cuda_local_eval_categorical, dist=None):
        super().__init__("categorical_logit", y, x, K, flags=flags, device=device,
                         local_eval=local_eval, dist=dist)
        t = self.torch
        self.n_classes = n_classes
        self.params = t.zeros(K * n_classes + n_classes, dtype=t.float64, device=device)
        self.out = t.zeros(2 + n_classes + K * n_classes, dtype=t.float64, device=device)

    @staticmethod
    def pack_params(alpha, beta):
        return np.concatenate([np.asarray(beta, dtype=np.float64).ravel(order="F"),
                               np.asarray(alpha, dtype=np.float64).ravel()])

    def evaluate(self, params_host=None, **row_outputs):
        t = self.torch
        self._adopt_stream()
        if self.rank == 0 and params_host is not None:
            p = t.as_tensor(np.ascontiguousarray(params_host, dtype=np.float64))
            self.params.copy_(p, non_blocking=True)
        if self.world > 1:
            self.dist.broadcast(self.params, src=0)
        self.local_eval(self.y, self.x, self.params, self.n_classes, self.flags, self.out,
                        **row_outputs)
        if self.world > 1:
            self.dist.all_reduce(self.out, op=self.dist.ReduceOp.SUM)
        return self.out

    def unpack(self, out_host):
        o = np.asarray(out_host, dtype=np.float64)
        Cc, K = self.n_classes, self.K
        return {"logp": float(o[0]), "nonfinite": float(o[1]),
                "d_alpha": o[2:2 + Cc].copy(),
                "d_beta": o[2 + Cc:2 + Cc + K * Cc].reshape((K, Cc), order="F").copy()}
