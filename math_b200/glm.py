"""Host-side mirror of the reference's GLM interface for the CUDA backend.

Same names and argument meaning as stan/math/prim/prob/<family>_glm_l{pdf,pmf}.hpp
and their device overloads in stan/math/opencl/prim/: ``y`` and ``x`` live on the
device (``MatrixCuda``), small parameters are host values, ``propto`` drops the
same constant terms, and errors map to the same exception kinds
(``ValueError`` <-> std::invalid_argument, ``DomainError`` <-> std::domain_error).

The reference returns a ``var`` whose partials are attached through
``make_partials_propagator``; here the value and the partials are returned in a
``GlmResult`` (the C++ drop-in that feeds them into the autodiff tape is
``include/stan/math/cuda/``).  ``var`` names which operands are autodiff
variables: any of "x", "alpha", "beta", "sigma"/"phi"/"cuts", "y".
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import (DX_FACTORED, PROPTO, VAR_ALPHA, VAR_AUX, VAR_BETA, VAR_X, VAR_Y,
                   check, lib)
from .matrix_cuda import MatrixCuda

_DP = C.POINTER(C.c_double)

# "x_factored": x is an autodiff variable and d_x comes back as the N x 1 factor d of
# d_x = d beta^T (SMC_DX_FACTORED; applied with MatrixCuda.rank1_update)
_FLAG = {"x": VAR_X, "x_factored": VAR_X | DX_FACTORED, "alpha": VAR_ALPHA, "beta": VAR_BETA, "sigma": VAR_AUX,
         "phi": VAR_AUX, "cuts": VAR_AUX, "y": VAR_Y}


@dataclass
class GlmResult:
    logp: float
    d_alpha: Optional[object] = None   # float (scalar alpha) or MatrixCuda (N x 1)
    d_beta: Optional[np.ndarray] = None
    d_aux: Optional[object] = None     # d_sigma / d_phi / d_cuts
    d_x: Optional[MatrixCuda] = None
    d_y: Optional[object] = None


def _flags(propto, var):
    f = PROPTO if propto else 0
    for v in var:
        f |= _FLAG[v]
    return f


def _dp(a):
    return a.ctypes.data_as(_DP)


def _h(m):
    return m.handle if m is not None else None


def _split(v, dtype, name):
    """scalar -> (None, value); MatrixCuda -> (matrix, 0)."""
    if isinstance(v, MatrixCuda):
        return v, 0
    if np.ndim(v) != 0:
        raise TypeError(f"{name}: pass a scalar or a MatrixCuda (use to_matrix_cuda)")
    return None, dtype(v)


def _beta(beta, K):
    b = np.ascontiguousarray(np.atleast_1d(np.asarray(beta, dtype=np.float64)).ravel())
    if b.size != K:
        raise ValueError(f"size of beta ({b.size}) does not match columns of x ({K})")
    return b


def _vec_out(cond, x):
    """An N x 1 output vector laid out (and sharded) like the rows of x."""
    return MatrixCuda.like(x, 1, np.float64) if cond else None


def _dx_out(flags, x):
    if flags & DX_FACTORED:
        return MatrixCuda.like(x, 1, np.float64)
    return MatrixCuda.like(x) if flags & VAR_X else None


def _glm4(fn, y, x, alpha, beta, propto, var):
    flags = _flags(propto, var)
    yv, ys = _split(y, int, "y")
    av, a0 = _split(alpha, float, "alpha")
    b = _beta(beta, x.cols)
    logp = C.c_double()
    d_alpha = C.c_double()
    d_beta = np.zeros(x.cols)
    d_av = _vec_out(av is not None and flags & VAR_ALPHA, x)
    d_x = _dx_out(flags, x)
    check(fn(_h(yv), ys, x.handle, _h(av), a0, _dp(b), flags, C.byref(logp),
             C.byref(d_alpha), _h(d_av), _dp(d_beta), _h(d_x)))
    return GlmResult(logp.value,
                     (d_av if av is not None else d_alpha.value)
                     if flags & VAR_ALPHA else None,
                     d_beta if flags & VAR_BETA else None, None, d_x)


def bernoulli_logit_glm_lpmf(y, x, alpha, beta, propto=False,
                             var=("alpha", "beta")):
    """prim/prob/bernoulli_logit_glm_lpmf.hpp L49-167."""
    return _glm4(lib().smc_bernoulli_logit_glm, y, x, alpha, beta, propto, var)


def poisson_log_glm_lpmf(y, x, alpha, beta, propto=False, var=("alpha", "beta")):
    """prim/prob/poisson_log_glm_lpmf.hpp L51-163."""
    return _glm4(lib().smc_poisson_log_glm, y, x, alpha, beta, propto, var)


def binomial_logit_glm_lpmf(n, trials, x, alpha, beta, propto=False,
                            var=("alpha", "beta")):
    """prim/prob/binomial_logit_glm_lpmf.hpp L54-160: `n` successes out of
    `trials` (the reference's N), each a scalar or an i32 MatrixCuda."""
    flags = _flags(propto, var)
    nv, ns = _split(n, int, "n")
    tv, ts = _split(trials, int, "trials")
    av, a0 = _split(alpha, float, "alpha")
    b = _beta(beta, x.cols)
    logp = C.c_double()
    d_alpha = C.c_double()
    d_beta = np.zeros(x.cols)
    d_av = _vec_out(av is not None and flags & VAR_ALPHA, x)
    d_x = _dx_out(flags, x)
    check(lib().smc_binomial_logit_glm(
        _h(nv), ns, _h(tv), ts, x.handle, _h(av), a0, _dp(b), flags,
        C.byref(logp), C.byref(d_alpha), _h(d_av), _dp(d_beta), _h(d_x)))
    return GlmResult(logp.value,
                     (d_av if av is not None else d_alpha.value)
                     if flags & VAR_ALPHA else None,
                     d_beta if flags & VAR_BETA else None, None, d_x)


def normal_id_glm_lpdf(y, x, alpha, beta, sigma, propto=False,
                       var=("alpha", "beta", "sigma")):
    """prim/prob/normal_id_glm_lpdf.hpp L54-216."""
    flags = _flags(propto, var)
    yv, ys = _split(y, float, "y")
    av, a0 = _split(alpha, float, "alpha")
    sv, s0 = _split(sigma, float, "sigma")
    b = _beta(beta, x.cols)
    N = x.rows
    logp, d_alpha, d_sigma, d_y = (C.c_double() for _ in range(4))
    d_beta = np.zeros(x.cols)
    d_av = _vec_out(av is not None and flags & VAR_ALPHA, x)
    d_sv = _vec_out(sv is not None and flags & VAR_AUX, x)
    d_yv = _vec_out(yv is not None and flags & VAR_Y, x)
    d_x = _dx_out(flags, x)
    check(lib().smc_normal_id_glm(
        _h(yv), ys, x.handle, _h(av), a0, _dp(b), _h(sv), s0, flags,
        C.byref(logp), C.byref(d_alpha), _h(d_av), _dp(d_beta), C.byref(d_sigma),
        _h(d_sv), C.byref(d_y), _h(d_yv), _h(d_x)))
    return GlmResult(
        logp.value,
        (d_av if av is not None else d_alpha.value) if flags & VAR_ALPHA else None,
        d_beta if flags & VAR_BETA else None,
        (d_sv if sv is not None else d_sigma.value) if flags & VAR_AUX else None,
        d_x, (d_yv if yv is not None else d_y.value) if flags & VAR_Y else None)


def neg_binomial_2_log_glm_lpmf(y, x, alpha, beta, phi, propto=False,
                                var=("alpha", "beta", "phi")):
    """prim/prob/neg_binomial_2_log_glm_lpmf.hpp L64-248."""
    flags = _flags(propto, var)
    yv, ys = _split(y, int, "y")
    av, a0 = _split(alpha, float, "alpha")
    pv, p0 = _split(phi, float, "phi")
    b = _beta(beta, x.cols)
    N = x.rows
    logp, d_alpha, d_phi = (C.c_double() for _ in range(3))
    d_beta = np.zeros(x.cols)
    d_av = _vec_out(av is not None and flags & VAR_ALPHA, x)
    d_pv = _vec_out(pv is not None and flags & VAR_AUX, x)
    d_x = _dx_out(flags, x)
    check(lib().smc_neg_binomial_2_log_glm(
        _h(yv), ys, x.handle, _h(av), a0, _dp(b), _h(pv), p0, flags,
        C.byref(logp), C.byref(d_alpha), _h(d_av), _dp(d_beta), C.byref(d_phi),
        _h(d_pv), _h(d_x)))
    return GlmResult(
        logp.value,
        (d_av if av is not None else d_alpha.value) if flags & VAR_ALPHA else None,
        d_beta if flags & VAR_BETA else None,
        (d_pv if pv is not None else d_phi.value) if flags & VAR_AUX else None,
        d_x)


def ordered_logistic_glm_lpmf(y, x, beta, cuts, propto=False,
                              var=("beta", "cuts")):
    """prim/prob/ordered_logistic_glm_lpmf.hpp L46-210."""
    flags = _flags(propto, var)
    yv, ys = _split(y, int, "y")
    b = _beta(beta, x.cols)
    c = np.ascontiguousarray(np.atleast_1d(np.asarray(cuts, dtype=np.float64)).ravel())
    logp = C.c_double()
    d_beta = np.zeros(x.cols)
    d_cuts = np.zeros(c.size)
    d_x = _dx_out(flags, x)
    check(lib().smc_ordered_logistic_glm(
        _h(yv), ys, x.handle, _dp(b), _dp(c), c.size, flags, C.byref(logp),
        _dp(d_beta), _dp(d_cuts), _h(d_x)))
    return GlmResult(logp.value, None, d_beta if flags & VAR_BETA else None,
                     d_cuts if flags & VAR_AUX else None, d_x)


def categorical_logit_glm_lpmf(y, x, alpha, beta, propto=False,
                               var=("alpha", "beta")):
    """prim/prob/categorical_logit_glm_lpmf.hpp L43-195 (beta is K x C)."""
    flags = _flags(propto, var)
    yv, ys = _split(y, int, "y")
    beta = np.asfortranarray(np.asarray(beta, dtype=np.float64))
    if beta.ndim != 2 or beta.shape[0] != x.cols:
        raise ValueError("x.cols() and beta.rows() must match in size")
    Cc = beta.shape[1]
    a = np.ascontiguousarray(np.atleast_1d(np.asarray(alpha, dtype=np.float64)).ravel())
    if a.size != Cc:
        raise ValueError("size of alpha does not match the number of classes")
    logp = C.c_double()
    d_alpha = np.zeros(Cc)
    d_beta = np.zeros((x.cols, Cc), order="F")
    d_x = _dx_out(flags, x)
    check(lib().smc_categorical_logit_glm(
        _h(yv), ys, x.handle, _dp(a), _dp(beta), Cc, flags, C.byref(logp),
        _dp(d_alpha), _dp(d_beta), _h(d_x)))
    return GlmResult(logp.value, d_alpha if flags & VAR_ALPHA else None,
                     d_beta if flags & VAR_BETA else None, None, d_x)
