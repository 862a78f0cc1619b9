"""Host-side mirror of the step either side of the fused GLMs.

``multiply`` / ``multiply_adjoint`` are the device matrix-vector product and its
reverse sweep (opencl/prim/multiply.hpp, opencl/rev/multiply.hpp); the
``*_lpmf`` functions are the un-fused densities on a device N-vector parameter
(prim/prob/bernoulli_logit_lpmf.hpp, poisson_log_lpmf.hpp,
neg_binomial_2_log_lpmf.hpp, ordered_logistic_lpmf.hpp, categorical_logit_lpmf.hpp)
for models that add
terms to ``x * beta`` before the likelihood.  Values and partials come back in
an ``LpmfResult`` (the C++ drop-in feeds them into the tape:
include/stan/math/cuda/prim/unfused_lpmf.hpp, rev/multiply.hpp).
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from ._lib import PROPTO, VAR_ALPHA, VAR_AUX, check, lib
from .glm import _beta, _dp, _h, _split
from .matrix_cuda import MatrixCuda


@dataclass
class LpmfResult:
    logp: float
    d_theta: Optional[MatrixCuda] = None  # N x 1, partial w.r.t. the vector parameter
    d_aux: Optional[object] = None        # d_phi (float) / d_cuts (ndarray)


def multiply(x, beta, alpha=0.0):
    """theta = x @ beta + alpha on the device (alpha: scalar or N x 1 MatrixCuda)."""
    av, a0 = _split(alpha, float, "alpha")
    b = _beta(beta, x.cols)
    theta = MatrixCuda(x.rows, 1, np.float64)
    check(lib().smc_linear_predictor(x.handle, _dp(b), _h(av), a0, theta.handle))
    return theta


def multiply_adjoint(x, v):
    """(x.T @ v, sum(v)) for an N x 1 device vector v: the reverse sweep of multiply."""
    g = np.zeros(x.cols)
    s = C.c_double()
    check(lib().smc_linear_predictor_adjoint(x.handle, v.handle, _dp(g), C.byref(s)))
    return g, s.value


def multiply_matrix(x, beta, alpha=None):
    """lin = x @ beta + alpha^T on the device for a K x C weight matrix (alpha: C
    values or None): the product in front of an un-fused categorical_logit_lpmf."""
    b = np.asfortranarray(np.asarray(beta, dtype=np.float64))
    if b.ndim != 2 or b.shape[0] != x.cols:
        raise ValueError(f"beta must be {x.cols} x C, got {b.shape}")
    a = None
    if alpha is not None:
        a = np.ascontiguousarray(np.asarray(alpha, dtype=np.float64).ravel())
        if a.size != b.shape[1]:
            raise ValueError("alpha must have one entry per column of beta")
    lin = MatrixCuda(x.rows, b.shape[1], np.float64)
    check(lib().smc_linear_predictor_matrix(x.handle, _dp(b), b.shape[1],
                                            _dp(a) if a is not None else None, lin.handle))
    return lin


def multiply_matrix_adjoint(x, adj):
    """(x.T @ adj, adj.sum(axis=0)) for an N x C device matrix adj: the reverse sweep
    of multiply_matrix (d_beta and d_alpha)."""
    g = np.zeros((x.cols, adj.cols), order="F")
    cs = np.zeros(adj.cols)
    check(lib().smc_linear_predictor_matrix_adjoint(x.handle, adj.handle, _dp(g), _dp(cs)))
    return g, cs


def indexing(z, idx):
    """out[i] = z[idx[i]] (0-based) on the device: opencl/kernel_generator/indexing.hpp."""
    zz = np.ascontiguousarray(np.atleast_1d(np.asarray(z, dtype=np.float64)).ravel())
    out = MatrixCuda(idx.size(), 1, np.float64)
    check(lib().smc_indexing(_dp(zz), zz.size, idx.handle, out.handle))
    return out


def indexing_rev(idx, res_adj, n_groups):
    """adj_z[g] = sum over rows with idx == g of res_adj (deterministic):
    opencl/indexing_rev.hpp L24-60."""
    g = np.zeros(int(n_groups))
    check(lib().smc_indexing_rev(idx.handle, res_adj.handle, int(n_groups), _dp(g)))
    return g


def vector_sum(v):
    s = C.c_double()
    check(lib().smc_vector_sum(v.handle, C.byref(s)))
    return s.value


def _flags(propto, theta_var, aux_var=False):
    return (PROPTO if propto else 0) | (VAR_ALPHA if theta_var else 0) \
        | (VAR_AUX if aux_var else 0)


def _theta_lpmf(fn, n, theta, propto, theta_var):
    nv, ns = _split(n, int, "n")
    flags = _flags(propto, theta_var)
    d = MatrixCuda(theta.size(), 1, np.float64) if theta_var else None
    logp = C.c_double()
    check(fn(_h(nv), ns, theta.handle, flags, C.byref(logp), _h(d)))
    return LpmfResult(logp.value, d)


def bernoulli_logit_lpmf(n, theta, propto=False, theta_var=True):
    """prim/prob/bernoulli_logit_lpmf.hpp L33-98."""
    return _theta_lpmf(lib().smc_bernoulli_logit_lpmf, n, theta, propto, theta_var)


def poisson_log_lpmf(n, alpha, propto=False, theta_var=True):
    """prim/prob/poisson_log_lpmf.hpp L27-100."""
    return _theta_lpmf(lib().smc_poisson_log_lpmf, n, alpha, propto, theta_var)


def neg_binomial_2_log_lpmf(n, eta, phi, propto=False, theta_var=True, phi_var=True):
    """prim/prob/neg_binomial_2_log_lpmf.hpp L24-134 (scalar phi)."""
    nv, ns = _split(n, int, "n")
    flags = _flags(propto, theta_var, phi_var)
    d = MatrixCuda(eta.size(), 1, np.float64) if theta_var else None
    logp, d_phi = C.c_double(), C.c_double()
    check(lib().smc_neg_binomial_2_log_lpmf(_h(nv), ns, eta.handle, None, float(phi),
                                            flags, C.byref(logp), _h(d),
                                            C.byref(d_phi), None))
    return LpmfResult(logp.value, d, d_phi.value if phi_var else None)


def ordered_logistic_lpmf(y, lam, cuts, propto=False, theta_var=True, cuts_var=True):
    """prim/prob/ordered_logistic_lpmf.hpp L72-214 (one cut-point vector)."""
    yv, ys = _split(y, int, "y")
    c = np.ascontiguousarray(np.atleast_1d(np.asarray(cuts, dtype=np.float64)).ravel())
    flags = _flags(propto, theta_var, cuts_var)
    d = MatrixCuda(lam.size(), 1, np.float64) if theta_var else None
    logp = C.c_double()
    d_cuts = np.zeros(c.size)
    check(lib().smc_ordered_logistic_lpmf(_h(yv), ys, lam.handle, _dp(c), c.size,
                                          flags, C.byref(logp), _h(d), _dp(d_cuts)))
    return LpmfResult(logp.value, d, d_cuts if cuts_var else None)


def ordered_logistic_lpmf_rows(y, lam, cuts, propto=False, theta_var=True, cuts_var=True):
    """prim/prob/ordered_logistic_lpmf.hpp L72-200 with one cut-point vector per outcome:
    ``cuts`` is a (C-1) x N (or (C-1) x 1) f64 MatrixCuda, column i the cut points of
    outcome i; ``d_aux`` is the device partial with the shape of ``cuts``."""
    yv, ys = _split(y, int, "y")
    flags = _flags(propto, theta_var, cuts_var)
    d = MatrixCuda(lam.size(), 1, np.float64) if theta_var else None
    d_cuts = MatrixCuda(cuts.rows, cuts.cols, np.float64) if cuts_var else None
    logp = C.c_double()
    check(lib().smc_ordered_logistic_lpmf_rows(_h(yv), ys, lam.handle, cuts.handle, flags,
                                               C.byref(logp), _h(d), _h(d_cuts)))
    return LpmfResult(logp.value, d, d_cuts)


def categorical_logit_lpmf(y, lin, propto=False, lin_var=True):
    """sum_i categorical_logit_lpmf(y_i | lin[i, :]) for an N x C device matrix of log
    odds (prim/prob/categorical_logit_lpmf.hpp L16-32, one row per outcome);
    ``d_theta`` is the N x C partial one-hot(y) - softmax(lin)."""
    yv, ys = _split(y, int, "y")
    flags = _flags(propto, lin_var)
    d = MatrixCuda(lin.rows, lin.cols, np.float64) if lin_var else None
    logp = C.c_double()
    check(lib().smc_categorical_logit_lpmf(_h(yv), ys, lin.handle, flags, C.byref(logp),
                                           _h(d)))
    return LpmfResult(logp.value, d)


def normal_lpdf(y, mu, sigma, propto=False, var=("mu", "sigma")):
    """prim/prob/normal_lpdf.hpp L41-104: y, mu scalars or N x 1 f64 MatrixCuda (at
    least one a vector), sigma a scalar.  `var` names the autodiff operands;
    returns (logp, d_y, d_mu, d_sigma) with device vectors for vector operands."""
    yv, ys = _split(y, float, "y")
    mv, ms = _split(mu, float, "mu")
    n = (mv if mv is not None else yv).size()
    flags = (PROPTO if propto else 0) | (VAR_ALPHA if "mu" in var else 0) \
        | (VAR_AUX if "sigma" in var else 0) | (32 if "y" in var else 0)
    d_yv = MatrixCuda(n, 1, np.float64) if ("y" in var and yv is not None) else None
    d_mv = MatrixCuda(n, 1, np.float64) if ("mu" in var and mv is not None) else None
    logp, d_y, d_mu, d_sigma = (C.c_double() for _ in range(4))
    check(lib().smc_normal_lpdf(_h(yv), ys, _h(mv), ms, float(sigma), flags,
                                C.byref(logp), _h(d_yv), C.byref(d_y), _h(d_mv),
                                C.byref(d_mu), C.byref(d_sigma)))
    return (logp.value,
            (d_yv if yv is not None else d_y.value) if "y" in var else None,
            (d_mv if mv is not None else d_mu.value) if "mu" in var else None,
            d_sigma.value if "sigma" in var else None)
