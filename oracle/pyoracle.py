"""ctypes front-end to the CHECKERS (test infrastructure, never the product).

Two back-ends behind the same call shape:

* ``impl="oracle"`` -> ``oracle/libglm_oracle.so`` (plain-C restatement,
  ``oracle/glm_oracle.c``; always available, built by ``oracle/Makefile``).
* ``impl="ref"``    -> ``oracle/_ref/libstan_ref.so`` (the unmodified reference
  compiled from /root/reference by ``oracle/ref_driver.cpp``; present only when
  it was built in the CPU container -- it then travels to the GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

PROPTO, VAR_X, VAR_ALPHA, VAR_BETA, VAR_AUX, VAR_Y = 1, 2, 4, 8, 16, 32
ALL_PARAMS = VAR_ALPHA | VAR_BETA | VAR_AUX

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_L = C.c_long


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


_oracle = None
_ref = {}


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(HERE, "libglm_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        _oracle = C.CDLL(path)
        for n in ("oracle_lgamma", "oracle_digamma", "oracle_log1p_exp",
                  "oracle_log1m_exp", "oracle_log_inv_logit",
                  "oracle_log1m_inv_logit"):
            getattr(_oracle, n).restype = C.c_double
            getattr(_oracle, n).argtypes = [C.c_double]
        _oracle.oracle_binomial_coefficient_log.restype = C.c_double
        _oracle.oracle_binomial_coefficient_log.argtypes = [C.c_double, C.c_double]
    return _oracle


def synthetic(seed, row0, nrows, ncols, kind=0, scale=1.0, lo=0, hi=1):
    """The inputs smc_matrix_fill_synthetic puts on the GPU, generated on the host by
    the C oracle (column-major; all host cores): bit-identical by construction."""
    lib = oracle_lib()
    if kind == 0:
        out = np.empty((nrows, ncols), dtype=np.float64, order="F")
        lib.oracle_synthetic_f64(out.ctypes.data_as(C.c_void_p), C.c_long(max(nrows, 1)),
                                 C.c_long(nrows), C.c_long(ncols), C.c_uint64(seed),
                                 C.c_long(row0), C.c_double(scale))
    else:
        out = np.empty((nrows, ncols), dtype=np.int32, order="F")
        lib.oracle_synthetic_i32(out.ctypes.data_as(C.c_void_p), C.c_long(max(nrows, 1)),
                                 C.c_long(nrows), C.c_long(ncols), C.c_uint64(seed),
                                 C.c_long(row0), C.c_int(lo), C.c_int(hi))
    return out


def ref_available(mt=False):
    return os.path.exists(os.path.join(
        HERE, "_ref", "libstan_ref_mt.so" if mt else "libstan_ref.so"))


def ref_lib(mt=False):
    if mt not in _ref:
        name = "libstan_ref_mt.so" if mt else "libstan_ref.so"
        lib = C.CDLL(os.path.join(HERE, "_ref", name))
        for n in ("ref_digamma", "ref_lgamma", "ref_log1p_exp", "ref_log1m_exp",
                  "ref_log_inv_logit", "ref_log1m_inv_logit"):
            getattr(lib, n).restype = C.c_double
            getattr(lib, n).argtypes = [C.c_double]
        lib.ref_binomial_coefficient_log.restype = C.c_double
        lib.ref_binomial_coefficient_log.argtypes = [C.c_double, C.c_double]
        lib.ref_time_glm.restype = C.c_double
        lib.ref_time_glm.argtypes = [C.c_int, _L, _L, C.c_void_p, _dp,
                                     C.c_double, _dp, C.c_double, C.c_int, _dp,
                                     _dp]
        if mt:
            lib.ref_time_glm_reduce_sum.restype = C.c_double
            lib.ref_time_glm_reduce_sum.argtypes = [
                C.c_int, _L, _L, C.c_void_p, _dp, C.c_double, _dp, C.c_double,
                C.c_int, _L, C.c_int, _dp, _dp]
        _ref[mt] = lib
    return _ref[mt]


def _prep_x(x):
    x = np.asfortranarray(np.asarray(x, dtype=np.float64))
    if x.ndim != 2:
        raise ValueError("x must be N x K")
    return x


def _vec(a, dtype):
    a = np.ascontiguousarray(np.atleast_1d(np.asarray(a, dtype=dtype)).ravel())
    return a


def _out(n):
    return np.full(max(int(n), 1), np.nan)[: int(n)].copy() if n else np.zeros(0)


def _finish(rc, logp, **grads):
    res = {"rc": int(rc), "logp": float(logp[0])}
    res.update(grads)
    return res


def _ref_flags(flags):
    """The reference driver makes every real parameter a var; x (and normal's
    y) follow VAR_X.  Callers must pass flags consistent with that."""
    want = ALL_PARAMS
    assert flags & want == want, "ref back-end: alpha/beta/aux are always var"
    return int(bool(flags & PROPTO)), int(bool(flags & VAR_X))


def bernoulli_logit_glm(y, x, alpha, beta, flags=VAR_ALPHA | VAR_BETA,
                        impl="oracle", poisson=False):
    x = _prep_x(x)
    N, K = x.shape
    y = _vec(y, np.int32)
    alpha = _vec(alpha, np.float64)
    beta = _vec(beta, np.float64)
    assert beta.size == K
    logp = np.zeros(1)
    d_alpha = np.zeros(alpha.size)
    d_beta = np.zeros(K)
    d_x = np.zeros((N, K), order="F") if flags & VAR_X else None
    if impl == "oracle":
        lib = oracle_lib()
        fn = lib.oracle_poisson_log_glm if poisson else lib.oracle_bernoulli_logit_glm
        rc = fn(_L(N), _L(K), _i(y), _L(y.size), _d(x), _L(max(N, 1)), _d(alpha),
                _L(alpha.size), _d(beta), C.c_uint(flags), _d(logp), _d(d_alpha),
                _d(d_beta), _d(d_x))
    else:
        lib = ref_lib()
        fn = lib.ref_poisson_log_glm if poisson else lib.ref_bernoulli_logit_glm
        propto, dv = _ref_flags(flags | VAR_AUX)
        rc = fn(_L(N), _L(K), _i(y), _L(y.size), _d(x), _d(alpha),
                _L(alpha.size), _d(beta), propto, dv, _d(logp), _d(d_alpha),
                _d(d_beta), _d(d_x))
    return _finish(rc, logp, d_alpha=d_alpha, d_beta=d_beta, d_x=d_x)


def poisson_log_glm(y, x, alpha, beta, flags=VAR_ALPHA | VAR_BETA, impl="oracle"):
    return bernoulli_logit_glm(y, x, alpha, beta, flags, impl, poisson=True)


def binomial_logit_glm(n, trials, x, alpha, beta, flags=VAR_ALPHA | VAR_BETA,
                       impl="oracle"):
    """binomial_logit_glm_lpmf(n | N, x, alpha, beta): `n` successes, `trials`
    the population sizes N (each a scalar or an N-vector)."""
    x = _prep_x(x)
    N, K = x.shape
    n = _vec(n, np.int32)
    trials = _vec(trials, np.int32)
    alpha = _vec(alpha, np.float64)
    beta = _vec(beta, np.float64)
    assert beta.size == K
    logp = np.zeros(1)
    d_alpha = np.zeros(alpha.size)
    d_beta = np.zeros(K)
    d_x = np.zeros((N, K), order="F") if flags & VAR_X else None
    if impl == "oracle":
        rc = oracle_lib().oracle_binomial_logit_glm(
            _L(N), _L(K), _i(n), _L(n.size), _i(trials), _L(trials.size), _d(x),
            _L(max(N, 1)), _d(alpha), _L(alpha.size), _d(beta), C.c_uint(flags),
            _d(logp), _d(d_alpha), _d(d_beta), _d(d_x))
    else:
        propto, dv = _ref_flags(flags | VAR_AUX)
        rc = ref_lib().ref_binomial_logit_glm(
            _L(N), _L(K), _i(n), _L(n.size), _i(trials), _L(trials.size), _d(x),
            _d(alpha), _L(alpha.size), _d(beta), propto, dv, _d(logp),
            _d(d_alpha), _d(d_beta), _d(d_x))
    return _finish(rc, logp, d_alpha=d_alpha, d_beta=d_beta, d_x=d_x)


def normal_id_glm(y, x, alpha, beta, sigma, flags=ALL_PARAMS, impl="oracle"):
    x = _prep_x(x)
    N, K = x.shape
    y = _vec(y, np.float64)
    alpha = _vec(alpha, np.float64)
    beta = _vec(beta, np.float64)
    sigma = _vec(sigma, np.float64)
    logp = np.zeros(1)
    d_alpha = np.zeros(alpha.size)
    d_beta = np.zeros(K)
    d_sigma = np.zeros(sigma.size)
    d_x = np.zeros((N, K), order="F") if flags & VAR_X else None
    d_y = np.zeros(y.size) if flags & VAR_Y else None
    if impl == "oracle":
        rc = oracle_lib().oracle_normal_id_glm(
            _L(N), _L(K), _d(y), _L(y.size), _d(x), _L(max(N, 1)), _d(alpha),
            _L(alpha.size), _d(beta), _d(sigma), _L(sigma.size), C.c_uint(flags),
            _d(logp), _d(d_alpha), _d(d_beta), _d(d_sigma), _d(d_x), _d(d_y))
    else:
        assert bool(flags & VAR_X) == bool(flags & VAR_Y), \
            "ref back-end ties x var and y var together"
        propto, dv = _ref_flags(flags)
        rc = ref_lib().ref_normal_id_glm(
            _L(N), _L(K), _d(y), _L(y.size), _d(x), _d(alpha), _L(alpha.size),
            _d(beta), _d(sigma), _L(sigma.size), propto, dv, _d(logp),
            _d(d_alpha), _d(d_beta), _d(d_sigma), _d(d_x), _d(d_y))
    return _finish(rc, logp, d_alpha=d_alpha, d_beta=d_beta, d_sigma=d_sigma,
                   d_x=d_x, d_y=d_y)


def neg_binomial_2_log_glm(y, x, alpha, beta, phi, flags=ALL_PARAMS,
                           impl="oracle"):
    x = _prep_x(x)
    N, K = x.shape
    y = _vec(y, np.int32)
    alpha = _vec(alpha, np.float64)
    beta = _vec(beta, np.float64)
    phi = _vec(phi, np.float64)
    logp = np.zeros(1)
    d_alpha = np.zeros(alpha.size)
    d_beta = np.zeros(K)
    d_phi = np.zeros(phi.size)
    d_x = np.zeros((N, K), order="F") if flags & VAR_X else None
    if impl == "oracle":
        rc = oracle_lib().oracle_neg_binomial_2_log_glm(
            _L(N), _L(K), _i(y), _L(y.size), _d(x), _L(max(N, 1)), _d(alpha),
            _L(alpha.size), _d(beta), _d(phi), _L(phi.size), C.c_uint(flags),
            _d(logp), _d(d_alpha), _d(d_beta), _d(d_phi), _d(d_x))
    else:
        propto, dv = _ref_flags(flags)
        rc = ref_lib().ref_neg_binomial_2_log_glm(
            _L(N), _L(K), _i(y), _L(y.size), _d(x), _d(alpha), _L(alpha.size),
            _d(beta), _d(phi), _L(phi.size), propto, dv, _d(logp), _d(d_alpha),
            _d(d_beta), _d(d_phi), _d(d_x))
    return _finish(rc, logp, d_alpha=d_alpha, d_beta=d_beta, d_phi=d_phi, d_x=d_x)


def ordered_logistic_glm(y, x, beta, cuts, flags=VAR_BETA | VAR_AUX,
                         impl="oracle"):
    x = _prep_x(x)
    N, K = x.shape
    y = _vec(y, np.int32)
    beta = _vec(beta, np.float64)
    cuts = _vec(cuts, np.float64)
    logp = np.zeros(1)
    d_beta = np.zeros(K)
    d_cuts = np.zeros(cuts.size)
    d_x = np.zeros((N, K), order="F") if flags & VAR_X else None
    if impl == "oracle":
        rc = oracle_lib().oracle_ordered_logistic_glm(
            _L(N), _L(K), _i(y), _L(y.size), _d(x), _L(max(N, 1)), _d(beta),
            _d(cuts), _L(cuts.size), C.c_uint(flags), _d(logp), _d(d_beta),
            _d(d_cuts), _d(d_x))
    else:
        propto, dv = _ref_flags(flags | VAR_ALPHA)
        rc = ref_lib().ref_ordered_logistic_glm(
            _L(N), _L(K), _i(y), _L(y.size), _d(x), _d(beta), _d(cuts),
            _L(cuts.size), propto, dv, _d(logp), _d(d_beta), _d(d_cuts), _d(d_x))
    return _finish(rc, logp, d_beta=d_beta, d_cuts=d_cuts, d_x=d_x)


def categorical_logit_glm(y, x, alpha, beta, flags=VAR_ALPHA | VAR_BETA,
                          impl="oracle"):
    x = _prep_x(x)
    N, K = x.shape
    beta = np.asfortranarray(np.asarray(beta, dtype=np.float64))
    assert beta.ndim == 2 and beta.shape[0] == K
    Cc = beta.shape[1]
    y = _vec(y, np.int32)
    alpha = _vec(alpha, np.float64)
    assert alpha.size == Cc
    logp = np.zeros(1)
    d_alpha = np.zeros(Cc)
    d_beta = np.zeros((K, Cc), order="F")
    d_x = np.zeros((N, K), order="F") if flags & VAR_X else None
    if impl == "oracle":
        rc = oracle_lib().oracle_categorical_logit_glm(
            _L(N), _L(K), _L(Cc), _i(y), _L(y.size), _d(x), _L(max(N, 1)),
            _d(alpha), _d(beta), C.c_uint(flags), _d(logp), _d(d_alpha),
            _d(d_beta), _d(d_x))
    else:
        propto, dv = _ref_flags(flags | VAR_AUX)
        rc = ref_lib().ref_categorical_logit_glm(
            _L(N), _L(K), _L(Cc), _i(y), _L(y.size), _d(x), _d(alpha), _d(beta),
            propto, dv, _d(logp), _d(d_alpha), _d(d_beta), _d(d_x))
    return _finish(rc, logp, d_alpha=d_alpha, d_beta=d_beta, d_x=d_x)


def categorical_logit_lpmf(y, lin, flags=VAR_ALPHA, impl="oracle"):
    """sum_i categorical_logit_lpmf(y_i | lin[i, :]) for an N x C matrix of log odds
    (prim/prob/categorical_logit_lpmf.hpp L16-32 per row); d_lin is N x C."""
    lin = _prep_x(lin)
    N, Cc = lin.shape
    y = _vec(y, np.int32)
    logp = np.zeros(1)
    d_lin = np.zeros((N, Cc), order="F")
    if impl == "oracle":
        rc = oracle_lib().oracle_categorical_logit_lpmf(
            _L(N), _L(Cc), _i(y), _L(y.size), _d(lin), _L(max(N, 1)), C.c_uint(flags),
            _d(logp), _d(d_lin))
    else:
        rc = ref_lib().ref_categorical_logit_lpmf(
            _L(N), _L(Cc), _i(y), _L(y.size), _d(lin), int(bool(flags & PROPTO)),
            _d(logp), _d(d_lin))
    return _finish(rc, logp, d_lin=d_lin)


FAMILY_ID = {"normal": 0, "bernoulli": 1, "poisson": 2, "neg_binomial": 3}


def ref_time(family, y, x, alpha, beta, aux=1.0, reps=3, threads=1,
             grainsize=None, single_call_in_mt_lib=False):
    """Seconds per lpdf+grad evaluation of the UNMODIFIED reference on the host
    (beta/alpha(/aux) var, x data).  threads>1 -> reduce_sum over TBB.
    single_call_in_mt_lib: take the plain single call from the STAN_THREADS build
    (the two builds define the same symbols and must not share a process)."""
    x = _prep_x(x)
    N, K = x.shape
    fam = FAMILY_ID[family]
    y = _vec(y, np.float64 if fam == 0 else np.int32)
    beta = _vec(beta, np.float64)
    logp = np.zeros(1)
    d_beta = np.zeros(K)
    if threads <= 1:
        t = ref_lib(mt=single_call_in_mt_lib).ref_time_glm(
            fam, _L(N), _L(K), y.ctypes.data_as(C.c_void_p),
                                   _d(x), float(alpha), _d(beta), float(aux),
                                   int(reps), _d(logp), _d(d_beta))
    else:
        if grainsize is None:
            grainsize = max(1, N // (8 * threads))
        t = ref_lib(mt=True).ref_time_glm_reduce_sum(
            fam, _L(N), _L(K), y.ctypes.data_as(C.c_void_p), _d(x), float(alpha),
            _d(beta), float(aux), int(threads), _L(grainsize), int(reps),
            _d(logp), _d(d_beta))
    return t, float(logp[0]), d_beta
