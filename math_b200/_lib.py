"""ctypes binding of libstanmath_cuda.so (the C ABI declared in include/stanmath_cuda.h).

The library is built in-tree by ``math_b200.build.build()`` (nvcc, sm_100a).
There is deliberately no fallback: if the shared object is missing or no CUDA
device is present, calls fail loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MATH_B200_LIB points at another build of the same library (A/B timing of two
# kernel variants in one GPU session); it is never a different implementation.
LIB_PATH = os.environ.get("MATH_B200_LIB") or os.path.join(HERE, "lib",
                                                           "libstanmath_cuda.so")

OK, ERR_INVALID_ARGUMENT, ERR_DOMAIN, ERR_CUDA, ERR_UNSUPPORTED = range(5)
F64, I32 = 0, 1
PROPTO, VAR_X, VAR_ALPHA, VAR_BETA, VAR_AUX, VAR_Y = 1, 2, 4, 8, 16, 32
DX_FACTORED = 64
OUT_HEADER, OUT_LOGP, OUT_SUM_D, OUT_AUX, OUT_NONFINITE, OUT_AUX2 = 8, 0, 1, 2, 3, 4


class DomainError(ArithmeticError):
    """Value-domain failure: the reference throws std::domain_error here."""


class BackendError(RuntimeError):
    """CUDA / backend failure (std::system_error in the reference's device path)."""


_P = C.c_void_p
_D = C.c_double
_I = C.c_int
_I64 = C.c_int64
_U = C.c_uint
_DP = C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/stanmath_cuda.h one to one
SIGNATURES = {
    "smc_device_count": (_I, [C.POINTER(_I)]),
    "smc_set_device": (_I, [_I]),
    "smc_get_device": (_I, [C.POINTER(_I)]),
    "smc_set_stream": (_I, [_P]),
    "smc_synchronize": (_I, []),
    "smc_trim_cache": (_I, []),
    "smc_device_info": (_I, [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I),
                             C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "smc_timer_start": (_I, []),
    "smc_timer_stop": (_I, [_DP]),
    "smc_measure_dmma_peak": (_I, [_DP]),
    "smc_last_error": (C.c_char_p, []),
    "smc_launch_count": (_I64, []),
    "smc_reset_launch_count": (None, []),
    "smc_matrix_create": (_I, [_I64, _I64, _I, C.POINTER(_P)]),
    "smc_matrix_wrap": (_I, [_P, _I64, _I64, _I64, _I, C.POINTER(_P)]),
    "smc_matrix_free": (_I, [_P]),
    "smc_matrix_rows": (_I64, [_P]),
    "smc_matrix_cols": (_I64, [_P]),
    "smc_matrix_ld": (_I64, [_P]),
    "smc_matrix_dtype": (_I, [_P]),
    "smc_matrix_data": (_P, [_P]),
    "smc_matrix_upload": (_I, [_P, _P, _I64]),
    "smc_matrix_upload_rows": (_I, [_P, _I64, _I64, _P, _I64]),
    "smc_matrix_download": (_I, [_P, _P, _I64]),
    "smc_matrix_download_rows": (_I, [_P, _I64, _I64, _P, _I64]),
    "smc_matrix_zero": (_I, [_P]),
    "smc_matrix_zero_lazy": (_I, [_P]),
    "smc_matrix_invalidate": (_I, [_P]),
    "smc_matrix_rank1_update": (_I, [_P, _D, _P, _DP]),
    "smc_matrix_copy": (_I, [_P, _P]),
    "smc_matrix_axpy": (_I, [_P, _D, _P]),
    "smc_matrix_outer": (_I, [_P, _P, _DP]),
    "smc_matrix_all_finite": (_I, [_P, C.POINTER(_I)]),
    "smc_matrix_int_range": (_I, [_P, C.POINTER(_I), C.POINTER(_I)]),
    "smc_matrix_fill_synthetic": (_I, [_P, C.c_uint64, _I64, _I, _D, _I, _I]),
    "smc_shard_init": (_I, [_I, C.POINTER(_I)]),
    "smc_shard_count": (_I, [C.POINTER(_I)]),
    "smc_shard_reduce_mode": (C.c_char_p, []),
    "smc_shard_synchronize": (_I, []),
    "smc_shard_shutdown": (_I, []),
    "smc_sharded_matrix_create": (_I, [_I64, _I64, _I, C.POINTER(_P)]),
    "smc_matrix_create_like": (_I, [_P, _I64, _I, C.POINTER(_P)]),
    "smc_matrix_view": (_I, [_P, C.POINTER(_P)]),
    "smc_matrix_shard_count": (_I, [_P]),
    "smc_matrix_shard": (_I, [_P, _I, C.POINTER(_P), C.POINTER(_I64), C.POINTER(_I)]),
    "smc_bernoulli_logit_glm": (_I, [_P, _I, _P, _P, _D, _DP, _U, _DP, _DP, _P,
                                     _DP, _P]),
    "smc_binomial_logit_glm": (_I, [_P, _I, _P, _I, _P, _P, _D, _DP, _U, _DP, _DP,
                                    _P, _DP, _P]),
    "smc_poisson_log_glm": (_I, [_P, _I, _P, _P, _D, _DP, _U, _DP, _DP, _P, _DP,
                                 _P]),
    "smc_normal_id_glm": (_I, [_P, _D, _P, _P, _D, _DP, _P, _D, _U, _DP, _DP, _P,
                               _DP, _DP, _P, _DP, _P, _P]),
    "smc_neg_binomial_2_log_glm": (_I, [_P, _I, _P, _P, _D, _DP, _P, _D, _U, _DP,
                                        _DP, _P, _DP, _DP, _P, _P]),
    "smc_ordered_logistic_glm": (_I, [_P, _I, _P, _DP, _DP, _I64, _U, _DP, _DP,
                                      _DP, _P]),
    "smc_categorical_logit_glm": (_I, [_P, _I, _P, _DP, _DP, _I64, _U, _DP, _DP,
                                       _DP, _P]),
    "smc_matrix_add_scalar": (_I, [_P, _D]),
    "smc_linear_predictor": (_I, [_P, _DP, _P, _D, _P]),
    "smc_linear_predictor_adjoint": (_I, [_P, _P, _DP, _DP]),
    "smc_linear_predictor_matrix": (_I, [_P, _DP, _I64, _DP, _P]),
    "smc_linear_predictor_matrix_adjoint": (_I, [_P, _P, _DP, _DP]),
    "smc_vector_sum": (_I, [_P, _DP]),
    "smc_indexing": (_I, [_DP, _I64, _P, _P]),
    "smc_indexing_rev": (_I, [_P, _P, _I64, _DP]),
    "smc_bernoulli_logit_lpmf": (_I, [_P, _I, _P, _U, _DP, _P]),
    "smc_poisson_log_lpmf": (_I, [_P, _I, _P, _U, _DP, _P]),
    "smc_neg_binomial_2_log_lpmf": (_I, [_P, _I, _P, _P, _D, _U, _DP, _P, _DP, _P]),
    "smc_normal_lpdf": (_I, [_P, _D, _P, _D, _D, _U, _DP, _P, _DP, _P, _DP, _DP]),
    "smc_ordered_logistic_lpmf": (_I, [_P, _I, _P, _DP, _I64, _U, _DP, _P, _DP]),
    "smc_ordered_logistic_lpmf_rows": (_I, [_P, _I, _P, _P, _U, _DP, _P, _P]),
    "smc_categorical_logit_lpmf": (_I, [_P, _I, _P, _U, _DP, _P]),
    "smc_categorical_logit_glm_device": (_I, [_P, _I, _P, _P, _I64, _U, _P, _P]),
    "smc_glm_eval_device": (_I, [_I, _P, _D, _P, _P, _D, _P, _D, _P, _I64, _U, _P,
                                 _P, _P, _P, _P]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc == OK:
        return
    msg = (lib().smc_last_error() or b"").decode(errors="replace")
    if rc == ERR_INVALID_ARGUMENT:
        raise ValueError(msg)  # std::invalid_argument
    if rc == ERR_DOMAIN:
        raise DomainError(msg)  # std::domain_error
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise BackendError(msg)
