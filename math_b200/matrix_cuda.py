"""Device-resident matrix, the analogue of the reference's ``matrix_cl<T>``
(stan/math/opencl/matrix_cl.hpp L46-55) and of ``to_matrix_cl`` /
``from_matrix_cl`` (stan/math/opencl/copy.hpp L45, L61-235).

Column-major, ``double`` or ``int`` elements.  The N x K design matrix is
uploaded once and reused by every log-density evaluation.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import F64, I32, check, lib

_NP = {F64: np.float64, I32: np.int32}


class MatrixCuda:
    def __init__(self, rows, cols=1, dtype=np.float64, _handle=None, _keep=None):
        self._keep = _keep  # keeps a wrapped buffer's owner alive
        if _handle is not None:
            self._h = _handle
            return
        code = I32 if np.dtype(dtype) == np.int32 else F64
        h = C.c_void_p()
        check(lib().smc_matrix_create(int(rows), int(cols), code, C.byref(h)))
        self._h = h

    # -- construction ------------------------------------------------------
    @classmethod
    def from_host(cls, a):
        """to_matrix_cuda: scalars are not accepted (they stay host scalars)."""
        a = np.asarray(a)
        if a.dtype.kind in "iub":
            a = a.astype(np.int32)
        else:
            a = a.astype(np.float64)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        if a.ndim != 2:
            raise ValueError("to_matrix_cuda: need a vector or a matrix")
        a = np.asfortranarray(a)
        m = cls(a.shape[0], a.shape[1], a.dtype)
        if a.size:
            check(lib().smc_matrix_upload(m._h, a.ctypes.data_as(C.c_void_p),
                                          max(a.shape[0], 1)))
        return m

    # -- row-sharded over the GPUs of the box (runtime.shard_init first) ---------
    @classmethod
    def sharded(cls, rows, cols=1, dtype=np.float64):
        """rows x cols partitioned contiguously over the shard set: accepted wherever
        a GLM takes a plain matrix (x and its per-row operands sharded alike)."""
        code = I32 if np.dtype(dtype) == np.int32 else F64
        h = C.c_void_p()
        check(lib().smc_sharded_matrix_create(int(rows), int(cols), code, C.byref(h)))
        return cls(rows, cols, dtype, _handle=h)

    @classmethod
    def from_host_sharded(cls, a):
        """Scatter a host vector / matrix over the shard set (once per model)."""
        a = np.asarray(a)
        a = a.astype(np.int32) if a.dtype.kind in "iub" else a.astype(np.float64)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        a = np.asfortranarray(a)
        m = cls.sharded(a.shape[0], a.shape[1], a.dtype)
        if a.size:
            check(lib().smc_matrix_upload(m._h, a.ctypes.data_as(C.c_void_p),
                                          max(a.shape[0], 1)))
        return m

    @classmethod
    def like(cls, other, cols=None, dtype=None):
        """A matrix with the rows -- and the row partition -- of `other`."""
        code = -1 if dtype is None else (I32 if np.dtype(dtype) == np.int32 else F64)
        h = C.c_void_p()
        check(lib().smc_matrix_create_like(other._h, -1 if cols is None else int(cols),
                                           code, C.byref(h)))
        return cls(0, 0, _handle=h)

    @property
    def shard_count(self):
        return lib().smc_matrix_shard_count(self._h)

    @classmethod
    def wrap(cls, device_ptr, rows, cols, ld, dtype=np.float64, keep=None):
        """Adopt an existing device buffer (e.g. a torch tensor's data_ptr)."""
        code = I32 if np.dtype(dtype) == np.int32 else F64
        h = C.c_void_p()
        check(lib().smc_matrix_wrap(C.c_void_p(device_ptr), int(rows), int(cols),
                                    int(ld), code, C.byref(h)))
        return cls(rows, cols, dtype, _handle=h, _keep=keep)

    def view_rows(self, row0, nrows):
        """Row block [row0, row0+nrows) as a non-owning view (same ld)."""
        es = 8 if self.dtype == np.float64 else 4
        return MatrixCuda.wrap(self.data_ptr + row0 * es, nrows, self.cols, self.ld,
                               self.dtype, keep=self)

    # -- properties ----------------------------------------------------------
    @property
    def handle(self):
        return self._h

    @property
    def rows(self):
        return lib().smc_matrix_rows(self._h)

    @property
    def cols(self):
        return lib().smc_matrix_cols(self._h)

    @property
    def ld(self):
        return lib().smc_matrix_ld(self._h)

    @property
    def dtype(self):
        return np.dtype(_NP[lib().smc_matrix_dtype(self._h)])

    @property
    def data_ptr(self):
        return lib().smc_matrix_data(self._h) or 0

    def size(self):
        return self.rows * self.cols

    @property
    def shape(self):
        return (self.rows, self.cols)

    # -- transfers -------------------------------------------------------------
    def upload(self, a):
        a = np.asfortranarray(np.asarray(a, dtype=self.dtype))
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        if a.shape != (self.rows, self.cols):
            raise ValueError("upload: shape mismatch")
        if a.size:
            check(lib().smc_matrix_upload(self._h, a.ctypes.data_as(C.c_void_p),
                                          max(a.shape[0], 1)))

    def upload_rows(self, row0, a):
        a = np.asfortranarray(np.asarray(a, dtype=self.dtype))
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        check(lib().smc_matrix_upload_rows(self._h, int(row0), a.shape[0],
                                           a.ctypes.data_as(C.c_void_p),
                                           max(a.shape[0], 1)))

    def to_host(self):
        out = np.zeros((self.rows, self.cols), dtype=self.dtype, order="F")
        if out.size:
            check(lib().smc_matrix_download(self._h, out.ctypes.data_as(C.c_void_p),
                                            max(self.rows, 1)))
        return out

    def rows_to_host(self, row0, nrows):
        out = np.zeros((nrows, self.cols), dtype=self.dtype, order="F")
        if out.size:
            check(lib().smc_matrix_download_rows(
                self._h, int(row0), int(nrows), out.ctypes.data_as(C.c_void_p),
                max(nrows, 1)))
        return out

    def zero(self):
        check(lib().smc_matrix_zero(self._h))

    def zero_lazy(self):
        """Declares the matrix zero without touching memory (the adjoint of a device
        var): the memset only runs if something reads it before a whole-matrix
        writer (axpy, rank1_update, upload, copy) gets there."""
        check(lib().smc_matrix_zero_lazy(self._h))

    def invalidate(self):
        """For wrapped memory whose contents changed behind the library's back."""
        check(lib().smc_matrix_invalidate(self._h))

    def axpy(self, a, x):
        check(lib().smc_matrix_axpy(self._h, float(a), x._h))

    def rank1_update(self, a, d, beta):
        """self[i, k] += a * d[i] * beta[k]: the reverse sweep of an autodiff design
        matrix from the factor d (rev/functor/operands_and_partials.hpp L28-38)."""
        b = np.ascontiguousarray(beta, dtype=np.float64).ravel()
        if b.size != self.cols:
            raise ValueError("rank1_update: size of beta does not match the columns")
        check(lib().smc_matrix_rank1_update(self._h, float(a), d._h,
                                            b.ctypes.data_as(C.POINTER(C.c_double))))

    def fill_synthetic(self, seed, row0=0, kind=0, scale=1.0, lo=0, hi=1):
        check(lib().smc_matrix_fill_synthetic(self._h, int(seed), int(row0),
                                              int(kind), float(scale), int(lo),
                                              int(hi)))

    def int_range(self):
        lo, hi = C.c_int(), C.c_int()
        check(lib().smc_matrix_int_range(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def all_finite(self):
        r = C.c_int()
        check(lib().smc_matrix_all_finite(self._h, C.byref(r)))
        return bool(r.value)

    def free(self):
        if getattr(self, "_h", None):
            lib().smc_matrix_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def to_matrix_cuda(a):
    return a if isinstance(a, MatrixCuda) else MatrixCuda.from_host(a)


def from_matrix_cuda(m):
    return m.to_host()


def synthetic_host(seed, row0, nrows, ncols, kind=0, scale=1.0, lo=0, hi=1):
    """Host replica of smc_matrix_fill_synthetic (bit-identical by construction:
    integer hash, exact int->double conversion, one correctly rounded multiply)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    rows = (np.arange(nrows, dtype=np.uint64) + np.uint64(row0)).reshape(-1, 1)
    cols = np.arange(ncols, dtype=np.uint64).reshape(1, -1)
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + rows * np.uint64(0x9E3779B97F4A7C15)
             + cols * np.uint64(0xBF58476D1CE4E5B9)) & M
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
        z = z ^ (z >> np.uint64(31))
    if kind == 0:
        u = ((z & np.uint64(0xFFFF)) + ((z >> np.uint64(16)) & np.uint64(0xFFFF))
             + ((z >> np.uint64(32)) & np.uint64(0xFFFF)) + (z >> np.uint64(48)))
        c = scale / np.sqrt(4294967295.0 / 3.0)
        return np.asfortranarray((u.astype(np.float64) - 131070.0) * c)
    span = np.uint64(hi - lo + 1)
    return np.asfortranarray((z % span).astype(np.int64) + lo).astype(np.int32)
