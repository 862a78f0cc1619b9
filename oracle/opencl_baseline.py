#!/usr/bin/env python
"""The reference's own STAN_OPENCL device path timed on this box's GPU (bench.py's
`cpu_baseline.opencl`; BASELINE.json north_star: "the reference STAN_OPENCL path on one
B200 where an OpenCL ICD is present").  Test / bench infrastructure.

Runs in its OWN process: the reference constructs its OpenCL context while the shared
object loads and aborts the process when no platform is found.  The NVIDIA ICD is
handed to the toolkit's loader through OCL_ICD_FILENAMES (no /etc/OpenCL/vendors on
the boxes of this pool).

    python oracle/opencl_baseline.py ROWS COLS REPS   -> one JSON line

The reference's kernels index x with 32-bit integers, so ROWS * COLS must stay below
2^31: the headline N=1e7, K=256 is timed on 8e6 rows and scaled linearly by bench.py.
"""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
SEED = 12345


def main():
    rows, cols, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    os.environ.setdefault("OCL_ICD_FILENAMES", "libnvidia-opencl.so.1")
    import numpy as np
    from oracle import pyoracle as po
    path = os.path.join(HERE, "_ref", "libstan_ref_cl.so")
    if not os.path.exists(path):
        print(json.dumps({"available": False, "why": "oracle/_ref/libstan_ref_cl.so not built"}))
        return
    lib = C.CDLL(path)  # aborts here when no OpenCL platform answers
    lib.ref_cl_last_error.restype = C.c_char_p
    lib.ref_time_glm_opencl.restype = C.c_double
    lib.ref_time_glm_opencl.argtypes = [C.c_int, C.c_long, C.c_long, C.c_void_p, C.c_void_p,
                                        C.c_double, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p]
    name = C.create_string_buffer(256)
    if lib.ref_cl_device_name(name, 256) != 0:
        print(json.dumps({"available": False,
                          "why": (lib.ref_cl_last_error() or b"").decode(errors="replace")}))
        return
    x = po.synthetic(SEED, 0, rows, cols)
    y = po.synthetic(SEED + 1, 0, rows, 1, kind=1, lo=0, hi=1).ravel()
    rng = np.random.default_rng(SEED)
    beta = rng.standard_normal(cols) / np.sqrt(cols)
    logp = np.zeros(1)
    d_beta = np.zeros(cols)
    sec = lib.ref_time_glm_opencl(1, rows, cols, y.ctypes.data_as(C.c_void_p),
                                  x.ctypes.data_as(C.c_void_p), 0.1,
                                  beta.ctypes.data_as(C.c_void_p), reps,
                                  logp.ctypes.data_as(C.c_void_p),
                                  d_beta.ctypes.data_as(C.c_void_p))
    if sec < 0:
        print(json.dumps({"available": False, "device": name.value.decode(errors="replace"),
                          "why": (lib.ref_cl_last_error() or b"").decode(errors="replace")}))
        return
    print(json.dumps({"available": True, "device": name.value.decode(errors="replace"),
                      "rows": rows, "cols": cols, "reps": reps, "sec_per_eval": sec,
                      "logp_per_row": float(logp[0]) / rows, "d_beta0": float(d_beta[0])}))


if __name__ == "__main__":
    main()
