#!/usr/bin/env python
"""Phase timeline of glm_fused_kernel from a -DSMC_FUSED_TRACE build (thread 0 of every
CTA stamps %globaltimer at: 0 entry, 1 set-up done, 2 first tile's partial published,
3 tiles done, 4 CTA partial written, 5 ticket taken, 6 (last CTA) result written).
    make -C math_b200/csrc OUT=$PWD/profiles/ab/ftrace EXTRA=-DSMC_FUSED_TRACE
    SMC_FUSED_TRACE_FILE=/tmp/f.bin MATH_B200_LIB=profiles/ab/ftrace/libstanmath_cuda.so \
        python profiles/time_configs.py 1;  python profiles/fused_trace_report.py /tmp/f.bin"""
import sys
import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.uint64)
grid = int(raw[0])
t = raw[1:1 + 8 * grid].reshape(grid, 8).astype(np.int64)
t0 = t[:, 0].min()
names = ["entry", "set-up done", "first partial published", "tiles done", "CTA partial written",
         "ticket taken", "result written (last CTA)"]
print(f"grid {grid}; ns after the first CTA's entry: min / median / max over CTAs")
for k, n in enumerate(names):
    v = t[:, k][t[:, k] > 0] - t0
    if len(v):
        print(f"  {k} {n:28s} {v.min():7d} {int(np.median(v)):7d} {v.max():7d}   (n={len(v)})")
