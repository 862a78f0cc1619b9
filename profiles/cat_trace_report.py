#!/usr/bin/env python
"""Reads the pass-1 trace of a -DSMC_CAT_TRACE build (CTA 0: per warp and stage load q,
clock64 at wait begin / data there / stage done / slice done, and the issue time of every
TMA stage load) and prints where the time of a stage goes.
    SMC_CAT_TRACE_FILE=/tmp/t.bin MATH_B200_LIB=profiles/ab/trace/libstanmath_cuda.so \
        python profiles/time_configs.py 5a;  python profiles/cat_trace_report.py /tmp/t.bin"""
import sys
import numpy as np

W = 16
raw = np.fromfile(sys.argv[1], dtype=np.int64)
t = raw[:512 * W * 4].reshape(512, W, 4).astype(np.float64)
issue = raw[512 * W * 4:].astype(np.float64)
nq = int((t[:, 0, 0] > 0).sum())
t, issue = t[:nq], issue[:nq]
t0 = t[0, :, 0].min()
wait = t[:, :, 1] - t[:, :, 0]
work = t[:, :, 2] - t[:, :, 1]
slic = t[:, :, 3] - t[:, :, 2]
period = np.diff(t[:, :, 0], axis=0)
print(f"stages traced {nq}; mean period per stage {period.mean():.0f} cycles "
      f"(4 warps per SMSP x 32 DMMA x 16 = 2048 at the pipe's rate)")
print(f"mean per warp-stage: wait {wait.mean():.0f}  dmma phase {work.mean():.0f}  slice {slic.mean():.0f}")
lat = t[:, :, 1].min(axis=1) - issue  # first warp to see the data - issue time
ok = issue > 0
print(f"TMA issue -> data visible: median {np.median(lat[ok]):.0f} max {lat[ok].max():.0f} cycles")
slack = t[:, :, 0].min(axis=1) - issue  # earliest consumer arrival - issue
print(f"TMA issue -> first consumer arrives: median {np.median(slack[ok]):.0f} min {slack[ok].min():.0f}")
spread = t[:, :, 0].max(axis=1) - t[:, :, 0].min(axis=1)
print(f"spread of the warps' arrival at a stage: median {np.median(spread):.0f} max {spread.max():.0f}")
print("per-stage detail (q, ks, period, max wait, min wait, mean work, mean slice, issue->first arrival):")
for q in range(0, min(nq, 100)):
    per = period[q].mean() if q < nq - 1 else 0
    print(f"  {q:4d} {q % 32:2d} {per:7.0f} {wait[q].max():7.0f} {wait[q].min():6.0f} {work[q].mean():7.0f} "
          f"{slic[q].mean():7.0f} {slack[q]:8.0f} {lat[q]:8.0f}")
byw = wait.mean(axis=0)
print("mean wait by warp:", " ".join(f"{v:.0f}" for v in byw))
print("mean dmma phase by warp:", " ".join(f"{v:.0f}" for v in work.mean(axis=0)))
