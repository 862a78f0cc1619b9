#!/bin/bash
# gpurun (1 GPU): pass-1 timeline of config 5a from the -DSMC_CAT_TRACE build
mkdir -p gpurun_out
for v in "" "SMC_CAT_NO_DEFER=1"; do
  env $v SMC_CAT_TRACE_FILE=/tmp/t.bin MATH_B200_LIB=profiles/ab/trace/libstanmath_cuda.so python profiles/time_configs.py 5a | cut -c1-120
  python profiles/cat_trace_report.py /tmp/t.bin > gpurun_out/r02_cat_trace${v:+_nodefer}.txt
  cp /tmp/t.bin gpurun_out/cat_trace${v:+_nodefer}.bin
done
