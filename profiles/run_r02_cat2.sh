#!/bin/bash
# gpurun (1 GPU): pass 1 of config 5a, production build vs the -DSMC_CAT_HACK=1 build (no epilogue; timing only)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_glm_gpu.py tests/test_fuzz_gpu.py tests/test_golden_gpu.py tests/test_unfused_gpu.py -x -q -k "cat or Cat or fuzz or golden" 2>&1 | tail -4
: > gpurun_out/r02_cat_hack.jsonl
for h in 0 1; do
  echo "== hack $h" | tee -a gpurun_out/r02_cat_hack.jsonl
  lib=profiles/ab/h$h/libstanmath_cuda.so; [ $h = 0 ] && lib=math_b200/lib/libstanmath_cuda.so
  MATH_B200_LIB=$lib timeout 60 python profiles/time_configs.py 5a 2>/dev/null | python -c "import sys,json; [print(json.dumps({k:d[k] for k in ('id','ms_per_eval')})) for d in map(json.loads,sys.stdin)]" | tee -a gpurun_out/r02_cat_hack.jsonl
done
