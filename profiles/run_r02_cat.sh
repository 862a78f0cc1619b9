#!/bin/bash
# gpurun (1 GPU): categorical parity (all class counts) and timings: config 5a, 33-64 classes
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_glm_gpu.py tests/test_fuzz_gpu.py tests/test_golden_gpu.py tests/test_unfused_gpu.py tests/test_full_size_gpu.py -x -q -k "cat or Cat or fuzz or golden" 2>&1 | tail -4
timeout 120 tests/cpp/_build/categorical_logit_glm_test 2>&1 | tail -2
timeout 120 tests/cpp/_build/ref_categorical_logit_glm_lpmf_test 2>&1 | tail -1
for lib in profiles/ab/base/libstanmath_cuda.so math_b200/lib/libstanmath_cuda.so; do
  echo "== $lib"; MATH_B200_LIB=$lib timeout 120 python profiles/time_categorical_wide.py 2>/dev/null
  MATH_B200_LIB=$lib timeout 60 python profiles/time_configs.py 5a | cut -c1-140
done | tee gpurun_out/r02_time_categorical_epilogue.txt
