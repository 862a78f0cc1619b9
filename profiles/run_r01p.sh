#!/bin/bash
# gpurun (1 GPU): categorical d_x on the DMMA pipe -- parity (pytest + gtest), then A/B against the
# plain-FMA kernel at the config 5a shape and two others.
mkdir -p gpurun_out
python -m pytest tests/test_glm_gpu.py tests/test_golden_gpu.py -m gpu -x -q -k "categorical or golden" 2>&1 | tail -4
tests/cpp/_build/categorical_logit_glm_test 2>&1 | tail -3
for shape in "2000000 512 32" "4000000 128 8" "1000000 256 64"; do
  SMC_CAT_DX_FMA=1 python profiles/time_categorical_dx.py $shape | sed 's/^/fma  /'
  python profiles/time_categorical_dx.py $shape | sed 's/^/dmma /'
done 2>&1 | tee gpurun_out/time_categorical_dx.txt
