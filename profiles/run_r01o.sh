#!/bin/bash
# gpurun (1 GPU): the matrix products either side of the un-fused categorical density
# (smc_linear_predictor_matrix[_adjoint]) -- parity, gtest vs prim, the whole GPU suite, timing.
mkdir -p gpurun_out
python -m pytest tests/test_categorical_lpmf.py -m gpu -x -q 2>&1 | tail -15
tests/cpp/_build/unfused_lpmf_test --gtest_filter='*categorical*' 2>&1 | tail -15
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python profiles/time_categorical_unfused.py 2>&1 | tee gpurun_out/time_categorical_unfused.json
