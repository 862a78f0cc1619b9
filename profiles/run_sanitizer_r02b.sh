#!/bin/bash
# gpurun (1 GPU): compute-sanitizer memcheck + racecheck over the kernels added late in round 2:
# the per-warp TMA pipeline of categorical_logit_lpmf, ordered_rows_kernel, outer_quad_kernel
mkdir -p gpurun_out
SEL='not full_size and not 50021 and not 200003 and not fuzz and not lsu'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file gpurun_out/sanitizer2_memcheck.log \
  python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_categorical_lpmf.py tests/test_unfused_gpu.py \
    tests/test_boundary_gpu.py -k "$SEL" > gpurun_out/sanitizer2_pytest.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer2_pytest.log; tail -2 gpurun_out/sanitizer2_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 77 --log-file gpurun_out/sanitizer2_racecheck.log \
  python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_categorical_lpmf.py -k "gpu_matches_oracle or golden" \
  > gpurun_out/sanitizer2_pytest_race.log 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/sanitizer2_pytest_race.log; tail -3 gpurun_out/sanitizer2_racecheck.log
grep -oE "[a-z_]+\.(cu|cuh):[0-9]+" gpurun_out/sanitizer2_racecheck.log | sort | uniq -c | head
