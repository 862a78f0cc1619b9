#!/bin/bash
# gpurun (1 GPU): compute-sanitizer memcheck + racecheck over the fused kernel (round 2: transposing
# column sums, bulk-copied partials in the last CTA, early TMA issue) and the categorical kernels, on
# the small parity cases of every family.
mkdir -p gpurun_out
SEL='not full_size and not 50021 and not 200003 and not fuzz'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file gpurun_out/sanitizer_memcheck.log \
  python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_glm_gpu.py tests/test_binomial_gpu.py \
    tests/test_categorical_lpmf.py tests/test_unfused_gpu.py -k "$SEL" > gpurun_out/sanitizer_pytest.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer_pytest.log; tail -2 gpurun_out/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 77 --log-file gpurun_out/sanitizer_racecheck.log \
  python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_glm_gpu.py -k "(bernoulli or normal or ordered or neg_binomial) and $SEL" \
  > gpurun_out/sanitizer_pytest_race.log 2>&1
echo "racecheck rc=$?"; tail -2 gpurun_out/sanitizer_pytest_race.log; tail -3 gpurun_out/sanitizer_racecheck.log
