#!/bin/bash
# gpurun (1 GPU): memcheck over the sharded evaluation (direct slots + flags) -> profiles/r02/r02_compute_sanitizer.txt
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file gpurun_out/sanitizer3_memcheck.log \
  python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_sharded_gpu.py -k "not 30011 and not 8191" > gpurun_out/sanitizer3_pytest.log 2>&1
echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer3_pytest.log; tail -2 gpurun_out/sanitizer3_memcheck.log
