#!/bin/bash
# gpurun (1 GPU), round-2 check: OpenCL ICD probe, GPU test suite, C++ drop-in timings
# (config 2, config 1, config 4 with the rank-one reverse sweep), per-config table.
mkdir -p gpurun_out
{
  echo "== /etc/OpenCL/vendors"; ls -la /etc/OpenCL/vendors 2>&1
  echo "== ldconfig"; ldconfig -p | grep -i -E "opencl|nvidia-opencl" 2>&1
  echo "== find"; find / -xdev \( -name 'libOpenCL*' -o -name 'libnvidia-opencl*' -o -name '*.icd' \) 2>/dev/null | head -20
  echo "== nvidia-smi"; nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv
  echo "== host"; nproc; free -g | head -2
} > gpurun_out/r02_opencl_probe.txt 2>&1
python -m pytest tests/test_boundary_gpu.py tests/test_fuzz_gpu.py -x -q > gpurun_out/pytest_new.log 2>&1; echo "pytest new rc=$?"; tail -15 gpurun_out/pytest_new.log
python -m pytest tests -m gpu -q --deselect tests/test_boundary_gpu.py --deselect tests/test_fuzz_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/pytest_gpu.log
B=tests/cpp/_build/glm_bench
$B 10000000 128 10 3 negbin_xvar | tee gpurun_out/r02_cpp_cfg4.json
$B 10000000 256 20 3 bernoulli | tee gpurun_out/r02_cpp_cfg2.json
$B 10000 100 2000 100 normal | tee gpurun_out/r02_cpp_cfg1.json
python profiles/time_configs.py 1 2 3 4 4b 5a 5b > gpurun_out/r02_configs.jsonl 2> gpurun_out/configs.err; cat gpurun_out/r02_configs.jsonl
