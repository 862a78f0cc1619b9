#!/bin/bash
# gpurun (1 GPU): config 1 (normal_id N=1e4 K=100): parity of the fused kernel, phase timeline, timings
mkdir -p gpurun_out
lscpu | grep -E "Model name|^CPU\(s\)|MHz" | head -4
timeout 600 python -m pytest tests/test_glm_gpu.py tests/test_golden_gpu.py tests/test_fuzz_gpu.py tests/test_binomial_gpu.py tests/test_unfused_gpu.py -x -q 2>&1 | tail -3
if [ -f profiles/ab/ftrace/libstanmath_cuda.so ]; then
SMC_FUSED_TRACE_FILE=/tmp/f.bin MATH_B200_LIB=profiles/ab/ftrace/libstanmath_cuda.so timeout 60 python profiles/time_configs.py 1 > /dev/null
python profiles/fused_trace_report.py /tmp/f.bin | tee gpurun_out/r02_fused_trace_cfg1.txt
fi
timeout 60 python profiles/time_configs.py 1 5b 2 | cut -c1-330 | tee gpurun_out/r02_configs_new.jsonl
timeout 60 tests/cpp/_build/glm_bench 10000 100 2000 100 normal | cut -c1-250 | tee gpurun_out/r02_cpp_cfg1_new.json
