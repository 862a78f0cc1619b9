#!/bin/bash
# gpurun (1 GPU): FP64 pipe microbenchmark + config 1 timing + quick parity of the fused kernel
mkdir -p gpurun_out
profiles/micro/fp64_pipes | tee gpurun_out/r02_fp64_pipes.jsonl
python -m pytest tests/test_glm_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -3
tests/cpp/_build/glm_bench 10000 100 2000 100 normal | tee gpurun_out/r02_cpp_cfg1_new.json
python profiles/time_configs.py 1 2 5b > gpurun_out/r02_configs_new.jsonl 2> gpurun_out/configs.err; cat gpurun_out/r02_configs_new.jsonl
