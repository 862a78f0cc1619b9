#!/bin/bash
# gpurun (1 GPU): ordered-family parity + timing of configs 5b / 2 / 4b after a kernel change.
mkdir -p gpurun_out
python -m pytest tests/test_glm_gpu.py tests/test_golden_gpu.py tests/test_full_size_gpu.py tests/test_unfused_gpu.py tests/test_cpp_backend_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python profiles/time_configs.py 5b 2 4b 2>&1 | tee gpurun_out/configs_quick.jsonl
