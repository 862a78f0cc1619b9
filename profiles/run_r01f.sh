#!/bin/bash
# gpurun (1 GPU): parity of the 16-column slab variant, then A/B timing of the slab width per config.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
SMC_FUSED_CPT=16 python -m pytest tests/test_glm_gpu.py tests/test_golden_gpu.py tests/test_binomial_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu_cpt16.log 2>&1; echo "pytest(cpt16) rc=$?"; tail -3 gpurun_out/pytest_gpu_cpt16.log
echo "--- default"; python profiles/time_configs.py 1 4 4b 5b 2>&1 | tee gpurun_out/configs_default.jsonl
echo "--- cpt16";  SMC_FUSED_CPT=16 python profiles/time_configs.py 1 4 4b 5b 2>&1 | tee gpurun_out/configs_cpt16.jsonl
echo "--- cpt32";  SMC_FUSED_CPT=32 python profiles/time_configs.py 5b 2>&1 | tee gpurun_out/configs_cpt32.jsonl
