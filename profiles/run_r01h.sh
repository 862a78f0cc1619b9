#!/bin/bash
# gpurun (1 GPU): parity + timing of the per-warp TMA pipeline (TMA-stored d_x) and the ordered inline-path fix.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python profiles/time_configs.py 1 2 3 4 4b 5b 2>&1 | tee gpurun_out/configs_new.jsonl
