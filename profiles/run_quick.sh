#!/bin/bash
# gpurun (1 GPU): GPU test suite + per-config timing table.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python profiles/time_configs.py ${CFGS:-1 2 4 4b 5a 5b} > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
cat gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
