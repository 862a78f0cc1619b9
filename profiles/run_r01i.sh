#!/bin/bash
# gpurun (1 GPU): full GPU suite, all-config timing incl. the un-fused pipeline (u2), bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python profiles/time_configs.py 1 2 3 4 4b 5a 5b u2 2>&1 | tee gpurun_out/configs.jsonl
python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
