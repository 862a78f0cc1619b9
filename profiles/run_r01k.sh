#!/bin/bash
# gpurun (1 GPU): GPU suite after the normal_lpdf widening + kernel time of the small config 1.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 10 --csv --log-file gpurun_out/launches_cfg1.csv \
   python profiles/time_configs.py 1 > gpurun_out/cfg1_under_ncu.log 2>&1
cut -d, -f5,15 gpurun_out/launches_cfg1.csv | tail -6
