#!/bin/bash
# gpurun (1 GPU): ncu evidence for the round's final kernels -- launch list of the bench command,
# then one --set full capture per config (2 headline, 4 d_x via TMA stores, 5b ordered, u2 multiply).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu.log 2>&1
CFGS="2 4 5b u2" bash profiles/run_ncu_all.sh
