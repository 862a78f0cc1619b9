#!/bin/bash
# Run under gpurun (1 GPU): bench line, ncu launch list, one full ncu capture of
# the fused kernel.  Outputs land in gpurun_out/ (copy summaries to profiles/).
set -x
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:glm_fused -s 3 -c 1 \
    -o gpurun_out/prof_fused python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
