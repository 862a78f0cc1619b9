#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "categor or cpp" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python profiles/time_configs.py 5a 4 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
cat gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cat_' -s 4 -c 4 \
    -o gpurun_out/prof3_cfg5a -f python profiles/time_configs.py 5a > gpurun_out/ncu3_5a.log 2>&1
