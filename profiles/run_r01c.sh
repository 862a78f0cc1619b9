#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python profiles/time_configs.py 1 2 3 4 4b 5a 5b > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
cat gpurun_out/configs.jsonl; tail -3 gpurun_out/configs.err
for cfg in 5b 4b 5a; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'glm_fused|cat_' -s 4 -c 4 \
    -o gpurun_out/prof2_cfg$cfg -f python profiles/time_configs.py $cfg > gpurun_out/ncu2_$cfg.log 2>&1
done
