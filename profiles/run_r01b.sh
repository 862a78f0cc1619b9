#!/bin/bash
# gpurun (1 GPU): GPU test suite, all-config timing, ncu captures of the kernels
# below the roofline (ordered, neg-binomial with x var, categorical).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python profiles/time_configs.py 1 2 4 4b 5a 5b > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
cat gpurun_out/configs.jsonl
for cfg in 5b 4 5a; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'glm_fused|cat_' -s 4 -c 4 \
    -o gpurun_out/prof_cfg$cfg -f python profiles/time_configs.py $cfg > gpurun_out/ncu_$cfg.log 2>&1
done
ls -la gpurun_out
