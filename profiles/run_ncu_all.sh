#!/bin/bash
# gpurun (1 GPU): one ncu --set full capture per config below the roofline
# (ordered 5b, neg-binomial with d_x 4, categorical 5a) and the headline (2).
# The reports are read ON the box (raw + source pages as csv); only the csv
# comes back (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
for cfg in ${CFGS:-5b 4 5a 2}; do
  rep=/tmp/ncu_cfg$cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'glm_fused|cat_lin_tma|cat_dbeta_tma' -s 4 -c 2 \
    -o $rep -f python profiles/time_configs.py $cfg > gpurun_out/ncu_$cfg.log 2>&1
  echo "cfg $cfg rc=$?"
  ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/ncu_cfg${cfg}_raw.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_cfg${cfg}_sass.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page details > gpurun_out/ncu_cfg${cfg}_details.txt 2>/dev/null
done
ls -la gpurun_out
