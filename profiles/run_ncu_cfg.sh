#!/bin/bash
# gpurun (1 GPU): `ncu --set full` of the kernels of one or more configs (CFGS, default
# "5a"), read ON the box into raw / source csv + details (gpurun_out/ is capped at
# 64 MiB, the .ncu-rep stays behind), then summarised here with
#   python profiles/summarize_ncu.py gpurun_out/ncu_cfg5a_raw.csv > profiles/r02/r02_ncu_cfg5a.txt
# KERNELS: regex of kernel names; SKIP / COUNT: launches to skip / capture; SASS=1 keeps
# the per-instruction source page too.
mkdir -p gpurun_out
for cfg in ${CFGS:-5a}; do
  rep=/tmp/ncu_cfg$cfg
  timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"${KERNELS:-glm_fused|cat_lin_tma|cat_dbeta_tma|outer_kernel|axpy_kernel|generic_}" \
    -s ${SKIP:-4} -c ${COUNT:-2} -o $rep -f python bench_configs.py $cfg > gpurun_out/ncu_$cfg.log 2>&1
  echo "cfg $cfg rc=$?"
  ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/ncu_cfg${cfg}_raw.csv 2>/dev/null
  [ -n "$SASS" ] && ncu -i $rep.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_cfg${cfg}_sass.csv 2>/dev/null
  ncu -i $rep.ncu-rep --page details > gpurun_out/ncu_cfg${cfg}_details.txt 2>/dev/null
done
ls -la gpurun_out | grep ncu_ | head -40
