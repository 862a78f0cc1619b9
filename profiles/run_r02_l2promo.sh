#!/bin/bash
# gpurun (1 GPU): L2 promotion of the tensor maps (none / 64 / 128 / 256 B) on configs 2, 3, 4b, 5b,
# alternating, one fresh process per measurement
mkdir -p gpurun_out; out=gpurun_out/r02_l2promo.txt; : > $out
for i in 1 2; do
  for p in 3 2 0; do
    for c in 2 3 4b 5b; do
      echo -n "promo=$p cfg=$c " | tee -a $out
      SMC_TMAP_L2PROMO=$p timeout 300 python bench_configs.py $c 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['ms_per_eval'],4))" | tee -a $out
    done
  done
done
