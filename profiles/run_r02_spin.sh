#!/bin/bash
# gpurun (1 GPU): poll budget of the synchronous call (SMC_SPIN_US 60 vs 5000) -> profiles/r02/r02_spin_budget.txt
mkdir -p gpurun_out; out=gpurun_out/r02_spin_budget.txt; : > $out
for i in 1 2; do for us in 60 5000; do for c in 2 4b 5b; do
  echo -n "SMC_SPIN_US=$us cfg=$c " | tee -a $out
  SMC_SPIN_US=$us timeout 300 python bench_configs.py $c 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('device', round(d['ms_per_eval'],4), 'e2e', round(d['e2e']['ms_per_eval'],4))" | tee -a $out
done; done; done
