#!/bin/bash
# gpurun (1 GPU): un-fused categorical_logit_lpmf -- A/B of the per-warp TMA pipeline over
# (lanes per row, warps, stages); each variant in a fresh process
mkdir -p gpurun_out
out=gpurun_out/r02_time_categorical_lpmf_tma2.txt; : > $out
run() { echo "== $*" | tee -a $out; env "$@" timeout 300 python profiles/time_categorical_lpmf.py 2>&1 | grep '"lin_var": true' | cut -c1-80 | tee -a $out; }
run SMC_CATL_L=2 SMC_CATL_W=16 SMC_CATL_S=3
run SMC_CATL_L=2 SMC_CATL_W=16 SMC_CATL_S=2
run SMC_CATL_L=4 SMC_CATL_W=16 SMC_CATL_S=3
run SMC_CATL_L=4 SMC_CATL_W=16 SMC_CATL_S=2
run SMC_CATL_L=1 SMC_CATL_W=16 SMC_CATL_S=2
run SMC_CATL_L=1 SMC_CATL_W=16 SMC_CATL_S=4
timeout 100 python profiles/micro/store_ceiling.py | tee gpurun_out/r02_store_ceiling.jsonl
