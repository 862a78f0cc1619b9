#!/bin/bash
# gpurun (1 GPU): un-fused categorical_logit_lpmf -- parity tests, then the per-warp TMA
# pipeline against the LSU kernels (value + d_lin, and data log odds), fresh process each
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "categorical or unfused or fuzz" > gpurun_out/check.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/check.log
out=gpurun_out/r02_time_categorical_lpmf_tma4.txt; : > $out
run() { echo "== $*" | tee -a $out; env "$@" timeout 300 python profiles/time_categorical_lpmf.py 2>&1 | cut -c1-84 | tee -a $out; }
run SMC_CATL_TMA=1
run SMC_CATL_TMA=0
