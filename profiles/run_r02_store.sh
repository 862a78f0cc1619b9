#!/bin/bash
# gpurun (1 GPU): whole GPU suite, config 4 with the 32-byte-store d_x stream, timing of the
# per-warp TMA pipeline of the un-fused categorical density
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for i in 1 2; do timeout 300 python bench_configs.py 4 2>/dev/null | tee -a gpurun_out/r02_cfg4_quad_store.jsonl | cut -c1-400; done
out=gpurun_out/r02_time_categorical_lpmf_tma3.txt; : > $out
run() { echo "== $*" | tee -a $out; env "$@" timeout 300 python profiles/time_categorical_lpmf.py 2>&1 | grep '"lin_var": true' | cut -c1-80 | tee -a $out; }
run SMC_CATL_TMA=1
run SMC_CATL_TMA=0
