#!/bin/bash
# gpurun (1 GPU): categorical_logit_lpmf with shared-memory staging -- parity, timing, one
# ncu --set full capture of its kernel at N=1e7, C=32 (value + d_lin, then value only).
mkdir -p gpurun_out
python -m pytest tests/test_categorical_lpmf.py -m gpu -x -q 2>&1 | tail -5
tests/cpp/_build/unfused_lpmf_test --gtest_filter='*categorical*' 2>&1 | tail -3
python profiles/time_categorical_lpmf.py 2>&1 | tee gpurun_out/time_categorical_lpmf.jsonl
rep=/tmp/ncu_catlpmf
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cat_lpmf_kernel' -s 10 -c 2 \
    -o $rep -f python profiles/time_categorical_lpmf.py 10000000 32 > gpurun_out/ncu_catlpmf.log 2>&1
echo "ncu rc=$?"
ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/ncu_catlpmf_raw.csv 2>/dev/null
ncu -i $rep.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_catlpmf_sass.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/ncu_catlpmf_raw.csv | head -80
