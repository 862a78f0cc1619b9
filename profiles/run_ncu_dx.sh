#!/bin/bash
# gpurun (1 GPU): ncu --set full of the categorical d_x kernel and of the stand-alone matrix
# product (lin_only epilogue) at the config 5a shape.
mkdir -p gpurun_out
rep=/tmp/ncu_catdx
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cat_dx_dmma' -s 3 -c 1 \
    -o $rep -f python profiles/time_categorical_dx.py > gpurun_out/ncu_catdx.log 2>&1
echo "ncu rc=$?"
ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/ncu_catdx_raw.csv 2>/dev/null
rep2=/tmp/ncu_catlin
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cat_lin_tma' -s 3 -c 1 \
    -o $rep2 -f python profiles/time_categorical_unfused.py > gpurun_out/ncu_catlin.log 2>&1
echo "ncu rc=$?"
ncu -i $rep2.ncu-rep --page raw --csv > gpurun_out/ncu_catlin_raw.csv 2>/dev/null
python profiles/summarize_ncu.py gpurun_out/ncu_catdx_raw.csv | grep -v "max\|min\|\.sum\.pct" | head -30
python profiles/summarize_ncu.py gpurun_out/ncu_catlin_raw.csv | grep -v "max\|min\|\.sum\.pct" | head -30
