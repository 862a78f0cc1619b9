#!/bin/bash
# gpurun (1 GPU): source-level ncu capture of selected configs (CFGS env, default ordered 5b + neg-binomial 4b).
mkdir -p gpurun_out
CFGS="${CFGS:-5b 4b}" bash profiles/run_ncu_all.sh
