#!/bin/bash
# gpurun (1 GPU): config 4 in parts, the reference device tests (logs -> profiles/r02/ref_opencl_tests/), ncu of config 4
mkdir -p gpurun_out
timeout 300 python profiles/time_cfg4_parts.py | tee gpurun_out/r02_cfg4_parts.json
bash profiles/run_r02_ref_tests.sh 2>&1 | tail -30
CFGS=4 KERNELS="outer_quad|glm_fused" bash profiles/run_ncu_cfg.sh
