#!/bin/bash
# gpurun (1 GPU): what a pure read stream reaches (linear loads, TMA tile loads by box shape) -> profiles/r02/r02_read_ceiling.jsonl
mkdir -p gpurun_out /tmp/micro
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/micro/read_ceiling profiles/micro/read_ceiling.cu || exit 1
timeout 300 /tmp/micro/read_ceiling | tee gpurun_out/r02_read_ceiling.jsonl
