#!/bin/bash
# gpurun (1 GPU): the FP64 microbenchmarks behind DESIGN.md section 4.3 (compiled on the box)
mkdir -p gpurun_out /tmp/micro
for m in fp64_pipes dmma_loop dmma_epilogue; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/micro/$m profiles/micro/$m.cu || exit 1
  timeout 120 /tmp/micro/$m | tee gpurun_out/r02_$m.jsonl
done
