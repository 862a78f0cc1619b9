#!/bin/bash
# gpurun (1 GPU): A/B of the synchronous-call latency (C++ drop-in incl. grad()) at small sizes.
# glm_bench links math_b200/lib/libstanmath_cuda.so through its rpath: LD_PRELOAD selects the build.
for i in 1 2; do
  for v in old new; do
    for shape in "10000 100" "100000 64" "10000000 256"; do
      set -- $shape; steps=2000; [ "$1" = "10000000" ] && steps=30
      echo -n "$v N=$1 K=$2: "; LD_PRELOAD=$PWD/profiles/ab/$v.so tests/cpp/_build/glm_bench $1 $2 $steps 50 | grep -o '"ms_per_eval": [0-9.]*'
    done
  done
done
