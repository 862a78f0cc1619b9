#!/bin/bash
# gpurun (1 GPU): A/B of two builds of the library (profiles/ab/old.so vs new.so), alternating,
# one fresh process per measurement so that allocation state is the same for both.
for i in 1 2 3; do
  for v in old new; do
    echo -n "$v "; MATH_B200_LIB=$PWD/profiles/ab/$v.so python profiles/time_configs.py ${CFGS:-5b} 2>&1 | cut -c1-14,80-130
  done
done
