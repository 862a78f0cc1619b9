#!/bin/bash
# gpurun (1 GPU): A/B of config 1 (latency-bound) incl. kernel time from an ncu launch list.
for v in old new old new; do
  echo -n "$v "; MATH_B200_LIB=$PWD/profiles/ab/$v.so python profiles/time_configs.py 1 2>&1 | cut -c1-14,70-130
done
for v in old new; do
  MATH_B200_LIB=$PWD/profiles/ab/$v.so ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 6 --csv --log-file /tmp/l_$v.csv python profiles/time_configs.py 1 > /dev/null 2>&1
  echo -n "$v kernel ns: "; tail -4 /tmp/l_$v.csv | awk -F'","' '{printf "%s ", $NF}'; echo
done
