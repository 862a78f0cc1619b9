#!/bin/bash
# gpurun (1 GPU): both bench arms as the driver runs them + the launch list of the bench command
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r02_bench_ref.json
timeout 1200 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "ours rc=$?"; tail -c 300 gpurun_out/r02_bench_n1.err; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
print(d['roofline']); print(d['e2e']); print(d['cpu_baseline'])
for c in d.get('configs',[]): print(c['id'], round(c['ms_per_eval'],4), c.get('roofline',{}).get('frac'), c.get('e2e',{}).get('ms_per_eval'), c.get('cpp_drop_in',{}).get('ms_per_eval'))
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline > /dev/null 2>&1; echo "ncu rc=$?"; grep -c glm_fused gpurun_out/r02_launches.csv
