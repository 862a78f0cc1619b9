#!/bin/bash
# gpurun (1 GPU): the whole GPU test suite, smoke, and both bench arms as the driver runs them
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 1500 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "ours rc=$?"; wc -l gpurun_out/r02_bench_n1.json; python - <<'P'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')})
r=d['roofline']; print(r['frac'], r['traffic'], r['traffic_source'][:80])
for c in d.get('configs',[]): print(c['id'], round(c['ms_per_eval'],4), c.get('roofline',{}).get('frac'), c.get('e2e',{}).get('ms_per_eval'), c.get('cpp_drop_in',{}).get('ms_per_eval'))
P
