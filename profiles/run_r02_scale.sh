#!/bin/bash
# gpurun --gpus N: bench.py at N ranks as the driver launches it (weak scaling of config 2, strong scaling of
# config 3, sharded e2e, NCCL parity, single-process sharded leg) + the single-process sharded gtest
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
python - <<P
import json
txt=open('gpurun_out/r02_bench_n$N.json').read().split('\n')
d=json.loads([l for l in txt if l.startswith('{')][0])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print('strong', {k:d['strong'][k] for k in ('ms_per_eval','efficiency_vs_n1','aggregate_GBps','logp_rel_vs_one_gpu')})
print('check', d['check'])
print('single process', json.dumps(d.get('single_process_sharded'))[:900])
P
timeout 300 tests/cpp/_build/sharded_glm_test > gpurun_out/sharded_gtest_${N}gpu.log 2>&1; echo "gtest rc=$?"; tail -3 gpurun_out/sharded_gtest_${N}gpu.log
