#!/bin/bash
# gpurun --gpus 8: weak-scaling bench (config 2 per GPU) at N = 1, 2, 4, 8 and strong scaling of
# config 3 (poisson N=1e8 K=64 row-sharded) at 1, 2, 4, 8 ranks; NCCL parity check at 8 ranks.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale8_n1.json 2> gpurun_out/scale8_n1.err; cut -c1-200 gpurun_out/scale8_n1.json
port=29520
for n in 2 4 8; do
  port=$((port+1))
  $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale8_n$n.json 2> gpurun_out/scale8_n$n.err; cut -c1-200 gpurun_out/scale8_n$n.json
done
python profiles/time_config3_sharded.py > gpurun_out/cfg3_g1.json 2> gpurun_out/cfg3_g1.err; cat gpurun_out/cfg3_g1.json
for n in 2 4 8; do
  port=$((port+1))
  $TR --nproc-per-node $n --master-port $port profiles/time_config3_sharded.py > gpurun_out/cfg3_g$n.json 2> gpurun_out/cfg3_g$n.err; cat gpurun_out/cfg3_g$n.json
done
port=$((port+1))
$TR --nproc-per-node 8 --master-port $port tests/multi_gpu_worker.py > gpurun_out/nccl_parity_8.log 2>&1; echo "nccl parity(8) rc=$?"; tail -2 gpurun_out/nccl_parity_8.log | cut -c1-300
