#!/bin/bash
# gpurun --gpus 2: multi-GPU parity test + the weak-scaling bench at N=1 and N=2, back to back.
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_mgpu.log
python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; cat gpurun_out/scale_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; cat gpurun_out/scale_n2.json; tail -3 gpurun_out/scale_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/scale_ref_n2.json 2> gpurun_out/scale_ref_n2.err; cat gpurun_out/scale_ref_n2.json
