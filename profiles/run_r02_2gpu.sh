#!/bin/bash
# gpurun --gpus 2: multi-GPU parity (NCCL ranks, single-process shards), sharded gtest, bench at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_sharded_gpu.py -x -q > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_2gpu.log
timeout 300 tests/cpp/_build/sharded_glm_test > gpurun_out/sharded_gtest_2gpu.log 2>&1; echo "gtest rc=$?"; tail -3 gpurun_out/sharded_gtest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_n2.json
