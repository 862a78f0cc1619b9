#!/bin/bash
# gpurun --gpus N: the single-process sharded path with the NCCL all-reduce and with the direct
# (zero-copy slots + completion flags) reduction, plus the sharded pytest / gtest on real GPUs
N=${N:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q > gpurun_out/pytest_sharded_${N}gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_sharded_${N}gpu.log
timeout 600 python bench_sharded.py $N 20 3 > gpurun_out/r02_sharded_direct_n$N.json 2> gpurun_out/r02_sharded_direct_n$N.err; echo "bench rc=$?"
python - <<P
import json
d=json.load(open('gpurun_out/r02_sharded_direct_n$N.json'))
for name, r in (("nccl", d), ("direct", d.get("direct", {}))):
    if "weak" in r:
        print(name, r["reduce"], "weak evals/s", round(r["weak"]["value"],1), "strong ms", round(r["strong"]["e2e_ms_per_eval"],4), "rel", r["vs_single_gpu_rel"])
    else:
        print(name, r)
P
SMC_SHARD_REDUCE=direct timeout 300 tests/cpp/_build/sharded_glm_test > gpurun_out/sharded_gtest_direct_${N}gpu.log 2>&1; echo "gtest(direct) rc=$?"; tail -3 gpurun_out/sharded_gtest_direct_${N}gpu.log
