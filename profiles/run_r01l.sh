#!/bin/bash
# gpurun (1 GPU): final evidence of round 1 -- GPU suite, bench (both arms), launch list of the bench
# command, ncu --set full of the kernels that changed last (5b ordered, 5a categorical).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-220 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu.log 2>&1
CFGS="5b 5a" bash profiles/run_ncu_all.sh > /dev/null 2>&1
ls gpurun_out | head -30
