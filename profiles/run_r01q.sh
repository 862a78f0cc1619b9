#!/bin/bash
# gpurun (1 GPU): evidence after the section 8(f)3 widening (categorical_logit_lpmf, the matrix
# products, wide-C and DMMA d_x): GPU suite incl. the C++ gtests, smoke, bench (both arms), the
# per-config table, the launch list of the bench command, ncu --set full of config 5a.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-200 gpurun_out/bench_ref.json
python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-220 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
python profiles/time_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; cat gpurun_out/configs.jsonl | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
    > gpurun_out/bench_under_ncu.log 2>&1
CFGS="5b" bash profiles/run_ncu_all.sh > /dev/null 2>&1
ls gpurun_out | wc -l
