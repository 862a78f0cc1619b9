#!/bin/bash
# gpurun (1 GPU): categorical parity + config 5a timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_glm_gpu.py tests/test_fuzz_gpu.py tests/test_golden_gpu.py tests/test_unfused_gpu.py tests/test_full_size_gpu.py -x -q -k "cat or Cat or fuzz or golden" 2>&1 | tail -4
: > gpurun_out/r02_cat_final.jsonl
for i in 1 2; do
  timeout 60 python profiles/time_configs.py 5a 2>/dev/null | python -c "import sys,json; [print(json.dumps({k:d[k] for k in ('id','ms_per_eval')}|{'frac':d['roofline']['frac']})) for d in map(json.loads,sys.stdin)]" | tee -a gpurun_out/r02_cat_final.jsonl
done
