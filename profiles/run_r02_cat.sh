#!/bin/bash
# gpurun (1 GPU): categorical parity + config 5a A/B of the pass-1 variants
mkdir -p gpurun_out
python -m pytest tests/test_glm_gpu.py tests/test_fuzz_gpu.py tests/test_golden_gpu.py tests/test_unfused_gpu.py -x -q -k "cat or Cat or fuzz or golden" 2>&1 | tail -4
: > gpurun_out/r02_cat_defer.jsonl
for v in "SMC_CAT_LEAD=-1" "SMC_CAT_LEAD=1" "SMC_CAT_LEAD=2" \
         "SMC_CAT_NO_DEFER=1 SMC_CAT_LEAD=-1" "SMC_CAT_NO_DEFER=1 SMC_CAT_LEAD=0" "SMC_CAT_NO_DEFER=1 SMC_CAT_LEAD=1" "SMC_CAT_NO_DEFER=1 SMC_CAT_LEAD=2"; do
  echo "== $v" | tee -a gpurun_out/r02_cat_defer.jsonl
  env $v python profiles/time_configs.py 5a 2>/dev/null | python -c "import sys,json; [print(json.dumps({k:d[k] for k in ('id','ms_per_eval')}|{'frac':d['roofline']['frac']})) for d in map(json.loads,sys.stdin)]" | tee -a gpurun_out/r02_cat_defer.jsonl
done
bash profiles/run_cat_trace.sh
