#!/bin/bash
# gpurun (1 GPU) development loop: selected GPU tests + selected config timings.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
for c in ${CFGS:-5a}; do python profiles/time_configs.py $c 2>&1 | tail -2; done
if [ -n "$EXTRA_CMD" ]; then bash -c "$EXTRA_CMD"; fi
