#!/bin/bash
# gpurun (1 GPU): quick check of a change -- the named pytest selection
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "${1:-ordered}" > gpurun_out/check.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/check.log
