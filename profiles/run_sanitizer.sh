#!/bin/bash
# gpurun (1 GPU): compute-sanitizer memcheck over the kernels added in the section 8(f)3 widening
# (cat_lpmf_kernel, lin_only epilogues, cat_dx_dmma_kernel, colsum, indexing_rev_sorted) and the
# ordered link, on the small parity cases.
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file gpurun_out/sanitizer_memcheck.log \
  python -m pytest -x -q -m gpu -p no:cacheprovider \
    "tests/test_categorical_lpmf.py" "tests/test_glm_gpu.py" "tests/test_unfused_gpu.py" \
    -k "(categorical or matrix_product or unfused_categorical or indexing or ordered) and not full_size and not 50021 and not 200003" \
  > gpurun_out/sanitizer_pytest.log 2>&1
echo "sanitizer rc=$?"
tail -3 gpurun_out/sanitizer_pytest.log
grep -c "Invalid\|Misaligned\|out of bounds" gpurun_out/sanitizer_memcheck.log
tail -5 gpurun_out/sanitizer_memcheck.log
