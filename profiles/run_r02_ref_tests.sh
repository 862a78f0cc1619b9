#!/bin/bash
# gpurun (1 GPU): the reference's own OpenCL-rev GLM tests against the CUDA backend
mkdir -p gpurun_out
for t in bernoulli_logit_glm_lpmf poisson_log_glm_lpmf normal_id_glm_lpdf neg_binomial_2_log_glm_lpmf ordered_logistic_glm_lpmf categorical_logit_glm_lpmf binomial_logit_glm_lpmf bernoulli_logit_lpmf poisson_log_lpmf neg_binomial_2_log_lpmf normal_lpdf ordered_logistic_lpmf copy; do
  timeout 300 tests/cpp/_build/ref_${t}_test > gpurun_out/ref_${t}.log 2>&1; echo "$t rc=$?"; grep -E "^\[  (PASSED|FAILED)|tests ran" gpurun_out/ref_${t}.log | head -12
done
