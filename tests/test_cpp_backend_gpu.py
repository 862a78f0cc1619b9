"""Runs the C++ parity tests of the CUDA backend (tests/cpp/*.cpp) on the GPU.

The binaries are built HERE against the reference's own headers
(`make -C tests/cpp`, done by __graft_entry__.build() where /root/reference
exists) and travel to the GPU box in tests/cpp/_build/.  Each one compares the
`stan::math::<family>_glm_*` overloads taking device matrices with the
reference's prim (Eigen) implementation -- linked into the same binary -- for
every prim/var combination of the arguments (see tests/cpp/cuda_test_util.hpp).
"""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "cpp", "_build")
TESTS = ["matrix_cuda_test", "bernoulli_logit_glm_test", "poisson_log_glm_test",
         "normal_id_glm_test", "neg_binomial_2_log_glm_test",
         "ordered_logistic_glm_test", "categorical_logit_glm_test",
         "binomial_logit_glm_test", "unfused_lpmf_test", "reduce_sum_threads_test",
         "sharded_glm_test"]
# The reference's OWN device tests of the GLMs (test/unit/math/opencl/rev/*_glm_*_test.cpp
# with test/unit/math/opencl/util.hpp), compiled unmodified against tests/cpp/ref_shim:
# error_checking, small_simple, broadcast_*, zero_instances, zero_attributes, big, ...
REF_TESTS = ["ref_bernoulli_logit_glm_lpmf_test", "ref_poisson_log_glm_lpmf_test",
             "ref_normal_id_glm_lpdf_test", "ref_neg_binomial_2_log_glm_lpmf_test",
             "ref_ordered_logistic_glm_lpmf_test", "ref_categorical_logit_glm_lpmf_test",
             "ref_binomial_logit_glm_lpmf_test",
             # the un-fused densities whose every device signature the backend has
             "ref_bernoulli_logit_lpmf_test", "ref_poisson_log_lpmf_test",
             "ref_neg_binomial_2_log_lpmf_test", "ref_normal_lpdf_test",
             # incl. one cut-point vector per outcome (a (C-1) x N device matrix)
             "ref_ordered_logistic_lpmf_test",
             # test/unit/math/opencl/rev/copy_test.cpp: host <-> device copies of every var form
             "ref_copy_test"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", TESTS + REF_TESTS)
def test_cpp_backend(gpu, name):
    exe = os.path.join(BUILD, name)
    if not os.path.exists(exe):
        pytest.fail(f"{exe} missing: run `make -C tests/cpp` where /root/reference exists")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    tail = "\n".join((p.stdout + p.stderr).splitlines()[-40:])
    assert p.returncode == 0, tail
    m = re.search(r"\[  PASSED  \] (\d+) tests?", p.stdout)
    assert m and int(m.group(1)) > 0, tail
