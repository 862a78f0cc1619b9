#ifndef STAN_MATH_CUDA_TEST_OPENCL_PRIM_SHIM_HPP
#define STAN_MATH_CUDA_TEST_OPENCL_PRIM_SHIM_HPP
// Two of the reference's GLM device tests include <stan/math/opencl/prim.hpp>: same shim.
#include <stan/math/opencl/rev.hpp>
#endif
