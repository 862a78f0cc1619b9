#ifndef STAN_MATH_CUDA_TEST_OPENCL_SHIM_HPP
#define STAN_MATH_CUDA_TEST_OPENCL_SHIM_HPP
// Shim that lets the reference's OWN device tests of the GLMs --
// test/unit/math/opencl/rev/*_glm_*_test.cpp with test/unit/math/opencl/util.hpp, both
// compiled unmodified from where they lie -- run against the CUDA backend: it stands
// where <stan/math/opencl/rev.hpp> stands on their include path and maps the handful of
// OpenCL-backend names they use onto the CUDA types.  Nothing of the OpenCL backend is
// included (tests/cpp/ref_opencl_test.cpp defines STAN_OPENCL only after the reference's
// <stan/math.hpp> has been read without it).
#include <stan/math/rev.hpp>
#include <stan/math/cuda.hpp>

#include <utility>
#include <vector>

namespace stan {

// (util.hpp L81-90 picks its device-side expect_eq with
// require_nonscalar_prim_or_rev_kernel_expression_t; matrix_cuda.hpp registers the CUDA
// device matrix as a non-scalar kernel expression already.)

namespace math {

template <typename T>
using matrix_cl = matrix_cuda<T>;

/** opencl/copy.hpp L57-121, opencl/rev/copy.hpp L24-97 */
template <typename T>
inline auto to_matrix_cl(T&& x) {
  return to_matrix_cuda(std::forward<T>(x));
}
/** opencl/copy.hpp L135-260, opencl/rev/copy.hpp L107-130 */
template <typename T_dst, typename T>
inline auto from_matrix_cl(const T& x) {
  return from_matrix_cuda<T_dst>(x);
}

/** from_matrix_cl(x) without a destination type (opencl/copy.hpp L233-237,
 * opencl/rev/copy.hpp L208-211). */
template <typename T>
inline auto from_matrix_cl(const T& x) {
  return from_matrix_cuda(x);
}
/** opencl/kernel_generator/constant.hpp L110-114 (the one kernel-generator expression
 * test/unit/math/opencl/rev/copy_test.cpp uses, to seed a device adjoint) */
inline matrix_cuda<double> constant(double v, int rows, int cols) {
  return matrix_cuda<double>(Eigen::MatrixXd::Constant(rows, cols, v));
}

}  // namespace math
}  // namespace stan
#endif
