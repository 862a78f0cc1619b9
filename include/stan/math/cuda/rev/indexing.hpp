#ifndef STAN_MATH_CUDA_REV_INDEXING_HPP
#define STAN_MATH_CUDA_REV_INDEXING_HPP
// indexing(z, idx): the N-vector z[idx] on the device for a small host vector z
// and a resident index vector -- the hierarchical intercept alpha_i = z[group_i]
// in front of a GLM.  Stands where the OpenCL backend's
// indexing(mat, idx) (opencl/kernel_generator/indexing.hpp L304) and
// indexing_rev (opencl/indexing_rev.hpp L24-60) stand; indices are 0-based as
// there.  G doubles go up per evaluation and G doubles of adjoint come down: the
// N-vector and its partials never leave the GPU, and the reverse sweep is
// deterministic (the reference's kernels use atomics).
#include <stan/math/cuda/prim/glm_common.hpp>
#include <stan/math/rev/core/arena_matrix.hpp>
#include <stan/math/rev/core/reverse_pass_callback.hpp>

#include <stdexcept>
#include <string>

namespace stan {
namespace math {

namespace cuda_internal {
/** Stan's multi-indexing throws std::out_of_range for a bad index
 * (prim/err/check_range.hpp); idx is data, so its range is cached on the device. */
inline void check_index_range(const char* function, const matrix_cuda<int>& idx,
                              int64_t n_elements) {
  if (idx.size() == 0) {
    return;
  }
  int lo = 0, hi = 0;
  check_cuda_status(function, smc_matrix_int_range(idx.handle(), &lo, &hi));
  if (lo < 0 || hi >= n_elements) {
    throw std::out_of_range(std::string(function) + ": accessing element out of range; "
                            "index range [" + std::to_string(lo) + ", "
                            + std::to_string(hi) + "], size "
                            + std::to_string(n_elements));
  }
}
}  // namespace cuda_internal

/** z[idx] for an arithmetic host vector z: an N x 1 device vector. */
template <typename T_z, require_eigen_vector_vt<std::is_arithmetic, T_z>* = nullptr>
inline matrix_cuda<double> indexing(const T_z& z, const matrix_cuda<int>& idx) {
  cuda_internal::check_index_range("indexing(CUDA)", idx, z.size());
  const Eigen::VectorXd zv = z;
  matrix_cuda<double> out(idx.size(), 1);
  check_cuda_status("indexing(CUDA)",
                    smc_indexing(zv.data(), zv.size(), idx.handle(), out.handle()));
  return out;
}

/** z[idx] for an autodiff host vector z (Eigen vector of var or
 * var_value<vector>): a device var whose reverse sweep adds, for every element g
 * of z, the adjoints of the rows that picked it. */
template <typename T_z, require_rev_vector_t<T_z>* = nullptr>
inline var_value<matrix_cuda<double>> indexing(const T_z& z, const matrix_cuda<int>& idx) {
  cuda_internal::check_index_range("indexing(CUDA)", idx, z.size());
  arena_t<T_z> z_arena = z;
  const Eigen::VectorXd zv = value_of(z_arena);
  matrix_cuda<double> out(idx.size(), 1);
  check_cuda_status("indexing(CUDA)",
                    smc_indexing(zv.data(), zv.size(), idx.handle(), out.handle()));
  var_value<matrix_cuda<double>> res(std::move(out));
  // idx is data and outlives the sweep (like every data matrix a callback views): the
  // callback keeps the caller's own handle, so what the library caches on it -- the
  // index range and, for many groups, the sorted row list -- is built once per upload,
  // not once per gradient evaluation on a throw-away view
  const smc_matrix* idx_handle = idx.handle();
  reverse_pass_callback([z_arena, idx_handle, res]() mutable {
    Eigen::VectorXd g = Eigen::VectorXd::Zero(z_arena.size());
    check_cuda_status("indexing(CUDA) reverse",
                      smc_indexing_rev(idx_handle, res.adj().handle(), g.size(),
                                       g.data()));
    if constexpr (std::decay_t<decltype(z_arena.adj())>::ColsAtCompileTime == 1) {
      z_arena.adj() += g;
    } else {
      z_arena.adj() += g.transpose();
    }
  });
  return res;
}

}  // namespace math
}  // namespace stan
#endif
