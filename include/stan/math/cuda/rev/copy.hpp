#ifndef STAN_MATH_CUDA_REV_COPY_HPP
#define STAN_MATH_CUDA_REV_COPY_HPP
// Autodiff-aware host <-> device copies, the analogue of
// stan/math/opencl/rev/copy.hpp: to_matrix_cuda(var types) returns a
// var_value<matrix_cuda<double>> whose reverse-pass callback does
// `a.adj() += from_matrix_cuda(res.adj())`; from_matrix_cuda(var) goes the other
// way.
#include <stan/math/cuda/copy.hpp>
#include <stan/math/cuda/rev/vari.hpp>
#include <stan/math/rev/core/arena_matrix.hpp>
#include <stan/math/rev/core/reverse_pass_callback.hpp>
#include <stan/math/rev/fun/value_of.hpp>
#include <stan/math/rev/meta.hpp>

#include <vector>

namespace stan {
namespace math {

/** var_value<Eigen> -> device var. */
template <typename T, require_var_t<T>* = nullptr,
          require_eigen_t<value_type_t<T>>* = nullptr>
inline var_value<matrix_cuda<double>> to_matrix_cuda(const T& a) {
  var_value<matrix_cuda<double>> res(to_matrix_cuda(a.val().eval()));
  reverse_pass_callback([a, res]() mutable {
    a.adj() += from_matrix_cuda<plain_type_t<decltype(a.val())>>(
        res.adj().to_matrix_cuda());
  });
  return res;
}

/** Eigen matrix of var -> device var. */
template <typename T, require_eigen_vt<is_var, T>* = nullptr>
inline var_value<matrix_cuda<double>> to_matrix_cuda(const T& src) {
  arena_t<plain_type_t<T>> src_arena(src);
  var_value<matrix_cuda<double>> res(to_matrix_cuda(src_arena.val().eval()));
  reverse_pass_callback([src_arena, res]() mutable {
    src_arena.adj() += from_matrix_cuda<
        Eigen::Matrix<double, T::RowsAtCompileTime, T::ColsAtCompileTime>>(
        res.adj().to_matrix_cuda());
  });
  return res;
}

/** Eigen matrix of var -> device var sharded over the GPUs of the shard set. */
template <typename T, require_eigen_vt<is_var, T>* = nullptr>
inline var_value<matrix_cuda<double>> to_matrix_cuda_sharded(const T& src) {
  arena_t<plain_type_t<T>> src_arena(src);
  var_value<matrix_cuda<double>> res(to_matrix_cuda_sharded(src_arena.val().eval()));
  reverse_pass_callback([src_arena, res]() mutable {
    src_arena.adj() += from_matrix_cuda<
        Eigen::Matrix<double, T::RowsAtCompileTime, T::ColsAtCompileTime>>(
        res.adj().to_matrix_cuda());
  });
  return res;
}

/** std::vector<var> -> device var column. */
inline var_value<matrix_cuda<double>> to_matrix_cuda(const std::vector<var>& src) {
  arena_t<Eigen::Matrix<var, Eigen::Dynamic, 1>> src_arena(
      Eigen::Map<const Eigen::Matrix<var, Eigen::Dynamic, 1>>(src.data(),
                                                               src.size()));
  var_value<matrix_cuda<double>> res(to_matrix_cuda(src_arena.val().eval()));
  reverse_pass_callback([src_arena, res]() mutable {
    src_arena.adj()
        += from_matrix_cuda<Eigen::VectorXd>(res.adj().to_matrix_cuda());
  });
  return res;
}

/** std::vector of Eigen vectors of var -> device var, one vector per column
 * (opencl/rev/copy.hpp L88-104): per-outcome cut points of ordered_logistic_lpmf. */
template <typename T, require_eigen_vt<is_var, T>* = nullptr>
inline var_value<matrix_cuda<double>> to_matrix_cuda(const std::vector<T>& src) {
  using arena_vec = arena_t<plain_type_t<T>>;
  const size_t n = src.size();
  // (arena storage, trivially destructible: captured by the callback)
  arena_vec* src_arena = ChainableStack::instance_->memalloc_.alloc_array<arena_vec>(n);
  std::vector<Eigen::VectorXd> vals;
  vals.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    new (src_arena + i) arena_vec(src[i]);
    vals.emplace_back(Eigen::Map<const Eigen::VectorXd>(src_arena[i].val().eval().data(),
                                                        src_arena[i].size()));
  }
  var_value<matrix_cuda<double>> res(to_matrix_cuda(vals));
  reverse_pass_callback([src_arena, n, res]() mutable {
    if (res.size() == 0) {
      return;
    }
    const Eigen::MatrixXd adj
        = from_matrix_cuda<Eigen::MatrixXd>(res.adj().to_matrix_cuda());
    for (size_t i = 0; i < n; ++i) {
      src_arena[i].adj() += Eigen::Map<const plain_type_t<decltype(src_arena[i].adj())>>(
          adj.data() + adj.rows() * i, src_arena[i].rows(), src_arena[i].cols());
    }
  });
  return res;
}

/** Values of a device var as an owning device matrix (copy). */
inline matrix_cuda<double> value_of(const var_value<matrix_cuda<double>>& a) {
  return a.val().to_matrix_cuda();
}

/** Device var -> host var_value<Eigen>; adjoints flow back to the device. */
template <typename T_dst = Eigen::MatrixXd, require_eigen_t<T_dst>* = nullptr>
inline var_value<T_dst> from_matrix_cuda(const var_value<matrix_cuda<double>>& a) {
  var_value<T_dst> res(from_matrix_cuda<T_dst>(a.val().to_matrix_cuda()));
  reverse_pass_callback([a, res]() mutable {
    // (laid out -- and sharded -- like the adjoint it is added to)
    matrix_cuda<double> g
        = matrix_cuda<double>::like_handle(a.adj().handle(), a.rows(), a.cols());
    const Eigen::MatrixXd g_host = res.adj();
    if (g_host.size() > 0) {
      check_cuda_status("from_matrix_cuda(var)",
                        smc_matrix_upload(g.handle(), g_host.data(), g_host.rows()));
    }
    check_cuda_status("from_matrix_cuda(var)",
                      smc_matrix_axpy(a.adj().handle(), 1.0, g.handle()));
  });
  return res;
}

}  // namespace math
}  // namespace stan
#endif
