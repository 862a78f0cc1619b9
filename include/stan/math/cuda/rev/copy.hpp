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
#include <stan/math/rev/core/callback_vari.hpp>
#include <stan/math/rev/core/reverse_pass_callback.hpp>
#include <stan/math/rev/fun/value_of.hpp>
#include <stan/math/rev/meta.hpp>

#include <vector>

namespace stan {
namespace math {

/** Device -> Eigen for an arena-owned device matrix (the value or adjoint of a device
 * var): no device copy in between. */
template <typename T_dst, typename T, require_eigen_vt<std::is_arithmetic, T_dst>* = nullptr>
inline T_dst from_matrix_cuda(const arena_matrix_cuda<T>& src) {
  return from_matrix_cuda<T_dst>(matrix_cuda<T>::view_of_handle(src.handle()));
}
template <int Unused = 0, typename T>
inline Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> from_matrix_cuda(
    const arena_matrix_cuda<T>& src) {
  return from_matrix_cuda<Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>>(src);
}

// The host -> device copies return a device var whose vari is the reference's own
// callback_vari (rev/core/callback_vari.hpp L11-22 through make_callback_var, as
// opencl/rev/copy.hpp does): its chain() moves the device adjoint back to the source.

/** var_value<Eigen> -> device var (opencl/rev/copy.hpp L32-38). */
template <typename T, require_var_t<T>* = nullptr,
          require_eigen_t<value_type_t<T>>* = nullptr>
inline var_value<matrix_cuda<double>> to_matrix_cuda(const T& a) {
  return make_callback_var(to_matrix_cuda(a.val().eval()), [a](auto& res_vari) mutable {
    a.adj() += from_matrix_cuda<plain_type_t<decltype(a.val())>>(res_vari.adj());
  });
}

/** Eigen matrix of var -> device var (L66-78). */
template <typename T, require_eigen_vt<is_var, T>* = nullptr>
inline var_value<matrix_cuda<double>> to_matrix_cuda(const T& src) {
  arena_t<plain_type_t<T>> src_arena(src);
  return make_callback_var(
      to_matrix_cuda(src_arena.val().eval()), [src_arena](auto& res_vari) mutable {
        src_arena.adj() += from_matrix_cuda<
            Eigen::Matrix<double, T::RowsAtCompileTime, T::ColsAtCompileTime>>(
            res_vari.adj());
      });
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

/** std::vector<var> -> device var column (L48-54). */
inline var_value<matrix_cuda<double>> to_matrix_cuda(const std::vector<var>& src) {
  return to_matrix_cuda(Eigen::Map<const Eigen::Matrix<var, Eigen::Dynamic, 1>>(
      src.data(), src.size()));
}

/** std::vector of Eigen vectors / matrices of var -> device var, one element per
 * column (L88-104): per-outcome cut points of ordered_logistic_lpmf. */
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
  return make_callback_var(to_matrix_cuda(vals), [src_arena, n](auto& res_vari) mutable {
    if (res_vari.size() == 0) {
      return;
    }
    const Eigen::MatrixXd adj = from_matrix_cuda<Eigen::MatrixXd>(res_vari.adj());
    for (size_t i = 0; i < n; ++i) {
      src_arena[i].adj() += Eigen::Map<const plain_type_t<decltype(src_arena[i].adj())>>(
          adj.data() + adj.rows() * i, src_arena[i].rows(), src_arena[i].cols());
    }
  });
}

/** Values of a device var as an owning device matrix (copy). */
inline matrix_cuda<double> value_of(const var_value<matrix_cuda<double>>& a) {
  return a.val().to_matrix_cuda();
}

namespace internal {
/** a.adj() += host adjoint (uploaded with the layout -- and the row partition -- of
 * the adjoint it is added to). */
template <typename Mat>
inline void add_host_adjoint(const var_value<matrix_cuda<double>>& a, const Mat& g_host) {
  if (g_host.size() == 0) {
    return;
  }
  const Eigen::MatrixXd g_cm = g_host;
  matrix_cuda<double> g
      = matrix_cuda<double>::like_handle(a.adj().handle(), a.rows(), a.cols());
  check_cuda_status("from_matrix_cuda(var)",
                    smc_matrix_upload(g.handle(), g_cm.data(), a.rows()));
  check_cuda_status("from_matrix_cuda(var)",
                    smc_matrix_axpy(a.adj().handle(), 1.0, g.handle()));
}
}  // namespace internal

/** Device var -> host var_value<Eigen>; adjoints flow back to the device
 * (opencl/rev/copy.hpp L115-125).  T_dst names the Eigen type. */
template <typename T_dst = Eigen::MatrixXd,
          require_eigen_vt<std::is_arithmetic, T_dst>* = nullptr>
inline var_value<T_dst> from_matrix_cuda(const var_value<matrix_cuda<double>>& a) {
  return make_callback_var(from_matrix_cuda<T_dst>(a.val()), [a](auto& res_vari) mutable {
    internal::add_host_adjoint(a, res_vari.adj());
  });
}
/** The same with the destination spelled as the reference spells it:
 * from_matrix_cuda<var_value<Eigen::MatrixXd>>(a). */
template <typename T_dst, require_var_vt<is_eigen, T_dst>* = nullptr>
inline T_dst from_matrix_cuda(const var_value<matrix_cuda<double>>& a) {
  return from_matrix_cuda<value_type_t<T_dst>>(a);
}

/** Device var -> Eigen matrix of var (L134-146). */
template <typename T_dst, require_eigen_vt<is_var, T_dst>* = nullptr>
inline T_dst from_matrix_cuda(const var_value<matrix_cuda<double>>& a) {
  arena_t<T_dst> res = from_matrix_cuda<
      Eigen::Matrix<double, T_dst::RowsAtCompileTime, T_dst::ColsAtCompileTime>>(a.val());
  reverse_pass_callback(
      [a, res]() mutable { internal::add_host_adjoint(a, res.adj()); });
  return res;
}

/** Device var (one column) -> std::vector<var> (L155-170). */
template <typename T_dst, require_std_vector_vt<is_var, T_dst>* = nullptr,
          require_all_stan_scalar_t<value_type_t<T_dst>>* = nullptr>
inline T_dst from_matrix_cuda(const var_value<matrix_cuda<double>>& a) {
  check_size_match("from_matrix_cuda<std::vector<var>>", "src.cols()", a.cols(),
                   "dst.cols()", 1);
  const Eigen::VectorXd val = from_matrix_cuda<Eigen::VectorXd>(a.val());
  arena_t<Eigen::Matrix<var, Eigen::Dynamic, 1>> res(val);
  reverse_pass_callback(
      [a, res]() mutable { internal::add_host_adjoint(a, res.adj()); });
  return T_dst(res.data(), res.data() + res.size());
}

/** Device var -> std::vector of Eigen vectors of var or of var_value<Eigen vector>,
 * one per column (L180-198). */
template <typename T_dst, require_std_vector_t<T_dst>* = nullptr,
          require_rev_vector_t<value_type_t<T_dst>>* = nullptr>
inline T_dst from_matrix_cuda(const var_value<matrix_cuda<double>>& a) {
  using elem_t = value_type_t<T_dst>;
  const Eigen::MatrixXd val = from_matrix_cuda<Eigen::MatrixXd>(a.val());
  const size_t n = static_cast<size_t>(a.cols());
  arena_t<elem_t>* res = ChainableStack::instance_->memalloc_.alloc_array<arena_t<elem_t>>(n);
  T_dst out;
  out.reserve(n);
  for (size_t i = 0; i < n; ++i) {
    new (res + i) arena_t<elem_t>(val.col(i));
    out.emplace_back(res[i]);
  }
  reverse_pass_callback([a, res, n]() mutable {
    Eigen::MatrixXd adj(a.rows(), a.cols());
    for (size_t i = 0; i < n; ++i) {
      adj.col(i) = res[i].adj();
    }
    internal::add_host_adjoint(a, adj);
  });
  return out;
}

}  // namespace math
}  // namespace stan
#endif
