#ifndef STAN_MATH_CUDA_PRIM_GLM_COMMON_HPP
#define STAN_MATH_CUDA_PRIM_GLM_COMMON_HPP
// Helpers shared by the six GLM overloads of the CUDA backend: they turn Stan's
// argument types into what the C ABI (stanmath_cuda.h) takes -- device handles
// for x / y / per-row vectors, host doubles for the small parameters -- and move
// the partials the kernel produced into the edges of the reference's own
// partials_propagator (prim/functor/partials_propagator.hpp L118-124,
// rev/functor/partials_propagator.hpp L49-86).
//
// Operand kinds accepted by the overloads
//   x            matrix_cuda<double>  |  var_value<matrix_cuda<double>>
//   per-row      (y, and alpha / sigma / phi when they are vectors)
//                scalar (broadcast)   |  matrix_cuda<T>  |  var_value<matrix_cuda<double>>
//                | any host vector the reference accepts (std::vector, Eigen
//                column / row vector, of arithmetic or var, var_value<Vector>):
//                uploaded on every call -- convenient, but N*8 bytes over PCIe
//                per evaluation; keep per-row data on the device on the hot path
//   parameters   (beta, cuts, categorical alpha / beta) any host type the
//                reference accepts: double, var, std::vector<var>,
//                Eigen::Matrix<var,...>, var_value<Eigen::VectorXd>, ...
#include <stan/math/cuda/copy.hpp>
#include <stan/math/cuda/matrix_cuda.hpp>
#include <stan/math/cuda/rev/copy.hpp>
#include <stan/math/cuda/rev/operands_and_partials.hpp>
#include <stan/math/prim/err.hpp>
#include <stan/math/prim/fun/as_column_vector_or_scalar.hpp>
#include <stan/math/prim/fun/size.hpp>
#include <stan/math/prim/fun/to_ref.hpp>
#include <stan/math/prim/fun/value_of.hpp>
#include <stan/math/prim/functor/partials_propagator.hpp>
#include <stan/math/rev/core/reverse_pass_callback.hpp>

#include <type_traits>
#include <vector>

namespace stan {

/** Any operand that lives on the device (data or autodiff). */
template <typename T>
struct is_cuda_operand
    : bool_constant<is_matrix_cuda<T>::value || is_var_matrix_cuda<T>::value> {};

/** The gate of the CUDA overloads: x is on the device (cf.
 * require_all_prim_or_rev_kernel_expression_t in the OpenCL overloads,
 * opencl/prim/bernoulli_logit_glm_lpmf.hpp L52-58). */
template <typename T_x>
using require_cuda_design_matrix_t = require_t<is_cuda_operand<T_x>>;

namespace math {
namespace cuda_internal {

template <typename... T>
constexpr unsigned var_flag(unsigned bit) {
  return is_constant_all<T...>::value ? 0u : bit;
}

/** Number of elements of an operand, 1 for scalars (math::size semantics). */
template <typename T>
inline int64_t operand_size(const T& v) {
  if constexpr (is_stan_scalar<T>::value) {
    return 1;
  } else if constexpr (is_cuda_operand<T>::value) {
    return v.size();
  } else {
    return static_cast<int64_t>(math::size(v));
  }
}

/** Host values of a small parameter (scalar or vector) as a dense column. */
template <typename T>
inline Eigen::VectorXd host_values(const T& v) {
  if constexpr (is_stan_scalar<T>::value) {
    Eigen::VectorXd r(1);
    r[0] = value_of(v);
    return r;
  } else {
    const auto& ref = to_ref(v);
    const auto& val = value_of(ref);
    return Eigen::VectorXd(as_column_vector_or_scalar(val));
  }
}

/** A small parameter (beta, cut points, the categorical alpha / beta) that the caller
 * keeps on the device -- the OpenCL overloads take every argument as matrix_cl, and a
 * model compiled for that backend passes them so
 * (opencl/prim/bernoulli_logit_glm_lpmf.hpp L52-58) -- comes to the host, K doubles;
 * for a device var the adjoints flow back through one upload + axpy in the reverse
 * sweep.  Host arguments pass through untouched. */
template <typename T, require_not_t<is_cuda_operand<T>>* = nullptr>
inline const T& host_param(const T& v) {
  return v;
}
inline Eigen::VectorXd host_param(const matrix_cuda<double>& v) {
  const Eigen::MatrixXd m = from_matrix_cuda<Eigen::MatrixXd>(v);
  return Eigen::VectorXd(Eigen::Map<const Eigen::VectorXd>(m.data(), m.size()));
}
inline var_value<Eigen::VectorXd> host_param(const var_value<matrix_cuda<double>>& a) {
  const Eigen::MatrixXd m = from_matrix_cuda<Eigen::MatrixXd>(a.val().to_matrix_cuda());
  var_value<Eigen::VectorXd> res(
      Eigen::VectorXd(Eigen::Map<const Eigen::VectorXd>(m.data(), m.size())));
  reverse_pass_callback([a, res]() mutable {
    if (res.size() == 0) {
      return;
    }
    // (column-major: the flat host adjoint has the layout of the n x 1 or 1 x n device
    // matrix it is added to)
    matrix_cuda<double> g
        = matrix_cuda<double>::like_handle(a.adj().handle(), a.rows(), a.cols());
    const Eigen::VectorXd g_host = res.adj();
    check_cuda_status("host_param(var)",
                      smc_matrix_upload(g.handle(), g_host.data(), a.rows()));
    check_cuda_status("host_param(var)",
                      smc_matrix_axpy(a.adj().handle(), 1.0, g.handle()));
  });
  return res;
}
/** The same for a matrix-valued parameter (the categorical K x C beta). */
template <typename T, require_not_t<is_cuda_operand<T>>* = nullptr>
inline const T& host_param_matrix(const T& v) {
  return v;
}
inline Eigen::MatrixXd host_param_matrix(const matrix_cuda<double>& v) {
  return from_matrix_cuda<Eigen::MatrixXd>(v);
}
inline var_value<Eigen::MatrixXd> host_param_matrix(
    const var_value<matrix_cuda<double>>& a) {
  return from_matrix_cuda<Eigen::MatrixXd>(a);
}

/** A per-row operand (y, vector alpha / sigma / phi) as the C ABI wants it:
 * a device handle, or NULL + a broadcast scalar. */
template <typename Elem, typename T>
class row_operand {
 public:
  static constexpr bool is_scalar = is_stan_scalar<T>::value;
  /** `like`: the handle of x; a host vector is uploaded with x's row partition. */
  explicit row_operand(const T& v, const smc_matrix* like = nullptr) {
    if constexpr (is_scalar) {
      scalar_ = static_cast<Elem>(value_of(v));
    } else if constexpr (is_matrix_cuda<T>::value) {
      handle_ = v.handle();
    } else if constexpr (is_var_matrix_cuda<T>::value) {
      handle_ = v.val().handle();
    } else {
      const auto& ref = to_ref(v);
      const auto& val = value_of(ref);
      Eigen::Matrix<Elem, Eigen::Dynamic, 1> col = as_column_vector_or_scalar(val);
      uploaded_ = matrix_cuda<Elem>::like_handle(like, col.size(), 1);
      if (col.size() > 0) {
        check_cuda_status("row_operand",
                          smc_matrix_upload(uploaded_.handle(), col.data(), col.size()));
      }
      handle_ = uploaded_.handle();
    }
  }
  const smc_matrix* handle() const noexcept { return handle_; }
  Elem scalar() const noexcept { return scalar_; }

 private:
  matrix_cuda<Elem> uploaded_;
  const smc_matrix* handle_{nullptr};
  Elem scalar_{0};
};

/** Where the kernel writes the N-vector partial of a per-row autodiff operand:
 * straight into the edge when the operand is a device var, else into a
 * temporary that store() downloads into the (host) edge. */
template <typename T>
class row_partial {
 public:
  static constexpr bool is_var_vector
      = !is_constant_all<T>::value && !is_stan_scalar<T>::value;
  template <typename Edge>
  row_partial(Edge& edge_partials, int64_t n, const smc_matrix* like = nullptr) {
    if constexpr (is_var_vector) {
      if constexpr (is_var_matrix_cuda<T>::value) {
        handle_ = edge_partials.handle();
      } else {
        tmp_ = matrix_cuda<double>::like_handle(like, n, 1);
        handle_ = tmp_.handle();
      }
    }
  }
  smc_matrix* handle() const noexcept { return handle_; }
  template <typename Edge>
  void store(Edge& edge_partials) {
    if constexpr (is_var_vector && !is_var_matrix_cuda<T>::value) {
      using P = std::decay_t<Edge>;
      edge_partials = from_matrix_cuda<
          Eigen::Matrix<double, P::RowsAtCompileTime, P::ColsAtCompileTime>>(tmp_);
    }
  }

 private:
  matrix_cuda<double> tmp_;
  smc_matrix* handle_{nullptr};
};

/** Assigns n doubles to the partials of a host vector (or scalar) edge. */
template <typename T_op, typename Edge>
inline void store_host_partial(Edge& edge_partials, const double* d, int64_t n) {
  if constexpr (is_stan_scalar<T_op>::value) {
    edge_partials[0] = d[0];
  } else {
    using P = std::decay_t<Edge>;
    constexpr int R = P::RowsAtCompileTime, C = P::ColsAtCompileTime;
    edge_partials = Eigen::Map<const Eigen::Matrix<double, R, C>>(
        d, R == 1 ? 1 : n, R == 1 ? n : 1);
  }
}

/** Device handle of the values of x. */
template <typename T_x>
inline const smc_matrix* x_handle(const T_x& x) {
  if constexpr (is_var_matrix_cuda<T_x>::value) {
    return x.val().handle();
  } else {
    return x.handle();
  }
}

/** Device handle the kernel writes the full N x K d_x into: the x edge's own
 * partial (the categorical GLM, whose d_x = T beta^T has rank C). */
template <typename T_x, typename Edge>
inline smc_matrix* dx_handle(Edge& edge_partials) {
  if constexpr (is_var_matrix_cuda<T_x>::value) {
    return edge_partials.handle();
  } else {
    return nullptr;
  }
}

/** The rank-one families (d_x = d beta^T): the x edge keeps the factor d and beta,
 * the product is applied by the reverse sweep (cuda_edge_partial::factored).
 * Returns the N x 1 device vector to pass as d_x together with dx_flags<T_x>(). */
template <typename T_x, typename Edge>
inline smc_matrix* dx_factor_handle(Edge& edge_partials, const double* beta) {
  if constexpr (is_var_matrix_cuda<T_x>::value) {
    return edge_partials.factored(beta);
  } else {
    return nullptr;
  }
}
template <typename T_x>
constexpr unsigned dx_flags() {
  return is_var_matrix_cuda<T_x>::value ? (SMC_VAR_X | SMC_DX_FACTORED) : 0u;
}

}  // namespace cuda_internal
}  // namespace math
}  // namespace stan
#endif
