#ifndef STAN_MATH_CUDA_REV_MULTIPLY_HPP
#define STAN_MATH_CUDA_REV_MULTIPLY_HPP
// multiply(x, beta) and add(theta, alpha) for a device-resident design matrix:
// the two steps a model takes when it builds its own linear predictor and calls
// an un-fused density on it.  They stand where the OpenCL
// backend's kernel-generator products stand (opencl/prim/multiply.hpp,
// opencl/rev/multiply.hpp L25-60, opencl/rev/add.hpp): the value is an N-vector
// in HBM, the reverse sweep is one more pass over x (beta.adj += x^T res.adj).
// x is data here; an autodiff x belongs to the fused GLMs, whose kernel writes
// d_x in the same sweep.  A K x C weight MATRIX (categorical models) takes the
// categorical GLM's FP64 tensor-core sweeps: multiply(x, beta_matrix) and
// linear_predictor(x, beta_matrix, alpha) return the N x C matrix of log odds.
#include <stan/math/cuda/prim/glm_common.hpp>
#include <stan/math/rev/core/arena_matrix.hpp>
#include <stan/math/rev/core/reverse_pass_callback.hpp>

namespace stan {
namespace math {

/** x * beta for a device matrix x (N x K, data) and a host K-vector beta of
 * arithmetic type: an N x 1 device vector. */
template <typename T_beta, require_eigen_vector_vt<std::is_arithmetic, T_beta>* = nullptr>
inline matrix_cuda<double> multiply(const matrix_cuda<double>& x, T_beta&& beta) {
  check_size_match("multiply(CUDA)", "Columns of ", "x", x.cols(), "size of ", "beta",
                   beta.size());
  const Eigen::VectorXd b = beta;
  matrix_cuda<double> theta(x.rows(), 1);
  check_cuda_status("multiply(CUDA)",
                    smc_linear_predictor(x.handle(), b.data(), nullptr, 0.0,
                                         theta.handle()));
  return theta;
}

/** x * beta for an autodiff beta (Eigen vector of var or var_value<vector>):
 * a device var; the reverse sweep adds x^T res.adj() to beta's adjoints. */
// (T_beta&& like the reference's own multiply overloads, rev/fun/multiply.hpp
// L137-140, so that this one is the more specialised candidate)
template <typename T_beta, require_rev_vector_t<T_beta>* = nullptr>
inline var_value<matrix_cuda<double>> multiply(const matrix_cuda<double>& x,
                                               T_beta&& beta) {
  check_size_match("multiply(CUDA)", "Columns of ", "x", x.cols(), "size of ", "beta",
                   beta.size());
  arena_t<std::decay_t<T_beta>> beta_arena = beta;
  const Eigen::VectorXd b = value_of(beta_arena);
  matrix_cuda<double> theta(x.rows(), 1);
  check_cuda_status("multiply(CUDA)",
                    smc_linear_predictor(x.handle(), b.data(), nullptr, 0.0,
                                         theta.handle()));
  var_value<matrix_cuda<double>> res(std::move(theta));
  arena_matrix_cuda<double> x_arena = arena_matrix_cuda<double>::view(x);
  reverse_pass_callback([x_arena, beta_arena, res]() mutable {
    Eigen::VectorXd g(beta_arena.size());
    check_cuda_status("multiply(CUDA) reverse",
                      smc_linear_predictor_adjoint(x_arena.handle(),
                                                   res.adj().handle(), g.data(),
                                                   nullptr));
    if constexpr (std::decay_t<decltype(beta_arena.adj())>::ColsAtCompileTime == 1) {
      beta_arena.adj() += g;
    } else {
      beta_arena.adj() += g.transpose();
    }
  });
  return res;
}

namespace cuda_internal {
/** A host matrix with both dimensions dynamic -- Eigen::Matrix<double / var, -1, -1>
 * or var_value<MatrixXd>: the K x C weight matrix of a categorical predictor. */
template <typename T, typename = void>
struct is_host_dynamic_matrix : std::false_type {};
template <typename T>
struct is_host_dynamic_matrix<
    T, std::enable_if_t<is_eigen<T>::value || is_var_matrix<T>::value>>
    : bool_constant<T::RowsAtCompileTime == Eigen::Dynamic
                    && T::ColsAtCompileTime == Eigen::Dynamic> {};

/** lin = x beta (+ 1 alpha^T): one sweep over x per 64 classes on the FP64 tensor
 * pipe (the categorical GLM's first pass with a plain store as its epilogue). */
inline matrix_cuda<double> matrix_product(const char* function, const matrix_cuda<double>& x,
                                          const Eigen::MatrixXd& beta,
                                          const double* alpha) {
  check_size_match(function, "Columns of ", "x", x.cols(), "rows of ", "beta", beta.rows());
  matrix_cuda<double> lin(x.rows(), beta.cols());
  check_cuda_status(function,
                    smc_linear_predictor_matrix(x.handle(), beta.data(), beta.cols(), alpha,
                                                lin.handle()));
  return lin;
}
}  // namespace cuda_internal

/** linear_predictor(x, beta, alpha) = x * beta + rep_matrix(alpha', rows(x)) for a
 * device matrix x (N x K, data), a host K x C weight matrix and a host C-vector of
 * intercepts (column or row; each arithmetic or autodiff): the N x C matrix of log
 * odds of a categorical model, left on the device for categorical_logit_lpmf.  The
 * reverse sweep is one more pass over x: beta.adj += x^T res.adj (opencl/rev/multiply.hpp
 * L25-60), alpha.adj += column sums of res.adj.  (An extension: Stan spells the
 * intercept as a rep_matrix sum; here it rides in the product's epilogue.) */
template <typename T_beta, typename T_alpha,
          require_t<cuda_internal::is_host_dynamic_matrix<std::decay_t<T_beta>>>* = nullptr,
          require_t<bool_constant<is_eigen_vector<T_alpha>::value
                                  || is_rev_vector<T_alpha>::value>>* = nullptr>
inline auto linear_predictor(const matrix_cuda<double>& x, const T_beta& beta,
                             const T_alpha& alpha) {
  static constexpr const char* function = "linear_predictor(CUDA)";
  check_size_match(function, "Columns of ", "beta", beta.cols(), "size of ", "alpha",
                   alpha.size());
  const Eigen::MatrixXd b = value_of(beta);
  Eigen::VectorXd a(alpha.size());
  {
    const auto& alpha_val = value_of(alpha);
    for (Eigen::Index c = 0; c < a.size(); ++c) {
      a[c] = alpha_val.coeff(c);
    }
  }
  matrix_cuda<double> lin = cuda_internal::matrix_product(function, x, b, a.data());
  if constexpr (is_constant_all<T_beta, T_alpha>::value) {
    return lin;
  } else {
    var_value<matrix_cuda<double>> res(std::move(lin));
    arena_matrix_cuda<double> x_arena = arena_matrix_cuda<double>::view(x);
    arena_t<T_beta> beta_arena = beta;
    arena_t<T_alpha> alpha_arena = alpha;
    reverse_pass_callback([x_arena, beta_arena, alpha_arena, res]() mutable {
      const Eigen::Index K = beta_arena.rows(), C = beta_arena.cols();
      Eigen::MatrixXd g(K, C);
      Eigen::VectorXd cs(C);
      check_cuda_status(
          "linear_predictor(CUDA) reverse",
          smc_linear_predictor_matrix_adjoint(
              x_arena.handle(), res.adj().handle(),
              is_constant_all<T_beta>::value ? nullptr : g.data(),
              is_constant_all<T_alpha>::value ? nullptr : cs.data()));
      if constexpr (!is_constant_all<T_beta>::value) {
        beta_arena.adj() += g;
      }
      if constexpr (!is_constant_all<T_alpha>::value) {
        if constexpr (std::decay_t<decltype(alpha_arena.adj())>::ColsAtCompileTime == 1) {
          alpha_arena.adj() += cs;
        } else {
          alpha_arena.adj() += cs.transpose();
        }
      }
    });
    return res;
  }
}

/** x * beta for a device matrix x (N x K, data) and a host K x C matrix beta of
 * arithmetic type: an N x C device matrix. */
template <typename T_beta, require_eigen_matrix_dynamic_vt<std::is_arithmetic, T_beta>* = nullptr>
inline matrix_cuda<double> multiply(const matrix_cuda<double>& x, T_beta&& beta) {
  const Eigen::MatrixXd b = beta;
  return cuda_internal::matrix_product("multiply(CUDA)", x, b, nullptr);
}

/** x * beta for an autodiff K x C beta (Eigen matrix of var or var_value<MatrixXd>):
 * a device var; the reverse sweep adds x^T res.adj() to beta's adjoints. */
template <typename T_beta,
          require_t<cuda_internal::is_host_dynamic_matrix<std::decay_t<T_beta>>>* = nullptr,
          require_st_var<T_beta>* = nullptr>
inline var_value<matrix_cuda<double>> multiply(const matrix_cuda<double>& x, T_beta&& beta) {
  arena_t<std::decay_t<T_beta>> beta_arena = beta;
  const Eigen::MatrixXd b = value_of(beta_arena);
  var_value<matrix_cuda<double>> res(
      cuda_internal::matrix_product("multiply(CUDA)", x, b, nullptr));
  arena_matrix_cuda<double> x_arena = arena_matrix_cuda<double>::view(x);
  reverse_pass_callback([x_arena, beta_arena, res]() mutable {
    Eigen::MatrixXd g(beta_arena.rows(), beta_arena.cols());
    check_cuda_status("multiply(CUDA) reverse",
                      smc_linear_predictor_matrix_adjoint(x_arena.handle(),
                                                          res.adj().handle(), g.data(),
                                                          nullptr));
    beta_arena.adj() += g;
  });
  return res;
}

/** theta + alpha for a device vector theta (data) and an arithmetic scalar. */
inline matrix_cuda<double> add(const matrix_cuda<double>& theta, double alpha) {
  matrix_cuda<double> out(theta.rows(), theta.cols());
  check_cuda_status("add(CUDA)", smc_matrix_copy(out.handle(), theta.handle()));
  check_cuda_status("add(CUDA)", smc_matrix_add_scalar(out.handle(), alpha));
  return out;
}
inline matrix_cuda<double> add(double alpha, const matrix_cuda<double>& theta) {
  return add(theta, alpha);
}

/** theta + alpha where theta (device vector) and / or alpha (scalar) are
 * autodiff variables. */
template <typename T_theta, typename T_alpha,
          require_t<is_cuda_operand<T_theta>>* = nullptr,
          require_stan_scalar_t<T_alpha>* = nullptr,
          require_any_t<is_var_matrix_cuda<T_theta>, is_var<T_alpha>>* = nullptr>
inline var_value<matrix_cuda<double>> add(const T_theta& theta, const T_alpha& alpha) {
  const smc_matrix* th = cuda_internal::x_handle(theta);
  matrix_cuda<double> out(theta.rows(), theta.cols());
  check_cuda_status("add(CUDA)", smc_matrix_copy(out.handle(), th));
  check_cuda_status("add(CUDA)", smc_matrix_add_scalar(out.handle(), value_of(alpha)));
  var_value<matrix_cuda<double>> res(std::move(out));
  // (a data theta is an owning matrix: it must not be captured by the callback)
  if constexpr (is_var_matrix_cuda<T_theta>::value) {
    reverse_pass_callback([theta, res]() mutable {
      check_cuda_status("add(CUDA) reverse",
                        smc_matrix_axpy(theta.adj().handle(), 1.0, res.adj().handle()));
    });
  }
  if constexpr (is_var<T_alpha>::value) {
    reverse_pass_callback([alpha, res]() mutable {
      double s = 0;
      check_cuda_status("add(CUDA) reverse", smc_vector_sum(res.adj().handle(), &s));
      alpha.adj() += s;
    });
  }
  return res;
}
template <typename T_alpha, typename T_theta,
          require_stan_scalar_t<T_alpha>* = nullptr,
          require_t<is_cuda_operand<T_theta>>* = nullptr,
          require_any_t<is_var_matrix_cuda<T_theta>, is_var<T_alpha>>* = nullptr>
inline var_value<matrix_cuda<double>> add(const T_alpha& alpha, const T_theta& theta) {
  return add(theta, alpha);
}

}  // namespace math
}  // namespace stan
#endif
