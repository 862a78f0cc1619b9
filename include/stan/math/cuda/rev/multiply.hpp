#ifndef STAN_MATH_CUDA_REV_MULTIPLY_HPP
#define STAN_MATH_CUDA_REV_MULTIPLY_HPP
// multiply(x, beta) and add(theta, alpha) for a device-resident design matrix:
// the two steps a model takes when it builds its own linear predictor and calls
// an un-fused density on it.  They stand where the OpenCL
// backend's kernel-generator products stand (opencl/prim/multiply.hpp,
// opencl/rev/multiply.hpp L25-60, opencl/rev/add.hpp): the value is an N-vector
// in HBM, the reverse sweep is one more pass over x (beta.adj += x^T res.adj).
// x is data here; an autodiff x belongs to the fused GLMs, whose kernel writes
// d_x in the same sweep.
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
