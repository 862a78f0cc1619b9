#ifndef STAN_MATH_CUDA_PRIM_NEG_BINOMIAL_2_LOG_GLM_LPMF_HPP
#define STAN_MATH_CUDA_PRIM_NEG_BINOMIAL_2_LOG_GLM_LPMF_HPP
// neg_binomial_2_log_glm_lpmf for a device-resident design matrix: the B200
// overload of stan/math/prim/prob/neg_binomial_2_log_glm_lpmf.hpp L64-248 (same
// name, template order and <propto> convention; cf.
// opencl/prim/neg_binomial_2_log_glm_lpmf.hpp L62-72).  One fused pass over x
// (smc_neg_binomial_2_log_glm) yields the value and every partial, including
// d_x = beta (x) d written into the device-resident x edge when x is a var.
#include <stan/math/cuda/prim/glm_common.hpp>

namespace stan {
namespace math {

template <bool propto, typename T_y, typename T_x, typename T_alpha,
          typename T_beta, typename T_precision,
          require_cuda_design_matrix_t<T_x>* = nullptr,
          require_not_t<is_cuda_operand<T_beta>>* = nullptr>
return_type_t<T_x, T_alpha, T_beta, T_precision> neg_binomial_2_log_glm_lpmf(
    const T_y& y, const T_x& x, const T_alpha& alpha, const T_beta& beta,
    const T_precision& phi) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "neg_binomial_2_log_glm_lpmf(CUDA)";
  const int64_t N = x.rows();
  const int64_t K = x.cols();

  // prim L102-107
  if (!is_stan_scalar<T_y>::value) {
    check_size_match(function, "Rows of ", "x", N, "rows of ", "y", operand_size(y));
  }
  check_size_match(function, "Columns of ", "x", K, "size of ", "beta",
                   operand_size(beta));
  if (!is_stan_scalar<T_precision>::value) {
    check_size_match(function, "Rows of ", "x", N, "size of ", "phi",
                     operand_size(phi));
  }
  if (!is_stan_scalar<T_alpha>::value) {
    check_size_match(function, "Rows of ", "x", N, "size of ", "alpha",
                     operand_size(alpha));
  }
  const Eigen::VectorXd beta_val = host_values(beta);
  check_finite(function, "Weight vector", beta_val);  // L114
  row_operand<double, T_alpha> alpha_op(alpha, x_handle(x));
  if (alpha_op.handle() == nullptr) {  // L115 (a vector alpha is checked in the sweep)
    check_finite(function, "Intercept", alpha_op.scalar());
  }
  if (N == 0) {  // size_zero(y, phi), L117-119
    return 0;
  }
  row_operand<int, T_y> y_op(y, x_handle(x));
  row_operand<double, T_precision> phi_op(phi, x_handle(x));
  if (y_op.handle() == nullptr) {  // L129
    check_nonnegative(function, "Failures variables", y_op.scalar());
  }
  if (phi_op.handle() == nullptr) {  // L130
    check_positive_finite(function, "Precision parameter", phi_op.scalar());
  }
  if (!include_summand<propto, T_x, T_alpha, T_beta, T_precision>::value) {
    int lo = 0, hi = 0;  // the y check of L129 precedes this return (L132-134)
    if (y_op.handle()) {
      check_cuda_status(function, smc_matrix_int_range(y_op.handle(), &lo, &hi));
      check_nonnegative(function, "Failures variables", lo);
    }
    return 0;
  }

  auto ops_partials = make_partials_propagator(x, alpha, beta, phi);
  row_partial<T_alpha> d_alpha_vec(partials<1>(ops_partials), N, x_handle(x));
  row_partial<T_precision> d_phi_vec(partials<3>(ops_partials), N, x_handle(x));

  const unsigned flags
      = (propto ? SMC_PROPTO : 0u) | dx_flags<T_x>()
        | var_flag<T_alpha>(SMC_VAR_ALPHA) | var_flag<T_beta>(SMC_VAR_BETA)
        | var_flag<T_precision>(SMC_VAR_AUX);
  double logp = 0, d_alpha = 0, d_phi = 0;
  Eigen::VectorXd d_beta(K);
  check_cuda_status(
      function,
      smc_neg_binomial_2_log_glm(
          y_op.handle(), y_op.scalar(), x_handle(x), alpha_op.handle(),
          alpha_op.scalar(), beta_val.data(), phi_op.handle(), phi_op.scalar(),
          flags, &logp, &d_alpha, d_alpha_vec.handle(), d_beta.data(), &d_phi,
          d_phi_vec.handle(), dx_factor_handle<T_x>(partials<0>(ops_partials), beta_val.data())));

  if constexpr (!is_constant_all<T_alpha>::value) {  // L225-231
    if constexpr (is_stan_scalar<T_alpha>::value) {
      store_host_partial<double>(partials<1>(ops_partials), &d_alpha, 1);
    } else {
      d_alpha_vec.store(partials<1>(ops_partials));
    }
  }
  if constexpr (!is_constant_all<T_beta>::value) {  // L211-212
    store_host_partial<T_beta>(partials<2>(ops_partials), d_beta.data(), K);
  }
  if constexpr (!is_constant_all<T_precision>::value) {  // L233-245
    if constexpr (is_stan_scalar<T_precision>::value) {
      store_host_partial<double>(partials<3>(ops_partials), &d_phi, 1);
    } else {
      d_phi_vec.store(partials<3>(ops_partials));
    }
  }
  return ops_partials.build(logp);
}

/** beta on the device (the OpenCL overloads' signature): K doubles come to the host,
 * see cuda_internal::host_param. */
template <bool propto, typename T_y, typename T_x, typename T_alpha,
          typename T_beta, typename T_precision,
          require_cuda_design_matrix_t<T_x>* = nullptr,
          require_t<is_cuda_operand<T_beta>>* = nullptr>
return_type_t<T_x, T_alpha, T_beta, T_precision> neg_binomial_2_log_glm_lpmf(
    const T_y& y, const T_x& x, const T_alpha& alpha, const T_beta& beta,
    const T_precision& phi) {
  return neg_binomial_2_log_glm_lpmf<propto>(y, x, alpha,
                                             cuda_internal::host_param(beta), phi);
}

// The propto = false forwarding overload is the reference's own
// (prim/prob/neg_binomial_2_log_glm_lpmf.hpp L250-257).

}  // namespace math
}  // namespace stan
#endif
