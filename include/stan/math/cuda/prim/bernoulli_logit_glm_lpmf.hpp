#ifndef STAN_MATH_CUDA_PRIM_BERNOULLI_LOGIT_GLM_LPMF_HPP
#define STAN_MATH_CUDA_PRIM_BERNOULLI_LOGIT_GLM_LPMF_HPP
// bernoulli_logit_glm_lpmf for a device-resident design matrix: the B200
// overload of stan/math/prim/prob/bernoulli_logit_glm_lpmf.hpp L49-167 (same
// name, template order and <propto> convention; selected by the type of x the
// way the OpenCL overload is, opencl/prim/bernoulli_logit_glm_lpmf.hpp L52-58).
// Value and partials come from ONE fused pass over x on the GPU
// (smc_bernoulli_logit_glm); they are attached to the tape through the
// reference's own make_partials_propagator(...).build(logp).
#include <stan/math/cuda/prim/glm_common.hpp>

namespace stan {
namespace math {

template <bool propto, typename T_y, typename T_x, typename T_alpha,
          typename T_beta, require_cuda_design_matrix_t<T_x>* = nullptr,
          require_not_t<is_cuda_operand<T_beta>>* = nullptr>
return_type_t<T_x, T_alpha, T_beta> bernoulli_logit_glm_lpmf(
    const T_y& y, const T_x& x, const T_alpha& alpha, const T_beta& beta) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "bernoulli_logit_glm_lpmf(CUDA)";
  const int64_t N = x.rows();
  const int64_t K = x.cols();

  // prim L76-79
  if (!is_stan_scalar<T_y>::value) {
    check_size_match(function, "Rows of ", "x", N, "rows of ", "y", operand_size(y));
  }
  check_size_match(function, "Columns of ", "x", K, "size of ", "beta",
                   operand_size(beta));
  if (!is_stan_scalar<T_alpha>::value) {
    check_size_match(function, "Rows of ", "x", N, "size of ", "alpha",
                     operand_size(alpha));
  }
  if (N == 0) {  // size_zero(y), L80-82
    return 0;
  }
  row_operand<int, T_y> y_op(y, x_handle(x));
  if (y_op.handle() == nullptr) {  // scalar y: check_bounded(y, 0, 1), L85
    check_bounded(function, "Vector of dependent variables", y_op.scalar(), 0, 1);
  }
  if (!include_summand<propto, T_x, T_alpha, T_beta>::value) {  // L87-89
    // the range check on a device y is part of the call; run it for parity
    int lo = 0, hi = 0;
    if (y_op.handle()) {
      check_cuda_status(function, smc_matrix_int_range(y_op.handle(), &lo, &hi));
      check_bounded(function, "Vector of dependent variables", lo, 0, 1);
      check_bounded(function, "Vector of dependent variables", hi, 0, 1);
    }
    return 0;
  }

  row_operand<double, T_alpha> alpha_op(alpha, x_handle(x));
  const Eigen::VectorXd beta_val = host_values(beta);

  auto ops_partials = make_partials_propagator(x, alpha, beta);
  row_partial<T_alpha> d_alpha_vec(partials<1>(ops_partials), N, x_handle(x));

  const unsigned flags = (propto ? SMC_PROPTO : 0u) | dx_flags<T_x>()
                         | var_flag<T_alpha>(SMC_VAR_ALPHA)
                         | var_flag<T_beta>(SMC_VAR_BETA);
  double logp = 0, d_alpha = 0;
  Eigen::VectorXd d_beta(K);
  check_cuda_status(
      function,
      smc_bernoulli_logit_glm(y_op.handle(), y_op.scalar(), x_handle(x),
                              alpha_op.handle(), alpha_op.scalar(), beta_val.data(),
                              flags, &logp, &d_alpha, d_alpha_vec.handle(),
                              d_beta.data(), dx_factor_handle<T_x>(partials<0>(ops_partials), beta_val.data())));

  // partials: d_x was written into the edge by the kernel (L158-159)
  if constexpr (!is_constant_all<T_alpha>::value) {  // L162-164
    if constexpr (is_stan_scalar<T_alpha>::value) {
      store_host_partial<double>(partials<1>(ops_partials), &d_alpha, 1);
    } else {
      d_alpha_vec.store(partials<1>(ops_partials));
    }
  }
  if constexpr (!is_constant_all<T_beta>::value) {  // L149
    store_host_partial<T_beta>(partials<2>(ops_partials), d_beta.data(), K);
  }
  return ops_partials.build(logp);
}

/** beta on the device (the OpenCL overloads' signature): K doubles come to the host,
 * see cuda_internal::host_param. */
template <bool propto, typename T_y, typename T_x, typename T_alpha,
          typename T_beta, require_cuda_design_matrix_t<T_x>* = nullptr,
          require_t<is_cuda_operand<T_beta>>* = nullptr>
return_type_t<T_x, T_alpha, T_beta> bernoulli_logit_glm_lpmf(
    const T_y& y, const T_x& x, const T_alpha& alpha, const T_beta& beta) {
  return bernoulli_logit_glm_lpmf<propto>(y, x, alpha, cuda_internal::host_param(beta));
}

// The propto = false forwarding overload is the reference's own
// (prim/prob/bernoulli_logit_glm_lpmf.hpp L169-174): its call to
// bernoulli_logit_glm_lpmf<false>(...) resolves to the overload above.

}  // namespace math
}  // namespace stan
#endif
