#ifndef STAN_MATH_CUDA_PRIM_BINOMIAL_LOGIT_GLM_LPMF_HPP
#define STAN_MATH_CUDA_PRIM_BINOMIAL_LOGIT_GLM_LPMF_HPP
// binomial_logit_glm_lpmf for a device-resident design matrix: the B200
// overload of stan/math/prim/prob/binomial_logit_glm_lpmf.hpp L54-160 (same
// name, template order and <propto> convention; selected by the type of x the
// way the OpenCL overload is, opencl/prim/binomial_logit_glm_lpmf.hpp L28-45).
// Value and partials come from ONE fused pass over x on the GPU
// (smc_binomial_logit_glm); they are attached to the tape through the
// reference's own make_partials_propagator(...).build(logp).
#include <stan/math/cuda/prim/glm_common.hpp>
#include <stan/math/prim/prob/binomial_logit_glm_lpmf.hpp>

namespace stan {
namespace math {

template <bool propto, typename T_n, typename T_N, typename T_x,
          typename T_alpha, typename T_beta,
          require_cuda_design_matrix_t<T_x>* = nullptr,
          require_not_t<is_cuda_operand<T_beta>>* = nullptr>
return_type_t<T_x, T_alpha, T_beta> binomial_logit_glm_lpmf(
    const T_n& n, const T_N& N, const T_x& x, const T_alpha& alpha,
    const T_beta& beta) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "binomial_logit_glm_lpmf(CUDA)";
  const int64_t N_instances = x.rows();
  const int64_t N_attributes = x.cols();

  // size_zero(n, N, alpha, beta, x) comes before every check, prim L76-78
  if (operand_size(n) == 0 || operand_size(N) == 0 || operand_size(alpha) == 0
      || operand_size(beta) == 0 || N_instances * N_attributes == 0) {
    return 0;
  }
  if (!include_summand<propto, T_x, T_alpha, T_beta>::value) {  // L80-82
    return 0;
  }
  // L88-93 (check_consistent_size: scalars broadcast, vectors must match)
  if (!is_stan_scalar<T_n>::value && !is_stan_scalar<T_N>::value) {
    check_size_match(function, "Size of ", "Successes variable", operand_size(n),
                     "size of ", "Population size parameter", operand_size(N));
  }
  if (!is_stan_scalar<T_n>::value) {
    check_size_match(function, "Rows of ", "x", N_instances, "size of ",
                     "Successes variable", operand_size(n));
  }
  if (!is_stan_scalar<T_N>::value) {
    check_size_match(function, "Rows of ", "x", N_instances, "size of ",
                     "Population size parameter", operand_size(N));
  }
  check_size_match(function, "Columns of ", "x", N_attributes, "size of ",
                   "Weight vector", operand_size(beta));
  if (!is_stan_scalar<T_alpha>::value) {
    check_size_match(function, "Rows of ", "x", N_instances, "size of ",
                     "Vector of intercepts", operand_size(alpha));
  }

  row_operand<int, T_n> n_op(n, x_handle(x));
  row_operand<int, T_N> N_op(N, x_handle(x));
  row_operand<double, T_alpha> alpha_op(alpha, x_handle(x));
  const Eigen::VectorXd beta_val = host_values(beta);

  auto ops_partials = make_partials_propagator(x, alpha, beta);
  row_partial<T_alpha> d_alpha_vec(partials<1>(ops_partials), N_instances, x_handle(x));

  const unsigned flags = (propto ? SMC_PROPTO : 0u) | dx_flags<T_x>()
                         | var_flag<T_alpha>(SMC_VAR_ALPHA)
                         | var_flag<T_beta>(SMC_VAR_BETA);
  double logp = 0, d_alpha = 0;
  Eigen::VectorXd d_beta(N_attributes);
  // check_bounded(n, 0, N), check_nonnegative(N) (L98-99) and the lazy
  // finiteness checks (L119-123) are part of the call
  check_cuda_status(
      function,
      smc_binomial_logit_glm(n_op.handle(), n_op.scalar(), N_op.handle(),
                             N_op.scalar(), x_handle(x), alpha_op.handle(),
                             alpha_op.scalar(), beta_val.data(), flags, &logp,
                             &d_alpha, d_alpha_vec.handle(), d_beta.data(),
                             dx_factor_handle<T_x>(partials<0>(ops_partials), beta_val.data())));

  // partials: d_x was written into the edge by the kernel (L146-149)
  if constexpr (!is_constant_all<T_alpha>::value) {  // L151-153
    if constexpr (is_stan_scalar<T_alpha>::value) {
      store_host_partial<double>(partials<1>(ops_partials), &d_alpha, 1);
    } else {
      d_alpha_vec.store(partials<1>(ops_partials));
    }
  }
  if constexpr (!is_constant_all<T_beta>::value) {  // L139
    store_host_partial<T_beta>(partials<2>(ops_partials), d_beta.data(),
                               N_attributes);
  }
  return ops_partials.build(logp);
}

/** beta on the device (the OpenCL overloads' signature): K doubles come to the host,
 * see cuda_internal::host_param. */
template <bool propto, typename T_n, typename T_N, typename T_x,
          typename T_alpha, typename T_beta,
          require_cuda_design_matrix_t<T_x>* = nullptr,
          require_t<is_cuda_operand<T_beta>>* = nullptr>
return_type_t<T_x, T_alpha, T_beta> binomial_logit_glm_lpmf(
    const T_n& n, const T_N& N, const T_x& x, const T_alpha& alpha,
    const T_beta& beta) {
  return binomial_logit_glm_lpmf<propto>(n, N, x, alpha,
                                         cuda_internal::host_param(beta));
}

// The propto = false forwarding overload is the reference's own
// (prim/prob/binomial_logit_glm_lpmf.hpp L159-165): its call to
// binomial_logit_glm_lpmf<false>(...) resolves to the overload above.

}  // namespace math
}  // namespace stan
#endif
