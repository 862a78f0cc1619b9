#ifndef STAN_MATH_CUDA_PRIM_NORMAL_ID_GLM_LPDF_HPP
#define STAN_MATH_CUDA_PRIM_NORMAL_ID_GLM_LPDF_HPP
// normal_id_glm_lpdf for a device-resident design matrix: the B200 overload of
// stan/math/prim/prob/normal_id_glm_lpdf.hpp L54-216 (same name, template order
// and <propto> convention; cf. opencl/prim/normal_id_glm_lpdf.hpp L54-62).
// One fused pass over x (smc_normal_id_glm) yields the value and every partial.
#include <stan/math/cuda/prim/glm_common.hpp>

namespace stan {
namespace math {

template <bool propto, typename T_y, typename T_x, typename T_alpha,
          typename T_beta, typename T_scale,
          require_cuda_design_matrix_t<T_x>* = nullptr,
          require_not_t<is_cuda_operand<T_beta>>* = nullptr>
return_type_t<T_y, T_x, T_alpha, T_beta, T_scale> normal_id_glm_lpdf(
    const T_y& y, const T_x& x, const T_alpha& alpha, const T_beta& beta,
    const T_scale& sigma) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "normal_id_glm_lpdf(CUDA)";
  const int64_t N = x.rows();
  const int64_t K = x.cols();

  // prim L84-89
  if (!is_stan_scalar<T_y>::value) {
    check_size_match(function, "Rows of ", "x", N, "rows of ", "y", operand_size(y));
  }
  check_size_match(function, "Columns of ", "x", K, "size of ", "beta",
                   operand_size(beta));
  if (!is_stan_scalar<T_scale>::value) {
    check_size_match(function, "Rows of ", "x", N, "size of ", "sigma",
                     operand_size(sigma));
  }
  if (!is_stan_scalar<T_alpha>::value) {
    check_size_match(function, "Rows of ", "x", N, "size of ", "alpha",
                     operand_size(alpha));
  }
  row_operand<double, T_scale> sigma_op(sigma, x_handle(x));
  if (sigma_op.handle() == nullptr) {  // check_positive_finite, L93
    check_positive_finite(function, "Scale vector", sigma_op.scalar());
  }
  if (N == 0) {  // size_zero(y, sigma), L95-97
    return 0;
  }
  if (!include_summand<propto, T_y, T_x, T_alpha, T_beta, T_scale>::value) {
    return 0;  // L98-100
  }

  row_operand<double, T_y> y_op(y, x_handle(x));
  row_operand<double, T_alpha> alpha_op(alpha, x_handle(x));
  const Eigen::VectorXd beta_val = host_values(beta);

  auto ops_partials = make_partials_propagator(y, x, alpha, beta, sigma);
  row_partial<T_y> d_y_vec(partials<0>(ops_partials), N, x_handle(x));
  row_partial<T_alpha> d_alpha_vec(partials<2>(ops_partials), N, x_handle(x));
  row_partial<T_scale> d_sigma_vec(partials<4>(ops_partials), N, x_handle(x));

  const unsigned flags
      = (propto ? SMC_PROPTO : 0u) | var_flag<T_y>(SMC_VAR_Y)
        | dx_flags<T_x>() | var_flag<T_alpha>(SMC_VAR_ALPHA)
        | var_flag<T_beta>(SMC_VAR_BETA) | var_flag<T_scale>(SMC_VAR_AUX);
  double logp = 0, d_alpha = 0, d_sigma = 0, d_y = 0;
  Eigen::VectorXd d_beta(K);
  check_cuda_status(
      function,
      smc_normal_id_glm(y_op.handle(), y_op.scalar(), x_handle(x), alpha_op.handle(),
                        alpha_op.scalar(), beta_val.data(), sigma_op.handle(),
                        sigma_op.scalar(), flags, &logp, &d_alpha,
                        d_alpha_vec.handle(), d_beta.data(), &d_sigma,
                        d_sigma_vec.handle(), &d_y, d_y_vec.handle(),
                        dx_factor_handle<T_x>(partials<1>(ops_partials), beta_val.data())));

  if constexpr (!is_constant_all<T_y>::value) {  // L141-147
    if constexpr (is_stan_scalar<T_y>::value) {
      store_host_partial<double>(partials<0>(ops_partials), &d_y, 1);
    } else {
      d_y_vec.store(partials<0>(ops_partials));
    }
  }
  if constexpr (!is_constant_all<T_alpha>::value) {  // L167-173
    if constexpr (is_stan_scalar<T_alpha>::value) {
      store_host_partial<double>(partials<2>(ops_partials), &d_alpha, 1);
    } else {
      d_alpha_vec.store(partials<2>(ops_partials));
    }
  }
  if constexpr (!is_constant_all<T_beta>::value) {  // L158-166
    store_host_partial<T_beta>(partials<3>(ops_partials), d_beta.data(), K);
  }
  if constexpr (!is_constant_all<T_scale>::value) {  // L174-184
    if constexpr (is_stan_scalar<T_scale>::value) {
      store_host_partial<double>(partials<4>(ops_partials), &d_sigma, 1);
    } else {
      d_sigma_vec.store(partials<4>(ops_partials));
    }
  }
  return ops_partials.build(logp);
}

/** beta on the device (the OpenCL overloads' signature): K doubles come to the host,
 * see cuda_internal::host_param. */
template <bool propto, typename T_y, typename T_x, typename T_alpha,
          typename T_beta, typename T_scale,
          require_cuda_design_matrix_t<T_x>* = nullptr,
          require_t<is_cuda_operand<T_beta>>* = nullptr>
return_type_t<T_y, T_x, T_alpha, T_beta, T_scale> normal_id_glm_lpdf(
    const T_y& y, const T_x& x, const T_alpha& alpha, const T_beta& beta,
    const T_scale& sigma) {
  return normal_id_glm_lpdf<propto>(y, x, alpha, cuda_internal::host_param(beta), sigma);
}

// The propto = false forwarding overload is the reference's own
// (prim/prob/normal_id_glm_lpdf.hpp L218-225).

}  // namespace math
}  // namespace stan
#endif
