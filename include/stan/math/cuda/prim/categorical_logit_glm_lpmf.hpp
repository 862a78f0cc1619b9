#ifndef STAN_MATH_CUDA_PRIM_CATEGORICAL_LOGIT_GLM_LPMF_HPP
#define STAN_MATH_CUDA_PRIM_CATEGORICAL_LOGIT_GLM_LPMF_HPP
// categorical_logit_glm_lpmf for a device-resident design matrix: the B200
// overload of stan/math/prim/prob/categorical_logit_glm_lpmf.hpp L43-195 (same
// name, template order and <propto> convention; cf.
// opencl/prim/categorical_logit_glm_lpmf.hpp L44-52).  Both N x K x C
// contractions run on the FP64 tensor pipe (smc_categorical_logit_glm).
#include <stan/math/cuda/prim/glm_common.hpp>

namespace stan {
namespace math {

template <bool propto, typename T_y, typename T_x, typename T_alpha,
          typename T_beta, require_cuda_design_matrix_t<T_x>* = nullptr,
          require_col_vector_t<T_alpha>* = nullptr,
          require_matrix_t<T_beta>* = nullptr>
return_type_t<T_x, T_alpha, T_beta> categorical_logit_glm_lpmf(
    const T_y& y, const T_x& x, const T_alpha& alpha, const T_beta& beta) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "categorical_logit_glm_lpmf(CUDA)";
  const int64_t N = x.rows();
  const int64_t K = x.cols();
  const int64_t C = beta.cols();

  // prim L68-72
  if (!is_stan_scalar<T_y>::value) {
    check_size_match(function, "Rows of ", "x", N, "rows of ", "y", operand_size(y));
  }
  check_size_match(function, "Columns of ", "beta", C, "size of ", "alpha",
                   operand_size(alpha));
  check_size_match(function, "x.cols()", K, "beta.rows()", beta.rows());
  if (N == 0 || C == 1) {  // L73-75
    return 0;
  }
  row_operand<int, T_y> y_op(y, x_handle(x));
  if (y_op.handle() == nullptr) {  // L77 (a device y is checked by the call)
    check_bounded(function, "categorical outcome out of support", y_op.scalar(), 1,
                  C);
  }
  if (!include_summand<propto, T_x, T_alpha, T_beta>::value) {  // L80-82
    int lo = 1, hi = 1;
    if (y_op.handle()) {
      check_cuda_status(function, smc_matrix_int_range(y_op.handle(), &lo, &hi));
      check_bounded(function, "categorical outcome out of support", lo, 1, C);
      check_bounded(function, "categorical outcome out of support", hi, 1, C);
    }
    return 0;
  }

  const Eigen::VectorXd alpha_val = host_values(alpha);
  const auto& beta_ref = to_ref(beta);
  const Eigen::MatrixXd beta_val = value_of(beta_ref);  // column-major K x C

  auto ops_partials = make_partials_propagator(x, alpha, beta);

  const unsigned flags = (propto ? SMC_PROPTO : 0u) | var_flag<T_x>(SMC_VAR_X)
                         | var_flag<T_alpha>(SMC_VAR_ALPHA)
                         | var_flag<T_beta>(SMC_VAR_BETA);
  double logp = 0;
  Eigen::VectorXd d_alpha(C);
  Eigen::MatrixXd d_beta(K, C);
  check_cuda_status(
      function,
      smc_categorical_logit_glm(y_op.handle(), y_op.scalar(), x_handle(x),
                                alpha_val.data(), beta_val.data(), C, flags, &logp,
                                d_alpha.data(), d_beta.data(),
                                dx_handle<T_x>(partials<0>(ops_partials))));

  if constexpr (!is_constant_all<T_alpha>::value) {  // L160-169
    partials<1>(ops_partials) = d_alpha;
  }
  if constexpr (!is_constant_all<T_beta>::value) {  // L171-190
    partials<2>(ops_partials) = d_beta;
  }
  return ops_partials.build(logp);
}

/** alpha and / or beta on the device (the OpenCL overloads' signature): C + K C doubles
 * come to the host, see cuda_internal::host_param. */
template <bool propto, typename T_y, typename T_x, typename T_alpha,
          typename T_beta, require_cuda_design_matrix_t<T_x>* = nullptr,
          require_any_t<is_cuda_operand<T_alpha>, is_cuda_operand<T_beta>>* = nullptr>
return_type_t<T_x, T_alpha, T_beta> categorical_logit_glm_lpmf(
    const T_y& y, const T_x& x, const T_alpha& alpha, const T_beta& beta) {
  return categorical_logit_glm_lpmf<propto>(y, x, cuda_internal::host_param(alpha),
                                            cuda_internal::host_param_matrix(beta));
}

// The propto = false forwarding overload is the reference's own
// (prim/prob/categorical_logit_glm_lpmf.hpp L197-201).

}  // namespace math
}  // namespace stan
#endif
