#ifndef STAN_MATH_CUDA_PRIM_ORDERED_LOGISTIC_GLM_LPMF_HPP
#define STAN_MATH_CUDA_PRIM_ORDERED_LOGISTIC_GLM_LPMF_HPP
// ordered_logistic_glm_lpmf for a device-resident design matrix: the B200
// overload of stan/math/prim/prob/ordered_logistic_glm_lpmf.hpp L46-210 (same
// name, template order and <propto> convention; cf.
// opencl/prim/ordered_logistic_glm_lpmf.hpp L53-61).  One fused pass over x
// (smc_ordered_logistic_glm) yields the value, d_beta, d_cuts and d_x.
#include <stan/math/cuda/prim/glm_common.hpp>
#include <stan/math/prim/err/check_ordered.hpp>

namespace stan {
namespace math {

template <bool propto, typename T_y, typename T_x, typename T_beta,
          typename T_cuts, require_cuda_design_matrix_t<T_x>* = nullptr,
          require_all_col_vector_t<T_beta, T_cuts>* = nullptr>
return_type_t<T_x, T_beta, T_cuts> ordered_logistic_glm_lpmf(
    const T_y& y, const T_x& x, const T_beta& beta, const T_cuts& cuts) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "ordered_logistic_glm_lpmf(CUDA)";
  const int64_t N = x.rows();
  const int64_t K = x.cols();
  const int64_t n_cuts = operand_size(cuts);
  const int64_t N_classes = n_cuts + 1;

  // prim L75-77
  if (!is_stan_scalar<T_y>::value) {
    check_size_match(function, "Rows of ", "x", N, "rows of ", "y", operand_size(y));
  }
  check_size_match(function, "Columns of ", "x", K, "size of ", "beta",
                   operand_size(beta));
  row_operand<int, T_y> y_op(y, x_handle(x));
  const Eigen::VectorXd cuts_val = host_values(cuts);
  // L82-89; the range of a device y is checked by the call itself (y is data:
  // the library caches its min / max at upload)
  int y_lo = 1, y_hi = 1;
  if (y_op.handle() == nullptr) {
    y_lo = y_hi = y_op.scalar();
  } else if (N > 0) {
    check_cuda_status(function, smc_matrix_int_range(y_op.handle(), &y_lo, &y_hi));
  }
  if (N > 0 || is_stan_scalar<T_y>::value) {
    check_bounded(function, "Vector of dependent variables", y_lo, 1, N_classes);
    check_bounded(function, "Vector of dependent variables", y_hi, 1, N_classes);
  }
  check_ordered(function, "Cut-points", cuts_val);
  if (N_classes > 1) {
    if (N_classes > 2) {
      check_finite(function, "Final cut-point", cuts_val[N_classes - 2]);
    }
    check_finite(function, "First cut-point", cuts_val[0]);
  }
  if (N == 0 || n_cuts == 0) {  // size_zero(y, cuts), L91-93
    return 0;
  }
  if (!include_summand<propto, T_x, T_beta, T_cuts>::value) {  // L94-96
    return 0;
  }

  const Eigen::VectorXd beta_val = host_values(beta);
  auto ops_partials = make_partials_propagator(x, beta, cuts);

  const unsigned flags = (propto ? SMC_PROPTO : 0u) | dx_flags<T_x>()
                         | var_flag<T_beta>(SMC_VAR_BETA)
                         | var_flag<T_cuts>(SMC_VAR_AUX);
  double logp = 0;
  Eigen::VectorXd d_beta(K), d_cuts(n_cuts);
  check_cuda_status(
      function,
      smc_ordered_logistic_glm(y_op.handle(), y_op.scalar(), x_handle(x),
                               beta_val.data(), cuts_val.data(), n_cuts, flags,
                               &logp, d_beta.data(), d_cuts.data(),
                               dx_factor_handle<T_x>(partials<0>(ops_partials), beta_val.data())));

  if constexpr (!is_constant_all<T_beta>::value) {  // L185-195
    store_host_partial<T_beta>(partials<1>(ops_partials), d_beta.data(), K);
  }
  if constexpr (!is_constant_all<T_cuts>::value) {  // L197-207
    store_host_partial<T_cuts>(partials<2>(ops_partials), d_cuts.data(), n_cuts);
  }
  return ops_partials.build(logp);
}

/** beta and / or the cut points on the device (the OpenCL overloads' signature): they
 * come to the host, see cuda_internal::host_param. */
template <bool propto, typename T_y, typename T_x, typename T_beta,
          typename T_cuts, require_cuda_design_matrix_t<T_x>* = nullptr,
          require_any_t<is_cuda_operand<T_beta>, is_cuda_operand<T_cuts>>* = nullptr>
return_type_t<T_x, T_beta, T_cuts> ordered_logistic_glm_lpmf(
    const T_y& y, const T_x& x, const T_beta& beta, const T_cuts& cuts) {
  return ordered_logistic_glm_lpmf<propto>(y, x, cuda_internal::host_param(beta),
                                           cuda_internal::host_param(cuts));
}

// The propto = false forwarding overload is the reference's own
// (prim/prob/ordered_logistic_glm_lpmf.hpp L212-216).

}  // namespace math
}  // namespace stan
#endif
