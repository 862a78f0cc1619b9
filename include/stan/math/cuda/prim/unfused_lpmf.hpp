#ifndef STAN_MATH_CUDA_PRIM_UNFUSED_LPMF_HPP
#define STAN_MATH_CUDA_PRIM_UNFUSED_LPMF_HPP
// The un-fused densities on a device-resident linear predictor (SURVEY.md
// 8(f)3): bernoulli_logit_lpmf, poisson_log_lpmf, neg_binomial_2_log_lpmf,
// ordered_logistic_lpmf, categorical_logit_lpmf and normal_lpdf for a theta that is a matrix_cuda<double>
// or a var_value<matrix_cuda<double>> -- the B200 overloads of
//   prim/prob/bernoulli_logit_lpmf.hpp L33-98      (opencl/prim/bernoulli_logit_lpmf.hpp)
//   prim/prob/poisson_log_lpmf.hpp L27-100         (opencl/prim/poisson_log_lpmf.hpp)
//   prim/prob/neg_binomial_2_log_lpmf.hpp L24-134  (opencl/prim/neg_binomial_2_log_lpmf.hpp)
//   prim/prob/ordered_logistic_lpmf.hpp L72-214    (opencl/prim/ordered_logistic_lpmf.hpp)
//   prim/prob/categorical_logit_lpmf.hpp L16-32    (no OpenCL twin; row-wise form)
//   prim/prob/normal_lpdf.hpp L41-104              (opencl/prim/normal_lpdf.hpp)
// for models that add terms to x * beta before the likelihood.  Same names,
// template order and <propto>; value and d/dtheta come from one kernel over the
// N-vector and are attached through make_partials_propagator(...).build(logp).
// The prim templates are switched off for device operands by the reference's own
// kernel-expression trait (see matrix_cuda.hpp).
#include <stan/math/cuda/prim/glm_common.hpp>
#include <stan/math/cuda/prim/normal_id_glm_lpdf.hpp>

namespace stan {
namespace math {

namespace cuda_internal {
/** Sizes of the random variable and the parameter agree (check_consistent_sizes). */
template <typename T_n, typename T_theta>
inline void check_rv_size(const char* function, const T_n& n, const T_theta& theta) {
  if (!is_stan_scalar<T_n>::value) {
    check_size_match(function, "Size of ", "Random variable", operand_size(n),
                     "size of ", "parameter", operand_size(theta));
  }
}
/** Device handle of the partials of a device-var edge, else NULL. */
template <typename T, typename Edge>
inline smc_matrix* dvec_handle(Edge& edge_partials) {
  if constexpr (is_var_matrix_cuda<T>::value) {
    return edge_partials.handle();
  } else {
    return nullptr;
  }
}
/** A scalar parameter next to a device random variable (the reference's device overloads
 * broadcast it, e.g. opencl/prim/bernoulli_logit_lpmf.hpp L25-86 with a scalar theta):
 * n copies on the device; a var collects the sum of the adjoints in the reverse sweep. */
template <typename T, require_arithmetic_t<T>* = nullptr>
inline matrix_cuda<double> broadcast_to_device(T v, int64_t n) {
  matrix_cuda<double> m(n, 1);
  if (n > 0) {
    check_cuda_status("broadcast_to_device", smc_matrix_zero(m.handle()));
    check_cuda_status("broadcast_to_device",
                      smc_matrix_add_scalar(m.handle(), static_cast<double>(v)));
  }
  return m;
}
inline var_value<matrix_cuda<double>> broadcast_to_device(const var& v, int64_t n) {
  var_value<matrix_cuda<double>> res(broadcast_to_device(v.val(), n));
  reverse_pass_callback([v, res]() mutable {
    double sum = 0;
    if (res.size() > 0) {
      check_cuda_status("broadcast_to_device(var)",
                        smc_vector_sum(res.adj().handle(), &sum));
    }
    v.adj() += sum;
  });
  return res;
}
}  // namespace cuda_internal

template <bool propto, typename T_n, typename T_prob,
          require_t<is_cuda_operand<T_prob>>* = nullptr>
return_type_t<T_prob> bernoulli_logit_lpmf(const T_n& n, const T_prob& theta) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "bernoulli_logit_lpmf(CUDA)";
  check_rv_size(function, n, theta);
  if (theta.size() == 0 || operand_size(n) == 0) {
    return 0.0;
  }
  row_operand<int, T_n> n_op(n);
  auto ops_partials = make_partials_propagator(theta);
  double logp = 0;
  const unsigned flags = (propto ? SMC_PROPTO : 0u) | var_flag<T_prob>(SMC_VAR_ALPHA);
  check_cuda_status(function,
                    smc_bernoulli_logit_lpmf(n_op.handle(), n_op.scalar(),
                                             x_handle(theta), flags, &logp,
                                             dvec_handle<T_prob>(partials<0>(ops_partials))));
  if (!include_summand<propto, T_prob>::value) {
    return 0.0;
  }
  return ops_partials.build(logp);
}

/** device n, scalar theta */
template <bool propto, typename T_n, typename T_prob,
          require_t<is_cuda_operand<T_n>>* = nullptr,
          require_stan_scalar_t<T_prob>* = nullptr>
return_type_t<T_prob> bernoulli_logit_lpmf(const T_n& n, const T_prob& theta) {
  return bernoulli_logit_lpmf<propto>(
      n, cuda_internal::broadcast_to_device(theta, n.size()));
}

template <bool propto, typename T_n, typename T_log_rate,
          require_t<is_cuda_operand<T_log_rate>>* = nullptr>
return_type_t<T_log_rate> poisson_log_lpmf(const T_n& n, const T_log_rate& alpha) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "poisson_log_lpmf(CUDA)";
  check_rv_size(function, n, alpha);
  row_operand<int, T_n> n_op(n);
  if (n_op.handle() == nullptr) {
    check_nonnegative(function, "Random variable", n_op.scalar());
  }
  if (alpha.size() == 0 || operand_size(n) == 0) {
    return 0.0;
  }
  auto ops_partials = make_partials_propagator(alpha);
  double logp = 0;
  const unsigned flags
      = (propto ? SMC_PROPTO : 0u) | var_flag<T_log_rate>(SMC_VAR_ALPHA);
  check_cuda_status(function,
                    smc_poisson_log_lpmf(n_op.handle(), n_op.scalar(), x_handle(alpha),
                                         flags, &logp,
                                         dvec_handle<T_log_rate>(partials<0>(ops_partials))));
  if (!include_summand<propto, T_log_rate>::value) {
    return 0.0;
  }
  return ops_partials.build(logp);
}

/** device n, scalar log rate */
template <bool propto, typename T_n, typename T_log_rate,
          require_t<is_cuda_operand<T_n>>* = nullptr,
          require_stan_scalar_t<T_log_rate>* = nullptr>
return_type_t<T_log_rate> poisson_log_lpmf(const T_n& n, const T_log_rate& alpha) {
  return poisson_log_lpmf<propto>(n,
                                  cuda_internal::broadcast_to_device(alpha, n.size()));
}

/** eta on the device; phi an arithmetic / var scalar or a device vector (data or var). */
template <bool propto, typename T_n, typename T_log_location, typename T_precision,
          require_t<is_cuda_operand<T_log_location>>* = nullptr,
          require_any_t<is_stan_scalar<T_precision>,
                        is_cuda_operand<T_precision>>* = nullptr>
return_type_t<T_log_location, T_precision> neg_binomial_2_log_lpmf(
    const T_n& n, const T_log_location& eta, const T_precision& phi) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "neg_binomial_2_log_lpmf(CUDA)";
  check_rv_size(function, n, eta);
  if (!is_stan_scalar<T_precision>::value) {
    check_size_match(function, "Size of ", "Log location parameter", operand_size(eta),
                     "size of ", "Precision parameter", operand_size(phi));
  }
  row_operand<int, T_n> n_op(n);
  if (n_op.handle() == nullptr) {
    check_nonnegative(function, "Failures variable", n_op.scalar());
  }
  row_operand<double, T_precision> phi_op(phi);
  if (phi_op.handle() == nullptr) {  // (a device phi is checked by the call)
    check_positive_finite(function, "Precision parameter", phi_op.scalar());
  }
  if (eta.size() == 0 || operand_size(n) == 0) {
    return 0.0;
  }
  auto ops_partials = make_partials_propagator(eta, phi);
  double logp = 0, d_phi = 0;
  const unsigned flags = (propto ? SMC_PROPTO : 0u)
                         | var_flag<T_log_location>(SMC_VAR_ALPHA)
                         | var_flag<T_precision>(SMC_VAR_AUX);
  check_cuda_status(
      function,
      smc_neg_binomial_2_log_lpmf(n_op.handle(), n_op.scalar(), x_handle(eta),
                                  phi_op.handle(), phi_op.scalar(), flags, &logp,
                                  dvec_handle<T_log_location>(partials<0>(ops_partials)),
                                  &d_phi,
                                  dvec_handle<T_precision>(partials<1>(ops_partials))));
  if (!include_summand<propto, T_log_location, T_precision>::value) {
    return 0.0;
  }
  if constexpr (!is_constant_all<T_precision>::value
                && is_stan_scalar<T_precision>::value) {
    partials<1>(ops_partials)[0] = d_phi;
  }
  return ops_partials.build(logp);
}

/** scalar eta next to a device n and / or a device phi: broadcast on the device */
template <bool propto, typename T_n, typename T_log_location, typename T_precision,
          require_stan_scalar_t<T_log_location>* = nullptr,
          require_any_t<is_cuda_operand<T_n>, is_cuda_operand<T_precision>>* = nullptr>
return_type_t<T_log_location, T_precision> neg_binomial_2_log_lpmf(
    const T_n& n, const T_log_location& eta, const T_precision& phi) {
  using namespace cuda_internal;  // NOLINT
  int64_t size = 0;
  if constexpr (is_cuda_operand<T_n>::value) {
    size = n.size();
  } else {
    size = phi.size();
  }
  return neg_binomial_2_log_lpmf<propto>(n, broadcast_to_device(eta, size), phi);
}

/** c: one host cut-point vector (Eigen column vector of double or var). */
template <bool propto, typename T_y, typename T_loc, typename T_cut,
          require_t<is_cuda_operand<T_loc>>* = nullptr,
          require_col_vector_t<T_cut>* = nullptr>
return_type_t<T_loc, T_cut> ordered_logistic_lpmf(const T_y& y, const T_loc& lambda,
                                                  const T_cut& c) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "ordered_logistic_lpmf(CUDA)";
  check_rv_size(function, y, lambda);
  row_operand<int, T_y> y_op(y);
  const Eigen::VectorXd cuts_val = host_values(c);
  const int64_t n_cuts = operand_size(c);
  auto ops_partials = make_partials_propagator(lambda, c);
  double logp = 0;
  Eigen::VectorXd d_cuts = Eigen::VectorXd::Zero(n_cuts);
  const unsigned flags = (propto ? SMC_PROPTO : 0u) | var_flag<T_loc>(SMC_VAR_ALPHA)
                         | var_flag<T_cut>(SMC_VAR_AUX);
  if (lambda.size() == 0) {
    return 0.0;
  }
  check_cuda_status(function,
                    smc_ordered_logistic_lpmf(y_op.handle(), y_op.scalar(),
                                              x_handle(lambda), cuts_val.data(), n_cuts,
                                              flags, &logp,
                                              dvec_handle<T_loc>(partials<0>(ops_partials)),
                                              d_cuts.data()));
  if (!include_summand<propto, T_loc, T_cut>::value) {
    return 0.0;
  }
  if constexpr (!is_constant_all<T_cut>::value) {
    store_host_partial<T_cut>(partials<1>(ops_partials), d_cuts.data(), n_cuts);
  }
  return ops_partials.build(logp);
}

/** cuts on the device (data or autodiff), as every argument of the OpenCL overload is
 * (opencl/prim/ordered_logistic_lpmf.hpp L68-160): a (C-1) x 1 matrix is one cut-point
 * vector for all outcomes, a (C-1) x N matrix holds one cut-point vector per outcome in
 * its columns -- prim's std::vector<Eigen::VectorXd> form (prim L72-200), which
 * to_matrix_cuda uploads that way.  The partial of the cut points is written on the
 * device, straight into the edge. */
template <bool propto, typename T_y, typename T_loc, typename T_cut,
          require_t<is_cuda_operand<T_loc>>* = nullptr,
          require_t<is_cuda_operand<T_cut>>* = nullptr>
return_type_t<T_loc, T_cut> ordered_logistic_lpmf(const T_y& y, const T_loc& lambda,
                                                  const T_cut& cuts) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "ordered_logistic_lpmf(CUDA)";
  check_rv_size(function, y, lambda);
  const int64_t N = lambda.size();
  if (cuts.cols() > 1) {
    check_size_match(function, "Length of location variables ", N,
                     "Number of cutpoint vectors ", cuts.cols());
  }
  row_operand<int, T_y> y_op(y);
  auto ops_partials = make_partials_propagator(lambda, cuts);
  double logp = 0;
  const unsigned flags = (propto ? SMC_PROPTO : 0u) | var_flag<T_loc>(SMC_VAR_ALPHA)
                         | var_flag<T_cut>(SMC_VAR_AUX);
  if (N == 0 || cuts.cols() == 0) {
    return 0.0;
  }
  check_cuda_status(function,
                    smc_ordered_logistic_lpmf_rows(
                        y_op.handle(), y_op.scalar(), x_handle(lambda), x_handle(cuts),
                        flags, &logp, dvec_handle<T_loc>(partials<0>(ops_partials)),
                        dvec_handle<T_cut>(partials<1>(ops_partials))));
  if (!include_summand<propto, T_loc, T_cut>::value) {
    return 0.0;
  }
  return ops_partials.build(logp);
}

/** normal_lpdf(y | mu, sigma) with y and / or mu on the device (data or autodiff)
 * and a host scalar sigma: prim/prob/normal_lpdf.hpp L41-104. */
template <bool propto, typename T_y, typename T_loc, typename T_scale,
          require_any_t<is_cuda_operand<T_y>, is_cuda_operand<T_loc>>* = nullptr,
          require_stan_scalar_t<T_scale>* = nullptr>
return_type_t<T_y, T_loc, T_scale> normal_lpdf(T_y&& y, T_loc&& mu, T_scale&& sigma) {
  using namespace cuda_internal;  // NOLINT
  using Ty = std::decay_t<T_y>;
  using Tm = std::decay_t<T_loc>;
  using Ts = std::decay_t<T_scale>;
  static constexpr const char* function = "normal_lpdf(CUDA)";
  if (!is_stan_scalar<Ty>::value && !is_stan_scalar<Tm>::value) {
    check_size_match(function, "Size of ", "Random variable", operand_size(y),
                     "size of ", "Location parameter", operand_size(mu));
  }
  static_assert(is_stan_scalar<Ty>::value || is_cuda_operand<Ty>::value,
                "normal_lpdf(CUDA): y is a scalar or a device vector");
  static_assert(is_stan_scalar<Tm>::value || is_cuda_operand<Tm>::value,
                "normal_lpdf(CUDA): mu is a scalar or a device vector");
  row_operand<double, Ty> y_op(y);
  row_operand<double, Tm> mu_op(mu);
  auto ops_partials = make_partials_propagator(y, mu, sigma);
  double logp = 0, d_y = 0, d_mu = 0, d_sigma = 0;
  const unsigned flags = (propto ? SMC_PROPTO : 0u) | var_flag<Ty>(SMC_VAR_Y)
                         | var_flag<Tm>(SMC_VAR_ALPHA) | var_flag<Ts>(SMC_VAR_AUX);
  check_cuda_status(
      function,
      smc_normal_lpdf(y_op.handle(), y_op.scalar(), mu_op.handle(), mu_op.scalar(),
                      value_of(sigma), flags, &logp,
                      dvec_handle<Ty>(partials<0>(ops_partials)), &d_y,
                      dvec_handle<Tm>(partials<1>(ops_partials)), &d_mu, &d_sigma));
  if (operand_size(y) == 0 || operand_size(mu) == 0
      || !include_summand<propto, Ty, Tm, Ts>::value) {
    return 0.0;
  }
  if constexpr (!is_constant_all<Ty>::value && is_stan_scalar<Ty>::value) {
    partials<0>(ops_partials)[0] = d_y;
  }
  if constexpr (!is_constant_all<Tm>::value && is_stan_scalar<Tm>::value) {
    partials<1>(ops_partials)[0] = d_mu;
  }
  if constexpr (!is_constant_all<Ts>::value) {
    partials<2>(ops_partials)[0] = d_sigma;
  }
  return ops_partials.build(logp);
}

/** normal_lpdf with a per-row scale on the device (data or autodiff): the linear-
 * regression GLM with no attributes -- normal_id_glm_lpdf(y | x = N x 0, alpha = mu,
 * beta = [], sigma) is the same density, partials and constant terms
 * (prim/prob/normal_id_glm_lpdf.hpp L122-213 with x beta = 0), so the fused entry takes it. */
template <bool propto, typename T_y, typename T_loc, typename T_scale,
          require_t<is_cuda_operand<T_scale>>* = nullptr>
return_type_t<T_y, T_loc, T_scale> normal_lpdf(const T_y& y, const T_loc& mu,
                                              const T_scale& sigma) {
  if (!is_stan_scalar<T_y>::value) {
    check_size_match("normal_lpdf(CUDA)", "Size of ", "Random variable",
                     cuda_internal::operand_size(y), "size of ", "Scale parameter",
                     sigma.size());
  }
  const matrix_cuda<double> no_attributes(sigma.size(), 0);
  return normal_id_glm_lpdf<propto>(y, no_attributes, mu, Eigen::VectorXd(0), sigma);
}

/** categorical_logit_lpmf with one row of log odds per outcome: `lin` is an N x C
 * device matrix (data or autodiff) and the result is
 * sum_i categorical_logit_lpmf(ns[i] | lin.row(i)^T) of
 * prim/prob/categorical_logit_lpmf.hpp L16-32 -- what a model that adds terms to
 * x * beta writes as a loop over the rows.  (The reference's own signatures take ONE
 * column vector of log odds for every outcome; a C-vector is not worth a device.)
 * d/dlin = one-hot(ns) - softmax(lin) is written straight into the device edge. */
template <bool propto, typename T_n, typename T_prob,
          require_t<is_cuda_operand<T_prob>>* = nullptr>
return_type_t<T_prob> categorical_logit_lpmf(const T_n& ns, const T_prob& lin) {
  using namespace cuda_internal;  // NOLINT
  static constexpr const char* function = "categorical_logit_lpmf(CUDA)";
  if (!is_stan_scalar<T_n>::value) {
    check_size_match(function, "Size of ", "Random variable", operand_size(ns),
                     "rows of ", "log odds parameter", lin.rows());
  }
  row_operand<int, T_n> n_op(ns);
  auto ops_partials = make_partials_propagator(lin);
  double logp = 0;
  const unsigned flags = (propto ? SMC_PROPTO : 0u) | var_flag<T_prob>(SMC_VAR_ALPHA);
  check_cuda_status(function,
                    smc_categorical_logit_lpmf(n_op.handle(), n_op.scalar(), x_handle(lin),
                                               flags, &logp,
                                               dvec_handle<T_prob>(partials<0>(ops_partials))));
  if (!include_summand<propto, T_prob>::value || lin.rows() == 0) {
    return 0.0;
  }
  return ops_partials.build(logp);
}

template <typename T_n, typename T_prob, require_t<is_cuda_operand<T_prob>>* = nullptr>
inline return_type_t<T_prob> categorical_logit_lpmf(const T_n& ns, const T_prob& lin) {
  return categorical_logit_lpmf<false>(ns, lin);
}

// The other propto = false forwarding overloads are the reference's own (last lines of
// each prim/prob/*_lpmf.hpp): their <false> calls resolve to the overloads above.

}  // namespace math
}  // namespace stan
#endif
