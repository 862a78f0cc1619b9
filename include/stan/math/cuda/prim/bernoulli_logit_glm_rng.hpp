#ifndef STAN_MATH_CUDA_PRIM_BERNOULLI_LOGIT_GLM_RNG_HPP
#define STAN_MATH_CUDA_PRIM_BERNOULLI_LOGIT_GLM_RNG_HPP
// bernoulli_logit_glm_rng for a device-resident design matrix (SURVEY.md
// 8(f)4): the B200 overload of prim/prob/bernoulli_logit_glm_rng.hpp L42-88 for
// posterior-predictive draws.  The N x K product x * beta -- all of the
// arithmetic -- and the finiteness check of x run on the GPU (one sweep over x
// each); the draws themselves consume the caller's generator one row at a time
// in row order exactly as prim does (L76-81), so for the same seed the variates
// are the ones prim returns.
#include <stan/math/cuda/rev/multiply.hpp>
#include <stan/math/prim/fun/inv_logit.hpp>
#include <stan/math/prim/meta/VectorBuilder.hpp>
#include <boost/random/bernoulli_distribution.hpp>
#include <boost/random/variate_generator.hpp>

namespace stan {
namespace math {

template <typename T_alpha, typename T_beta, class RNG>
inline typename VectorBuilder<true, int, T_alpha>::type bernoulli_logit_glm_rng(
    const matrix_cuda<double>& x, const T_alpha& alpha, const T_beta& beta, RNG& rng) {
  using boost::bernoulli_distribution;
  using boost::variate_generator;
  static constexpr const char* function = "bernoulli_logit_glm_rng(CUDA)";
  const size_t N = x.cols();
  const size_t M = x.rows();
  check_consistent_size(function, "Weight vector", beta, N);       // L55
  check_consistent_size(function, "Vector of intercepts", alpha, M);  // L56
  int x_finite = 1;
  check_cuda_status(function, smc_matrix_all_finite(x.handle(), &x_finite));
  if (!x_finite) {  // L60
    throw_domain_error(function, "Matrix of independent variables", "", "",
                       "is not finite");
  }
  check_finite(function, "Weight vector", beta);  // L61
  check_finite(function, "Intercept", alpha);     // L62

  // x * beta on the device (a scalar beta multiplies every column, L67-69)
  Eigen::VectorXd beta_vector(N);
  if constexpr (is_vector<T_beta>::value) {
    beta_vector = cuda_internal::host_values(beta);
  } else {
    beta_vector.setConstant(value_of(beta));
  }
  Eigen::VectorXd x_beta = Eigen::VectorXd::Zero(M);
  if (M > 0) {
    x_beta = from_matrix_cuda<Eigen::VectorXd>(multiply(x, beta_vector));
  }

  scalar_seq_view<T_alpha> alpha_vec(alpha);
  VectorBuilder<true, int, T_alpha> output(M);
  for (size_t m = 0; m < M; ++m) {  // L76-81
    double theta_m = alpha_vec[m] + x_beta(m);
    variate_generator<RNG&, bernoulli_distribution<>> bernoulli_rng(
        rng, bernoulli_distribution<>(inv_logit(theta_m)));
    output[m] = bernoulli_rng();
  }
  return output.data();
}

}  // namespace math
}  // namespace stan
#endif
