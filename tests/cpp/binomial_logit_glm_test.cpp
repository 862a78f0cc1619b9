// CUDA-vs-CPU parity for binomial_logit_glm_lpmf; cases follow the reference's
// device test test/unit/math/opencl/rev/binomial_logit_glm_lpmf_test.cpp
// (error_checking, small_simple, broadcast_n, zero_instances, zero_attributes,
// small_vector_alpha, big) plus the shapes the fused kernel splits on.
#include "cuda_test_util.hpp"

using Eigen::Dynamic;
using Eigen::Matrix;
using Eigen::MatrixXd;
using Eigen::VectorXd;
using stan::math::matrix_cuda;
using stan::math::var;
using std::vector;
using namespace cuda_test;  // NOLINT

namespace {
auto f = [](const auto& n, const auto& trials, const auto& x, const auto& alpha,
            const auto& beta) {
  return stan::math::binomial_logit_glm_lpmf(n, trials, x, alpha, beta);
};
auto f_propto = [](const auto& n, const auto& trials, const auto& x,
                   const auto& alpha, const auto& beta) {
  return stan::math::binomial_logit_glm_lpmf<true>(n, trials, x, alpha, beta);
};
}  // namespace

TEST(CudaBinomialLogitGLM, error_checking) {
  int N = 3, M = 2;
  vector<int> n{1, 0, 1}, n_size{1, 0, 1, 0}, n_value{0, 1, -23};
  vector<int> trials{10, 5, 2}, trials_size{1, 0, 1, 0}, trials_value{5, 1, -1};
  MatrixXd x(N, M), x_size1(N - 1, M), x_size2(N, M - 1), x_value(N, M);
  x << -12, 46, -42, 24, 25, 27;
  x_size1 << -12, 46, -42, 24;
  x_size2 << -12, 46, -42;
  x_value << -12, 46, -42, 24, 25, NAN;
  VectorXd beta(M), beta_size(M + 1), beta_value(M);
  beta << 0.3, 2;
  beta_size << 0.3, 2, 0.4;
  beta_value << 0.3, INFINITY;
  VectorXd alpha(N), alpha_size(N - 1), alpha_value(N);
  alpha << 0.3, -0.8, 1.8;
  alpha_size << 0.3, -0.8;
  alpha_value << 0.3, -0.8, NAN;

  matrix_cuda<double> x_d(x), x_size1_d(x_size1), x_size2_d(x_size2), x_value_d(x_value);
  matrix_cuda<int> n_d(n), n_size_d(n_size), n_value_d(n_value);
  matrix_cuda<int> t_d(trials), t_size_d(trials_size), t_value_d(trials_value);
  matrix_cuda<double> alpha_d(alpha), alpha_size_d(alpha_size), alpha_value_d(alpha_value);

  using stan::math::binomial_logit_glm_lpmf;
  EXPECT_NO_THROW(binomial_logit_glm_lpmf(n_d, t_d, x_d, alpha_d, beta));
  EXPECT_THROW(binomial_logit_glm_lpmf(n_size_d, t_d, x_d, alpha_d, beta), std::invalid_argument);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_d, t_size_d, x_d, alpha_d, beta), std::invalid_argument);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_d, t_d, x_size1_d, alpha_d, beta), std::invalid_argument);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_d, t_d, x_size2_d, alpha_d, beta), std::invalid_argument);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_d, t_d, x_d, alpha_size_d, beta), std::invalid_argument);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_d, t_d, x_d, alpha_d, beta_size), std::invalid_argument);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_value_d, t_d, x_d, alpha_d, beta), std::domain_error);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_d, t_value_d, x_d, alpha_d, beta), std::domain_error);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_d, t_d, x_value_d, alpha_d, beta), std::domain_error);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_d, t_d, x_d, alpha_value_d, beta), std::domain_error);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_d, t_d, x_d, alpha_d, beta_value), std::domain_error);
  // host per-row containers take the same route
  EXPECT_THROW(binomial_logit_glm_lpmf(n_size, trials, x_d, alpha, beta), std::invalid_argument);
  EXPECT_THROW(binomial_logit_glm_lpmf(n_value, trials, x_d, alpha, beta), std::domain_error);
  EXPECT_THROW(binomial_logit_glm_lpmf(7, 5, x_d, alpha, beta), std::domain_error);
}

TEST(CudaBinomialLogitGLM, small_simple) {
  vector<int> n{0, 1, 0}, trials{10, 4, 15};
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST, HOST), n, trials, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, HOST, HOST), n, trials, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, HOST, DEV, HOST, HOST), n, trials, x, alpha, beta);
}

TEST(CudaBinomialLogitGLM, moderate_theta) {
  // |theta| = O(1): both log_inv_logit branches away from saturation
  vector<int> n{3, 1, 7, 0, 5}, trials{10, 4, 15, 3, 5};
  MatrixXd x(5, 2);
  x << -1.2, 0.46, -0.42, 2.4, 0.25, 0.27, 0.9, -1.1, 0.0, 0.3;
  VectorXd beta(2);
  beta << 0.3, -0.7;
  double alpha = -0.2;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST, HOST), n, trials, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, HOST, HOST), n, trials, x, alpha, beta);
}

TEST(CudaBinomialLogitGLM, broadcast_n_and_trials) {
  vector<int> trials{10, 4, 15}, n{1, 0, 2};
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, DEV, HOST, HOST), 1, trials, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(HOST, DEV, DEV, HOST, HOST), 1, trials, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, HOST, DEV, HOST, HOST), n, 6, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, HOST, DEV, HOST, HOST), 2, 6, x, alpha, beta);
}

TEST(CudaBinomialLogitGLM, zero_instances) {
  vector<int> n{}, trials{};
  MatrixXd x(0, 2);
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST, HOST), n, trials, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, HOST, HOST), n, trials, x, alpha, beta);
}

TEST(CudaBinomialLogitGLM, zero_attributes) {
  vector<int> n{0, 1, 0}, trials{10, 5, 4};
  MatrixXd x(3, 0);
  VectorXd beta(0);
  double alpha = 0.3;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST, HOST), n, trials, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, HOST, HOST), n, trials, x, alpha, beta);
}

TEST(CudaBinomialLogitGLM, small_vector_alpha) {
  vector<int> n{0, 1, 0}, trials{0, 1, 0};
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  VectorXd beta(2);
  beta << 0.3, 2;
  VectorXd alpha(3);
  alpha << 0.3, -0.8, 1.8;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, DEV, HOST), n, trials, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, DEV, HOST), n, trials, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST, HOST), n, trials, x, alpha, beta);
}

TEST(CudaBinomialLogitGLM, big) {
  int N = 153, M = 71;  // deliberately not multiples of any tile size
  srand(1);
  vector<int> n(N), trials(N);
  for (int i = 0; i < N; i++) {
    trials[i] = std::abs(Eigen::Array<int, 1, 1>::Random()[0]) % 50;
    n[i] = trials[i] ? std::abs(Eigen::Array<int, 1, 1>::Random()[0]) % (trials[i] + 1) : 0;
  }
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M);
  VectorXd alpha = VectorXd::Random(N);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, DEV, HOST), n, trials, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, DEV, HOST), n, trials, x, alpha, beta);
}

TEST(CudaBinomialLogitGLM, wide_rows_more_than_one_tile) {
  // N spans several row tiles, K > 256 takes the general two-pass path
  srand(2);
  for (int M : {64, 256, 300}) {
    int N = 4099;
    vector<int> n(N), trials(N);
    for (int i = 0; i < N; i++) {
      trials[i] = 1 + (i * 13 + M) % 20;
      n[i] = (i * 7 + M) % (trials[i] + 1);
    }
    MatrixXd x = MatrixXd::Random(N, M);
    VectorXd beta = VectorXd::Random(M) / std::sqrt(M);
    double alpha = 0.1;
    compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST, HOST), n, trials, x, alpha, beta);
  }
}

TEST(CudaBinomialLogitGLM, interface_types) {
  vector<int> n{1, 0, 3}, trials{4, 2, 3};
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> n_d(n), t_d(trials);
  VectorXd beta(2);
  beta << 0.3, 2;
  const double expect = stan::math::binomial_logit_glm_lpmf(n, trials, x, 0.3, beta);
  using stan::math::binomial_logit_glm_lpmf;
  EXPECT_NEAR(binomial_logit_glm_lpmf(n_d, t_d, x_d, 0.3, beta), expect, 1e-12);
  Eigen::RowVectorXd beta_row = beta.transpose();
  EXPECT_NEAR(binomial_logit_glm_lpmf(n_d, t_d, x_d, 0.3, beta_row), expect, 1e-12);
  vector<double> beta_std{0.3, 2};
  EXPECT_NEAR(binomial_logit_glm_lpmf(n_d, t_d, x_d, 0.3, beta_std), expect, 1e-12);
  {
    vector<var> b{0.3, 2};
    var lp = binomial_logit_glm_lpmf(n_d, t_d, x_d, var(0.3), b);
    vector<var> b2{0.3, 2};
    var lp2 = binomial_logit_glm_lpmf(n, trials, x, var(0.3), b2);
    (lp + lp2).grad();
    EXPECT_NEAR(lp.val(), expect, 1e-12);
    for (int k = 0; k < 2; ++k) EXPECT_NEAR(b[k].adj(), b2[k].adj(), 1e-10);
    stan::math::recover_memory();
  }
  {
    stan::math::var_value<VectorXd> b(beta), b2(beta);
    var lp = binomial_logit_glm_lpmf(n_d, t_d, x_d, 0.3, b);
    var lp2 = binomial_logit_glm_lpmf(n, trials, x, 0.3, b2);
    (lp + lp2).grad();
    EXPECT_NEAR(lp.val(), expect, 1e-12);
    for (int k = 0; k < 2; ++k) EXPECT_NEAR(b.adj()[k], b2.adj()[k], 1e-10);
    stan::math::recover_memory();
  }
  {  // x as a device autodiff variable: d_x stays in HBM
    stan::math::var_value<matrix_cuda<double>> xv = stan::math::to_matrix_cuda(
        Matrix<var, Dynamic, Dynamic>(x));
    Matrix<var, Dynamic, Dynamic> xh = x;
    var lp = binomial_logit_glm_lpmf(n_d, t_d, xv, 0.3, beta);
    var lp2 = binomial_logit_glm_lpmf(n, trials, xh, 0.3, beta);
    (lp + lp2).grad();
    EXPECT_NEAR(lp.val(), lp2.val(), 1e-12);
    MatrixXd dx = stan::math::from_matrix_cuda(xv.adj().to_matrix_cuda());
    for (int j = 0; j < 2; ++j)
      for (int i = 0; i < 3; ++i) EXPECT_NEAR(dx(i, j), xh(i, j).adj(), 1e-10);
    stan::math::recover_memory();
  }
}
