// CUDA-vs-CPU parity for categorical_logit_glm_lpmf; cases follow the
// reference's device test
// test/unit/math/opencl/rev/categorical_logit_glm_lpmf_test.cpp (error_checking,
// small_simple, broadcast_y, zero_instances, zero_attributes, single_class, big
// with C = 43) plus the known answer of SURVEY.md 8(c).
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
auto f = [](const auto& y, const auto& x, const auto& alpha, const auto& beta) {
  return stan::math::categorical_logit_glm_lpmf(y, x, alpha, beta);
};
auto f_propto = [](const auto& y, const auto& x, const auto& alpha, const auto& beta) {
  return stan::math::categorical_logit_glm_lpmf<true>(y, x, alpha, beta);
};
}  // namespace

TEST(CudaCategoricalLogitGLM, error_checking) {
  int N = 3, M = 2, C = 3;
  vector<int> y{1, 3, 2}, y_size{1, 3, 1, 2}, y_value{1, 2, -23};
  MatrixXd x(N, M), x_size1(N - 1, M), x_size2(N, M - 1), x_value(N, M);
  x << -12, 46, -42, 24, 25, 27;
  x_size1 << -12, 46, -42, 24;
  x_size2 << -12, 46, -42;
  x_value << -12, 46, -42, 24, 25, -INFINITY;
  MatrixXd beta(M, C), beta_size1(M + 1, C), beta_size2(M, C + 1), beta_value(M, C);
  beta << 0.3, 2, 0.4, -0.1, -1.3, 1;
  beta_size1 << 0.3, 2, 0.4, -0.1, -1.3, 1, 0, 0, 0;
  beta_size2 << 0.3, 2, 0.4, -0.1, -1.3, 1, 0, 0;
  beta_value << 0.3, 2, 0.4, -0.1, -1.3, NAN;
  VectorXd alpha(C), alpha_size(C - 1), alpha_value(C);
  alpha << 0.3, -0.8, 1.8;
  alpha_size << 0.3, -0.8;
  alpha_value << 0.3, -0.8, NAN;

  matrix_cuda<double> x_d(x), x_size1_d(x_size1), x_size2_d(x_size2), x_value_d(x_value);
  matrix_cuda<int> y_d(y), y_size_d(y_size), y_value_d(y_value);

  using stan::math::categorical_logit_glm_lpmf;
  EXPECT_NO_THROW(categorical_logit_glm_lpmf(y_d, x_d, alpha, beta));
  EXPECT_THROW(categorical_logit_glm_lpmf(y_size_d, x_d, alpha, beta), std::invalid_argument);
  EXPECT_THROW(categorical_logit_glm_lpmf(y_d, x_size1_d, alpha, beta), std::invalid_argument);
  EXPECT_THROW(categorical_logit_glm_lpmf(y_d, x_size2_d, alpha, beta), std::invalid_argument);
  EXPECT_THROW(categorical_logit_glm_lpmf(y_d, x_d, alpha_size, beta), std::invalid_argument);
  EXPECT_THROW(categorical_logit_glm_lpmf(y_d, x_d, alpha, beta_size1), std::invalid_argument);
  EXPECT_THROW(categorical_logit_glm_lpmf(y_d, x_d, alpha, beta_size2), std::invalid_argument);
  EXPECT_THROW(categorical_logit_glm_lpmf(y_value_d, x_d, alpha, beta), std::domain_error);
  EXPECT_THROW(categorical_logit_glm_lpmf(y_d, x_value_d, alpha, beta), std::domain_error);
  EXPECT_THROW(categorical_logit_glm_lpmf(y_d, x_d, alpha_value, beta), std::domain_error);
  EXPECT_THROW(categorical_logit_glm_lpmf(y_d, x_d, alpha, beta_value), std::domain_error);
  EXPECT_THROW(categorical_logit_glm_lpmf(4, x_d, alpha, beta), std::domain_error);
}

TEST(CudaCategoricalLogitGLM, known_answer) {
  vector<int> y{1, 3, 1, 2, 2};
  MatrixXd x(5, 2);
  x << -12, 46, -42, 24, 25, 27, -14, -11, 5, 18;
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> y_d(y);
  Matrix<var, Dynamic, 1> alpha(3);
  alpha << 0.5, -2, 4;
  Matrix<var, Dynamic, Dynamic> beta(2, 3);
  beta << 0.3, 2, 0.4, -0.1, -1.3, 1;
  var lp = stan::math::categorical_logit_glm_lpmf(y_d, x_d, alpha, beta);
  lp.grad();
  EXPECT_NEAR(lp.val(), -141.10004744408855, 1e-10);
  const double da[3] = {1.0000474428564439, 1.9999979548657816, -3.0000453977222259};
  for (int c = 0; c < 3; ++c) EXPECT_NEAR(alpha[c].adj(), da[c], 1e-10);
  const double db[6] = {26.999335799326794, 83.999478127000543, -8.9999713681453422,
                        7.0000224964526279, -17.999364431181448, -90.999500623453187};
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < 2; ++k) EXPECT_NEAR(beta(k, c).adj(), db[c * 2 + k], 1e-9);
  stan::math::recover_memory();
}

TEST(CudaCategoricalLogitGLM, small_simple) {
  vector<int> y{1, 3, 2};
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  MatrixXd beta(2, 3);
  beta << 0.3, 2, 0.4, -0.1, -1.3, 1;
  VectorXd alpha(3);
  alpha << 0.3, -0.8, 1.8;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaCategoricalLogitGLM, broadcast_y) {
  int y = 2;
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  MatrixXd beta(2, 3);
  beta << 0.3, 2, 0.4, -0.1, -1.3, 1;
  VectorXd alpha(3);
  alpha << 0.3, -0.8, 1.8;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(HOST, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaCategoricalLogitGLM, zero_instances) {
  vector<int> y{};
  MatrixXd x(0, 2);
  MatrixXd beta(2, 3);
  beta << 0.3, 2, 0.4, -0.1, -1.3, 1;
  VectorXd alpha(3);
  alpha << 0.3, -0.8, 1.8;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaCategoricalLogitGLM, zero_attributes) {
  vector<int> y{1, 3, 2};
  MatrixXd x(3, 0);
  MatrixXd beta(0, 3);
  VectorXd alpha(3);
  alpha << 0.3, -0.8, 1.8;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaCategoricalLogitGLM, single_class) {
  vector<int> y{1, 1, 1};
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  MatrixXd beta(2, 1);
  beta << 0.3, 2;
  VectorXd alpha(1);
  alpha << 0.3;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaCategoricalLogitGLM, big) {
  int N = 153, M = 71, C = 43;
  srand(9);
  vector<int> y(N);
  for (int i = 0; i < N; i++) y[i] = std::abs(Eigen::Array<int, 1, 1>::Random()[0]) % C + 1;
  MatrixXd x = MatrixXd::Random(N, M);
  MatrixXd beta = MatrixXd::Random(M, C);
  VectorXd alpha = VectorXd::Random(C);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaCategoricalLogitGLM, more_than_64_classes) {
  // beyond one launch of the DMMA kernels: composed from class blocks of 64
  int N = 307, M = 23, C = 97;
  srand(12);
  vector<int> y(N);
  for (int i = 0; i < N; i++) y[i] = (i * 11) % C + 1;
  MatrixXd x = MatrixXd::Random(N, M);
  MatrixXd beta = MatrixXd::Random(M, C);
  VectorXd alpha = VectorXd::Random(C);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaCategoricalLogitGLM, config5a_shape) {
  // BASELINE.json configs[4] (categorical part) at reduced N: K = 512, C = 32
  int N = 4099, M = 512, C = 32;
  srand(10);
  vector<int> y(N);
  for (int i = 0; i < N; i++) y[i] = (i * 5) % C + 1;
  MatrixXd x = MatrixXd::Random(N, M);
  MatrixXd beta = MatrixXd::Random(M, C) / std::sqrt(M);
  VectorXd alpha = VectorXd::Random(C) * 0.1;
  matrix_cuda<int> y_d(y);
  matrix_cuda<double> x_d(x);
  Matrix<var, Dynamic, 1> a1 = alpha, a2 = alpha;
  Matrix<var, Dynamic, Dynamic> b1 = beta, b2 = beta;
  var lp_dev = stan::math::categorical_logit_glm_lpmf(y_d, x_d, a1, b1);
  var lp_cpu = stan::math::categorical_logit_glm_lpmf(y, x, a2, b2);
  (lp_dev + lp_cpu).grad();
  expect_close("logp", lp_dev.val(), lp_cpu.val(), kRelLogp, 0);
  compare_adj("d_alpha", a1, a2);
  compare_adj("d_beta", b1, b2);
  stan::math::recover_memory();
}
