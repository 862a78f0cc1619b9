// CUDA-vs-CPU parity for normal_id_glm_lpdf; cases follow the reference's device
// test test/unit/math/opencl/rev/normal_id_glm_lpdf_test.cpp (error_checking,
// small_simple, broadcast_y, zero_instances, zero_attributes,
// small_vector_alpha_sigma, big) plus the known answer of SURVEY.md 8(c).
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
auto f = [](const auto& y, const auto& x, const auto& alpha, const auto& beta,
            const auto& sigma) {
  return stan::math::normal_id_glm_lpdf(y, x, alpha, beta, sigma);
};
auto f_propto = [](const auto& y, const auto& x, const auto& alpha, const auto& beta,
                   const auto& sigma) {
  return stan::math::normal_id_glm_lpdf<true>(y, x, alpha, beta, sigma);
};
}  // namespace

TEST(CudaNormalIdGLM, error_checking) {
  int N = 3, M = 2;
  VectorXd y(N), y_size(N + 1), y_value(N);
  y << 14, 32, 21;
  y_size << 14, 32, 21, 3;
  y_value << 14, 32, NAN;
  MatrixXd x(N, M), x_size1(N - 1, M), x_size2(N, M - 1), x_value(N, M);
  x << -12, 46, -42, 24, 25, 27;
  x_size1 << -12, 46, -42, 24;
  x_size2 << -12, 46, -42;
  x_value << -12, 46, -42, 24, 25, -INFINITY;
  VectorXd beta(M), beta_size(M + 1), beta_value(M);
  beta << 0.3, 2;
  beta_size << 0.3, 2, 0.4;
  beta_value << 0.3, INFINITY;
  VectorXd alpha(N), alpha_size(N - 1), alpha_value(N);
  alpha << 0.3, -0.8, 1.8;
  alpha_size << 0.3, -0.8;
  alpha_value << 0.3, -0.8, NAN;
  VectorXd sigma(N), sigma_size(N - 1), sigma_value(N);
  sigma << 10, 4, 6;
  sigma_size << 10, 4;
  sigma_value << 10, 4, -6;

  matrix_cuda<double> x_d(x), x_size1_d(x_size1), x_size2_d(x_size2), x_value_d(x_value);
  matrix_cuda<double> y_d(y), y_size_d(y_size), y_value_d(y_value);
  matrix_cuda<double> alpha_d(alpha), alpha_size_d(alpha_size), alpha_value_d(alpha_value);
  matrix_cuda<double> sigma_d(sigma), sigma_size_d(sigma_size), sigma_value_d(sigma_value);

  using stan::math::normal_id_glm_lpdf;
  EXPECT_NO_THROW(normal_id_glm_lpdf(y_d, x_d, alpha_d, beta, sigma_d));
  EXPECT_THROW(normal_id_glm_lpdf(y_size_d, x_d, alpha_d, beta, sigma_d), std::invalid_argument);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_size1_d, alpha_d, beta, sigma_d), std::invalid_argument);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_size2_d, alpha_d, beta, sigma_d), std::invalid_argument);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_d, alpha_size_d, beta, sigma_d), std::invalid_argument);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_d, alpha_d, beta_size, sigma_d), std::invalid_argument);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_d, alpha_d, beta, sigma_size_d), std::invalid_argument);
  EXPECT_THROW(normal_id_glm_lpdf(y_value_d, x_d, alpha_d, beta, sigma_d), std::domain_error);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_value_d, alpha_d, beta, sigma_d), std::domain_error);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_d, alpha_value_d, beta, sigma_d), std::domain_error);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_d, alpha_d, beta_value, sigma_d), std::domain_error);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_d, alpha_d, beta, sigma_value_d), std::domain_error);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_d, alpha_d, beta, -1.0), std::domain_error);
  EXPECT_THROW(normal_id_glm_lpdf(y_d, x_d, alpha_d, beta, 0.0), std::domain_error);
}

TEST(CudaNormalIdGLM, known_answer) {
  VectorXd y(3);
  y << 14, 32, 21;
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  matrix_cuda<double> y_d(y);
  var alpha = 0.3, sigma = 10;
  Matrix<var, Dynamic, 1> beta(2);
  beta << 0.3, 2;
  Matrix<var, Dynamic, Dynamic> xv = x;
  auto xv_d = stan::math::to_matrix_cuda(xv);
  var lp = stan::math::normal_id_glm_lpdf(y_d, xv_d, alpha, beta, sigma);
  lp.grad();
  EXPECT_NEAR(lp.val(), -45.956670878596157, 1e-11);
  EXPECT_NEAR(alpha.adj(), -1.1920000000000002, 1e-11);
  EXPECT_NEAR(sigma.adj(), 6.9584200000000012, 1e-11);
  EXPECT_NEAR(beta[0].adj(), 0.31800000000000139, 1e-10);
  EXPECT_NEAR(beta[1].adj(), -46.265999999999998, 1e-10);
  const double dx[6] = {-0.2241, -0.0111, -0.1224, -1.494, -0.074, -0.816};
  for (int j = 0; j < 2; ++j)
    for (int i = 0; i < 3; ++i) EXPECT_NEAR(xv(i, j).adj(), dx[j * 3 + i], 1e-12);
  stan::math::recover_memory();
}

TEST(CudaNormalIdGLM, small_simple) {
  VectorXd y(3);
  y << 14, 32, 21;
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3, sigma = 11;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, sigma);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, sigma);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST, HOST), y, x, alpha, beta, sigma);
}

TEST(CudaNormalIdGLM, broadcast_y) {
  double y = 13;
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3, sigma = 11;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST, HOST), y, x, alpha, beta, sigma);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(HOST, DEV, HOST, HOST, HOST), y, x, alpha, beta, sigma);
}

TEST(CudaNormalIdGLM, zero_instances) {
  VectorXd y(0);
  MatrixXd x(0, 2);
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3, sigma = 11;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, sigma);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, sigma);
}

TEST(CudaNormalIdGLM, zero_attributes) {
  VectorXd y(3);
  y << 14, 32, 21;
  MatrixXd x(3, 0);
  VectorXd beta(0);
  double alpha = 0.3, sigma = 11;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, sigma);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, sigma);
}

TEST(CudaNormalIdGLM, small_vector_alpha_sigma) {
  VectorXd y(3);
  y << 14, 32, 21;
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  VectorXd beta(2);
  beta << 0.3, 2;
  VectorXd alpha(3), sigma(3);
  alpha << 0.3, -0.8, 1.8;
  sigma << 11, 12, 13;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST, DEV), y, x, alpha, beta, sigma);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, HOST, DEV), y, x, alpha, beta, sigma);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST, HOST), y, x, alpha, beta, sigma);
}

TEST(CudaNormalIdGLM, big) {
  int N = 153, M = 71;
  srand(3);
  VectorXd y = VectorXd::Random(N);
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M);
  VectorXd alpha = VectorXd::Random(N);
  VectorXd sigma = VectorXd::Random(N).array() + 1.1;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST, DEV), y, x, alpha, beta, sigma);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, HOST, DEV), y, x, alpha, beta, sigma);
}

TEST(CudaNormalIdGLM, config1_shape) {
  // BASELINE.json configs[0]: N = 10,000, K = 100, alpha / beta / sigma var
  int N = 10000, M = 100;
  srand(4);
  VectorXd y = VectorXd::Random(N) * 3;
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M) / 10.0;
  double alpha = 0.1, sigma = 1.3;
  matrix_cuda<double> x_d(x), y_d(y);
  var a1 = alpha, s1 = sigma, a2 = alpha, s2 = sigma;
  Matrix<var, Dynamic, 1> b1 = beta, b2 = beta;
  var lp_dev = stan::math::normal_id_glm_lpdf(y_d, x_d, a1, b1, s1);
  var lp_cpu = stan::math::normal_id_glm_lpdf(y, x, a2, b2, s2);
  (lp_dev + lp_cpu).grad();
  expect_close("logp", lp_dev.val(), lp_cpu.val(), kRelLogp, 0);
  expect_close("d_alpha", a1.adj(), a2.adj(), kRelGrad, b2.adj().cwiseAbs().maxCoeff());
  expect_close("d_sigma", s1.adj(), s2.adj(), kRelGrad, 0);
  compare_adj("d_beta", b1, b2);
  stan::math::recover_memory();
}
