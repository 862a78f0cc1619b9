// CUDA-vs-CPU parity for neg_binomial_2_log_glm_lpmf; cases follow the
// reference's device test
// test/unit/math/opencl/rev/neg_binomial_2_log_glm_lpmf_test.cpp (error_checking,
// small_simple, broadcast_y, zero_instances, zero_attributes,
// small_vector_alpha_phi, big) plus the known answer of SURVEY.md 8(c).
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
            const auto& phi) {
  return stan::math::neg_binomial_2_log_glm_lpmf(y, x, alpha, beta, phi);
};
auto f_propto = [](const auto& y, const auto& x, const auto& alpha, const auto& beta,
                   const auto& phi) {
  return stan::math::neg_binomial_2_log_glm_lpmf<true>(y, x, alpha, beta, phi);
};
}  // namespace

TEST(CudaNegBinomial2LogGLM, error_checking) {
  int N = 3, M = 2;
  vector<int> y{1, 0, 1}, y_size{1, 0, 1, 0}, y_value{1, 4, -23};
  MatrixXd x(N, M), x_size1(N - 1, M), x_size2(N, M - 1), x_value(N, M);
  x << -12, 46, -42, 24, 25, 27;
  x /= 10;
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
  VectorXd phi(N), phi_size(N - 1), phi_value(N);
  phi << 10, 4, 6;
  phi_size << 10, 4;
  phi_value << 10, 4, -6;

  matrix_cuda<double> x_d(x), x_size1_d(x_size1), x_size2_d(x_size2), x_value_d(x_value);
  matrix_cuda<int> y_d(y), y_size_d(y_size), y_value_d(y_value);
  matrix_cuda<double> alpha_d(alpha), alpha_size_d(alpha_size), alpha_value_d(alpha_value);
  matrix_cuda<double> phi_d(phi), phi_size_d(phi_size), phi_value_d(phi_value);

  using stan::math::neg_binomial_2_log_glm_lpmf;
  EXPECT_NO_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_d, alpha_d, beta, phi_d));
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_size_d, x_d, alpha_d, beta, phi_d), std::invalid_argument);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_size1_d, alpha_d, beta, phi_d), std::invalid_argument);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_size2_d, alpha_d, beta, phi_d), std::invalid_argument);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_d, alpha_size_d, beta, phi_d), std::invalid_argument);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_d, alpha_d, beta_size, phi_d), std::invalid_argument);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_d, alpha_d, beta, phi_size_d), std::invalid_argument);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_value_d, x_d, alpha_d, beta, phi_d), std::domain_error);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_value_d, alpha_d, beta, phi_d), std::domain_error);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_d, alpha_value_d, beta, phi_d), std::domain_error);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_d, alpha_d, beta_value, phi_d), std::domain_error);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_d, alpha_d, beta, phi_value_d), std::domain_error);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(y_d, x_d, alpha_d, beta, -2.0), std::domain_error);
  EXPECT_THROW(neg_binomial_2_log_glm_lpmf(-1, x_d, alpha_d, beta, 2.0), std::domain_error);
}

TEST(CudaNegBinomial2LogGLM, known_answer) {
  vector<int> y{14, 2, 5};
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  x /= 100.0;
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> y_d(y);
  var alpha = 0.3, phi = 2;
  Matrix<var, Dynamic, 1> beta(2);
  beta << 0.3, 2;
  var lp = stan::math::neg_binomial_2_log_glm_lpmf(y_d, x_d, alpha, beta, phi);
  lp.grad();
  EXPECT_NEAR(lp.val(), -10.359512037713642, 1e-12);
  EXPECT_NEAR(alpha.adj(), 5.2275673197452948, 1e-11);
  EXPECT_NEAR(phi.adj(), -0.46459289780560997, 1e-11);
  EXPECT_NEAR(beta[0].adj(), -0.22711413174797268, 1e-11);
  EXPECT_NEAR(beta[1].adj(), 2.1845346756425257, 1e-11);
  stan::math::recover_memory();
}

TEST(CudaNegBinomial2LogGLM, small_simple) {
  vector<int> y{2, 0, 5};
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3, phi = 13.2;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, phi);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, phi);
}

TEST(CudaNegBinomial2LogGLM, broadcast_y) {
  int y = 2;
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3, phi = 13.2;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST, HOST), y, x, alpha, beta, phi);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(HOST, DEV, HOST, HOST, HOST), y, x, alpha, beta, phi);
}

TEST(CudaNegBinomial2LogGLM, zero_instances) {
  vector<int> y{};
  MatrixXd x(0, 2);
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3, phi = 13.2;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, phi);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, phi);
}

TEST(CudaNegBinomial2LogGLM, zero_attributes) {
  vector<int> y{2, 0, 5};
  MatrixXd x(3, 0);
  VectorXd beta(0);
  double alpha = 0.3, phi = 13.2;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, phi);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST, HOST), y, x, alpha, beta, phi);
}

TEST(CudaNegBinomial2LogGLM, small_vector_alpha_phi) {
  vector<int> y{2, 0, 5};
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  VectorXd beta(2);
  beta << 0.3, 2;
  VectorXd alpha(3), phi(3);
  alpha << 0.3, -0.8, 1.8;
  phi << 13.2, 0.7, 4;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST, DEV), y, x, alpha, beta, phi);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, HOST, DEV), y, x, alpha, beta, phi);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST, HOST), y, x, alpha, beta, phi);
}

TEST(CudaNegBinomial2LogGLM, big) {
  int N = 153, M = 71;
  srand(5);
  vector<int> y(N);
  for (int i = 0; i < N; i++) y[i] = std::abs(Eigen::Array<int, 1, 1>::Random()[0]) % 200;
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M);
  VectorXd alpha = VectorXd::Random(N);
  VectorXd phi = VectorXd::Random(N).array() + 1.1;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST, DEV), y, x, alpha, beta, phi);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, HOST, DEV), y, x, alpha, beta, phi);
}

TEST(CudaNegBinomial2LogGLM, config4_shape_x_var_on_device) {
  // BASELINE.json configs[3] at reduced N: phi var AND x var, the N x K adjoint
  // of x stays on the device (var_value<matrix_cuda<double>>)
  int N = 20011, M = 128;
  srand(6);
  vector<int> y(N);
  for (int i = 0; i < N; i++) y[i] = (i * 13) % 5;
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M) / std::sqrt(M);
  matrix_cuda<int> y_d(y);
  stan::math::var_value<matrix_cuda<double>> x_d{matrix_cuda<double>(x)};
  Matrix<var, Dynamic, Dynamic> x_c = x;
  var p1 = 2.5, p2 = 2.5;
  Matrix<var, Dynamic, 1> b1 = beta, b2 = beta;
  var lp_dev = stan::math::neg_binomial_2_log_glm_lpmf(y_d, x_d, 0.1, b1, p1);
  var lp_cpu = stan::math::neg_binomial_2_log_glm_lpmf(y, x_c, 0.1, b2, p2);
  (lp_dev + lp_cpu).grad();
  expect_close("logp", lp_dev.val(), lp_cpu.val(), kRelLogp, 0);
  expect_close("d_phi", p1.adj(), p2.adj(), kRelGrad, 0);
  compare_adj("d_beta", b1, b2);
  MatrixXd dx_dev = stan::math::from_matrix_cuda(x_d.adj().to_matrix_cuda());
  MatrixXd dx_cpu = x_c.adj();
  const double scale = dx_cpu.cwiseAbs().maxCoeff();
  double worst = 0;
  for (int j = 0; j < M; ++j)
    for (int i = 0; i < N; ++i)
      worst = std::max(worst, std::fabs(dx_dev(i, j) - dx_cpu(i, j))
                                  / (kRelGrad * std::fabs(dx_cpu(i, j)) + kAbsFloor * scale));
  EXPECT_LE(worst, 1.0) << "d_x (device adjoint) vs prim";
  stan::math::recover_memory();
}
