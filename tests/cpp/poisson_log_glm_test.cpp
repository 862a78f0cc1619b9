// CUDA-vs-CPU parity for poisson_log_glm_lpmf; cases follow the reference's
// device test test/unit/math/opencl/rev/poisson_log_glm_lpmf_test.cpp
// (error_checking, small_simple, broadcast_y, zero_instances, zero_attributes,
// small_vector_alpha, big) plus the known answers of SURVEY.md 8(c) and the
// interface types of test/unit/math/rev/prob/poisson_log_glm_lpmf_test.cpp
// L466-518.
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
  return stan::math::poisson_log_glm_lpmf(y, x, alpha, beta);
};
auto f_propto = [](const auto& y, const auto& x, const auto& alpha, const auto& beta) {
  return stan::math::poisson_log_glm_lpmf<true>(y, x, alpha, beta);
};
}  // namespace

TEST(CudaPoissonLogGLM, error_checking) {
  int N = 3, M = 2;
  vector<int> y{1, 0, 1}, y_size{1, 0, 1, 0}, y_value{0, 1, -23};
  MatrixXd x(N, M), x_size1(N - 1, M), x_size2(N, M - 1), x_value(N, M);
  x << -12, 46, -42, 24, 25, 27;
  x_size1 << -12, 46, -42, 24;
  x_size2 << -12, 46, -42;
  x_value << -12, 46, -42, 24, 25, NAN;  // (-INFINITY: prim returns -inf, the device overloads throw: ref_poisson_log_glm_lpmf_test)
  VectorXd beta(M), beta_size(M + 1), beta_value(M);
  beta << 0.3, 2;
  beta_size << 0.3, 2, 0.4;
  beta_value << 0.3, INFINITY;
  VectorXd alpha(N), alpha_size(N - 1), alpha_value(N);
  alpha << 0.3, -0.8, 1.8;
  alpha_size << 0.3, -0.8;
  alpha_value << 0.3, -0.8, NAN;

  matrix_cuda<double> x_d(x), x_size1_d(x_size1), x_size2_d(x_size2), x_value_d(x_value);
  matrix_cuda<int> y_d(y), y_size_d(y_size), y_value_d(y_value);
  matrix_cuda<double> alpha_d(alpha), alpha_size_d(alpha_size), alpha_value_d(alpha_value);

  using stan::math::poisson_log_glm_lpmf;
  EXPECT_NO_THROW(poisson_log_glm_lpmf(y_d, x_d, alpha_d, beta));
  EXPECT_THROW(poisson_log_glm_lpmf(y_size_d, x_d, alpha_d, beta), std::invalid_argument);
  EXPECT_THROW(poisson_log_glm_lpmf(y_d, x_size1_d, alpha_d, beta), std::invalid_argument);
  EXPECT_THROW(poisson_log_glm_lpmf(y_d, x_size2_d, alpha_d, beta), std::invalid_argument);
  EXPECT_THROW(poisson_log_glm_lpmf(y_d, x_d, alpha_size_d, beta), std::invalid_argument);
  EXPECT_THROW(poisson_log_glm_lpmf(y_d, x_d, alpha_d, beta_size), std::invalid_argument);
  EXPECT_THROW(poisson_log_glm_lpmf(y_value_d, x_d, alpha_d, beta), std::domain_error);
  EXPECT_THROW(poisson_log_glm_lpmf(y_d, x_value_d, alpha_d, beta), std::domain_error);
  EXPECT_THROW(poisson_log_glm_lpmf(y_d, x_d, alpha_value_d, beta), std::domain_error);
  EXPECT_THROW(poisson_log_glm_lpmf(y_d, x_d, alpha_d, beta_value), std::domain_error);
  // host per-row containers take the same route
  EXPECT_THROW(poisson_log_glm_lpmf(y_size, x_d, alpha, beta), std::invalid_argument);
  EXPECT_THROW(poisson_log_glm_lpmf(y_value, x_d, alpha, beta), std::domain_error);
  EXPECT_THROW(poisson_log_glm_lpmf(-5, x_d, alpha, beta), std::domain_error);
}

TEST(CudaPoissonLogGLM, known_answer) {
  // reference prim output on the reference test's fixed inputs
  vector<int> y{14, 2, 5};
  MatrixXd x(3, 2);
  x << -12, 46, -42, 24, 25, 27;
  x /= 100.0;
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> y_d(y);
  var alpha = 0.3;
  Matrix<var, Dynamic, 1> beta(2);
  beta << 0.3, 2;
  var lp = stan::math::poisson_log_glm_lpmf(y_d, x_d, alpha, beta);
  lp.grad();
  EXPECT_NEAR(lp.val(), -15.900271464537942, 1e-12);
  EXPECT_NEAR(alpha.adj(), 13.312588641542726, 1e-11);
  EXPECT_NEAR(beta[0].adj(), -0.69435197905709523, 1e-11);
  EXPECT_NEAR(beta[1].adj(), 5.631286107135665, 1e-11);
  stan::math::recover_memory();
}

TEST(CudaPoissonLogGLM, small_simple) {
  vector<int> y{2, 0, 5};
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  // y kept on the host (uploaded per call)
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaPoissonLogGLM, broadcast_y) {
  int y = 1;
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(HOST, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaPoissonLogGLM, zero_instances) {
  vector<int> y{};
  MatrixXd x(0, 2);
  VectorXd beta(2);
  beta << 0.3, 2;
  double alpha = 0.3;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaPoissonLogGLM, zero_attributes) {
  vector<int> y{2, 0, 5};
  MatrixXd x(3, 0);
  VectorXd beta(0);
  double alpha = 0.3;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaPoissonLogGLM, small_vector_alpha) {
  vector<int> y{2, 0, 5};
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  VectorXd beta(2);
  beta << 0.3, 2;
  VectorXd alpha(3);
  alpha << 0.3, -0.8, 1.8;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
}

TEST(CudaPoissonLogGLM, big) {
  int N = 153, M = 71;  // deliberately not multiples of any tile size
  srand(1);
  vector<int> y(N);
  for (int i = 0; i < N; i++) y[i] = std::abs(Eigen::Array<int, 1, 1>::Random()[0]) % 200;
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M);
  VectorXd alpha = VectorXd::Random(N);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, DEV, HOST), y, x, alpha, beta);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, DEV, HOST), y, x, alpha, beta);
}

TEST(CudaPoissonLogGLM, wide_rows_more_than_one_tile) {
  // N spans several row tiles, K > 256 takes the general two-pass path
  srand(2);
  for (int M : {64, 256, 300}) {
    int N = 4099;
    vector<int> y(N);
    for (int i = 0; i < N; i++) y[i] = (i * 7 + M) % 6;
    MatrixXd x = MatrixXd::Random(N, M);
    VectorXd beta = VectorXd::Random(M) / std::sqrt(M);
    double alpha = 0.1;
    compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, alpha, beta);
  }
}

TEST(CudaPoissonLogGLM, interface_types) {
  // every parameter container the reference's interface test accepts
  vector<int> y{1, 0, 1};
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> y_d(y);
  VectorXd beta(2);
  beta << 0.3, 2;
  const double expect = stan::math::poisson_log_glm_lpmf(y, x, 0.3, beta);
  using stan::math::poisson_log_glm_lpmf;
  EXPECT_NEAR(poisson_log_glm_lpmf(y_d, x_d, 0.3, beta), expect, 1e-12);
  Eigen::RowVectorXd beta_row = beta.transpose();
  EXPECT_NEAR(poisson_log_glm_lpmf(y_d, x_d, 0.3, beta_row), expect, 1e-12);
  vector<double> beta_std{0.3, 2};
  EXPECT_NEAR(poisson_log_glm_lpmf(y_d, x_d, 0.3, beta_std), expect, 1e-12);
  {
    vector<var> b{0.3, 2};
    var lp = poisson_log_glm_lpmf(y_d, x_d, var(0.3), b);
    vector<var> b2{0.3, 2};
    var lp2 = poisson_log_glm_lpmf(y, x, var(0.3), b2);
    (lp + lp2).grad();
    EXPECT_NEAR(lp.val(), expect, 1e-12);
    for (int k = 0; k < 2; ++k) EXPECT_NEAR(b[k].adj(), b2[k].adj(), 1e-10);
    stan::math::recover_memory();
  }
  {
    stan::math::var_value<VectorXd> b(beta), b2(beta);
    var lp = poisson_log_glm_lpmf(y_d, x_d, 0.3, b);
    var lp2 = poisson_log_glm_lpmf(y, x, 0.3, b2);
    (lp + lp2).grad();
    EXPECT_NEAR(lp.val(), expect, 1e-12);
    for (int k = 0; k < 2; ++k) EXPECT_NEAR(b.adj()[k], b2.adj()[k], 1e-10);
    stan::math::recover_memory();
  }
  {
    Matrix<var, 1, Dynamic> b = beta_row, b2 = beta_row;
    var lp = poisson_log_glm_lpmf(y_d, x_d, 0.3, b);
    var lp2 = poisson_log_glm_lpmf(y, x, 0.3, b2);
    (lp + lp2).grad();
    for (int k = 0; k < 2; ++k) EXPECT_NEAR(b[k].adj(), b2[k].adj(), 1e-10);
    stan::math::recover_memory();
  }
  {  // K = 1 with a scalar beta
    MatrixXd x1 = x.col(0);
    matrix_cuda<double> x1_d(x1);
    var b = 0.7, b2 = 0.7;
    var lp = poisson_log_glm_lpmf(y_d, x1_d, 0.3, b);
    var lp2 = poisson_log_glm_lpmf(y, x1, 0.3, b2);
    (lp + lp2).grad();
    EXPECT_NEAR(lp.val(), lp2.val(), 1e-12);
    EXPECT_NEAR(b.adj(), b2.adj(), 1e-10);
    stan::math::recover_memory();
  }
}
