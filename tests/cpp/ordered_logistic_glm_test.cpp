// CUDA-vs-CPU parity for ordered_logistic_glm_lpmf; cases follow the reference's
// device test test/unit/math/opencl/rev/ordered_logistic_glm_lpmf_test.cpp
// (error_checking, small_simple, broadcast_y, zero_instances, zero_attributes,
// single_class, big with C = 43) plus the known answer of SURVEY.md 8(c).
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
auto f = [](const auto& y, const auto& x, const auto& beta, const auto& cuts) {
  return stan::math::ordered_logistic_glm_lpmf(y, x, beta, cuts);
};
auto f_propto = [](const auto& y, const auto& x, const auto& beta, const auto& cuts) {
  return stan::math::ordered_logistic_glm_lpmf<true>(y, x, beta, cuts);
};
}  // namespace

TEST(CudaOrderedLogisticGLM, error_checking) {
  int N = 3, M = 2, C = 5;
  vector<int> y{1, 3, 2}, y_size{1, 3, 1, 2}, y_value1{0, 1, 2}, y_value2{1, 2, 6};
  MatrixXd x(N, M), x_size1(N - 1, M), x_size2(N, M - 1), x_value(N, M);
  x << -12, 46, -42, 24, 25, 27;
  x_size1 << -12, 46, -42, 24;
  x_size2 << -12, 46, -42;
  x_value << -12, 46, -42, 24, 25, -INFINITY;
  VectorXd beta(M), beta_size(M + 1), beta_value(M);
  beta << 0.3, 2;
  beta_size << 0.3, 2, 0.4;
  beta_value << 0.3, INFINITY;
  VectorXd cuts(C - 1), cuts_value1(C - 1), cuts_value2(C - 1);
  cuts << -0.4, 0.1, 0.3, 4.5;
  cuts_value1 << 0.3, -0.8, 1.8, 4.5;  // not ordered
  cuts_value2 << -0.4, 0.1, 0.3, INFINITY;

  matrix_cuda<double> x_d(x), x_size1_d(x_size1), x_size2_d(x_size2), x_value_d(x_value);
  matrix_cuda<int> y_d(y), y_size_d(y_size), y_value1_d(y_value1), y_value2_d(y_value2);

  using stan::math::ordered_logistic_glm_lpmf;
  EXPECT_NO_THROW(ordered_logistic_glm_lpmf(y_d, x_d, beta, cuts));
  EXPECT_THROW(ordered_logistic_glm_lpmf(y_size_d, x_d, beta, cuts), std::invalid_argument);
  EXPECT_THROW(ordered_logistic_glm_lpmf(y_d, x_size1_d, beta, cuts), std::invalid_argument);
  EXPECT_THROW(ordered_logistic_glm_lpmf(y_d, x_size2_d, beta, cuts), std::invalid_argument);
  EXPECT_THROW(ordered_logistic_glm_lpmf(y_d, x_d, beta_size, cuts), std::invalid_argument);
  EXPECT_THROW(ordered_logistic_glm_lpmf(y_value1_d, x_d, beta, cuts), std::domain_error);
  EXPECT_THROW(ordered_logistic_glm_lpmf(y_value2_d, x_d, beta, cuts), std::domain_error);
  EXPECT_THROW(ordered_logistic_glm_lpmf(y_d, x_value_d, beta, cuts), std::domain_error);
  EXPECT_THROW(ordered_logistic_glm_lpmf(y_d, x_d, beta_value, cuts), std::domain_error);
  EXPECT_THROW(ordered_logistic_glm_lpmf(y_d, x_d, beta, cuts_value1), std::domain_error);
  EXPECT_THROW(ordered_logistic_glm_lpmf(y_d, x_d, beta, cuts_value2), std::domain_error);
  EXPECT_THROW(ordered_logistic_glm_lpmf(7, x_d, beta, cuts), std::domain_error);
}

TEST(CudaOrderedLogisticGLM, known_answer) {
  vector<int> y{1, 1, 2, 4, 4};
  MatrixXd x(5, 2);
  x << 1, 2, 3, 4, 5, 6, 7, 8, 9, 0;
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> y_d(y);
  Matrix<var, Dynamic, 1> beta(2), cuts(3);
  beta << 1.1, 0.4;
  cuts << 0.9, 1.1, 7;
  var lp = stan::math::ordered_logistic_glm_lpmf(y_d, x_d, beta, cuts);
  lp.grad();
  EXPECT_NEAR(lp.val(), -13.914810581699157, 1e-11);
  EXPECT_NEAR(beta[0].adj(), -8.0587178047631252, 1e-10);
  EXPECT_NEAR(beta[1].adj(), -11.219308348175458, 1e-10);
  EXPECT_NEAR(cuts[0].adj(), -2.804494248653481, 1e-10);
  EXPECT_NEAR(cuts[1].adj(), 5.5155430300941335, 1e-10);
  EXPECT_NEAR(cuts[2].adj(), -0.071993868812495185, 1e-10);
  stan::math::recover_memory();
}

TEST(CudaOrderedLogisticGLM, small_simple) {
  vector<int> y{2, 1, 5};
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  VectorXd beta(2), cuts(4);
  beta << 0.3, 2;
  cuts << -0.4, 0.1, 0.3, 4.5;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST), y, x, beta, cuts);
}

TEST(CudaOrderedLogisticGLM, broadcast_y) {
  int y = 2;
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  VectorXd beta(2), cuts(4);
  beta << 0.3, 2;
  cuts << -0.4, 0.1, 0.3, 4.5;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, DEV, HOST, HOST), y, x, beta, cuts);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(HOST, DEV, HOST, HOST), y, x, beta, cuts);
}

TEST(CudaOrderedLogisticGLM, zero_instances) {
  vector<int> y{};
  MatrixXd x(0, 2);
  VectorXd beta(2), cuts(4);
  beta << 0.3, 2;
  cuts << -0.4, 0.1, 0.3, 4.5;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
}

TEST(CudaOrderedLogisticGLM, zero_attributes) {
  vector<int> y{2, 1, 5};
  MatrixXd x(3, 0);
  VectorXd beta(0), cuts(4);
  cuts << -0.4, 0.1, 0.3, 4.5;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
}

TEST(CudaOrderedLogisticGLM, single_class) {
  vector<int> y{1, 1, 1};
  MatrixXd x(3, 2);
  x << -1.2, 4.6, -4.2, 2.4, 2.5, 2.7;
  VectorXd beta(2), cuts(0);
  beta << 0.3, 2;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
}

TEST(CudaOrderedLogisticGLM, big) {
  int N = 153, M = 71, C = 43;
  srand(7);
  vector<int> y(N);
  for (int i = 0; i < N; i++) y[i] = std::abs(Eigen::Array<int, 1, 1>::Random()[0]) % C + 1;
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M);
  VectorXd cuts = VectorXd::Random(C - 1);
  std::sort(cuts.data(), cuts.data() + cuts.size());
  for (int c = 1; c < C - 1; ++c) cuts[c] = std::max(cuts[c], cuts[c - 1] + 1e-3);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
  compare_cpu_cuda_prim_rev(f_propto, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
}

TEST(CudaOrderedLogisticGLM, config5b_shape) {
  // BASELINE.json configs[4] (ordered part) at reduced N: K = 64, 8 cuts
  int N = 30011, M = 64;
  srand(8);
  vector<int> y(N);
  for (int i = 0; i < N; i++) y[i] = (i * 11) % 9 + 1;
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M) / std::sqrt(M);
  VectorXd cuts = VectorXd::LinSpaced(8, -2, 2);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(DEV, DEV, HOST, HOST), y, x, beta, cuts);
}
