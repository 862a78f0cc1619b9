// matrix_cuda / var_value<matrix_cuda<double>> round trips: the cases of the
// reference's test/unit/math/opencl/copy_test.cpp and rev/copy_test.cpp that
// the GLM path relies on (Eigen / std::vector / scalar copies, copy and move
// semantics, var round trip with adjoint flow, arena lifetime).
#include "cuda_test_util.hpp"

using Eigen::MatrixXd;
using Eigen::VectorXd;
using stan::math::from_matrix_cuda;
using stan::math::matrix_cuda;
using stan::math::to_matrix_cuda;
using stan::math::var;
using stan::math::var_value;

TEST(CudaMatrix, eigen_round_trip) {
  MatrixXd a = MatrixXd::Random(7, 5);
  matrix_cuda<double> d(a);
  EXPECT_EQ(d.rows(), 7);
  EXPECT_EQ(d.cols(), 5);
  EXPECT_EQ(d.size(), 35);
  MatrixXd b = from_matrix_cuda(d);
  EXPECT_TRUE((a.array() == b.array()).all());
  VectorXd v = VectorXd::Random(11);
  EXPECT_TRUE((from_matrix_cuda<VectorXd>(to_matrix_cuda(v)).array() == v.array()).all());
  Eigen::RowVectorXd r = Eigen::RowVectorXd::Random(9);
  EXPECT_TRUE((from_matrix_cuda<Eigen::RowVectorXd>(to_matrix_cuda(r)).array() == r.array()).all());
  // expressions and blocks upload their evaluated value
  MatrixXd c = from_matrix_cuda(to_matrix_cuda(a.block(1, 1, 3, 2) * 2.0));
  EXPECT_TRUE((c.array() == (a.block(1, 1, 3, 2) * 2.0).array()).all());
}

TEST(CudaMatrix, std_vector_and_scalar) {
  std::vector<int> y{3, 1, 4, 1, 5, 9, 2, 6};
  matrix_cuda<int> d(y);
  EXPECT_EQ(d.rows(), 8);
  EXPECT_EQ(d.cols(), 1);
  EXPECT_EQ(from_matrix_cuda<std::vector<int>>(d), y);
  EXPECT_EQ(from_matrix_cuda<double>(to_matrix_cuda(2.5)), 2.5);
  EXPECT_EQ(from_matrix_cuda<int>(to_matrix_cuda(7)), 7);
  MatrixXd two(2, 1);
  two << 1, 2;
  EXPECT_THROW(from_matrix_cuda<double>(to_matrix_cuda(two)), std::invalid_argument);
}

TEST(CudaMatrix, copy_move_view) {
  MatrixXd a = MatrixXd::Random(33, 3);
  matrix_cuda<double> d(a);
  matrix_cuda<double> copy(d);  // deep copy
  EXPECT_NE(copy.handle(), d.handle());
  EXPECT_TRUE((from_matrix_cuda(copy).array() == a.array()).all());
  smc_matrix* h = copy.handle();
  matrix_cuda<double> moved(std::move(copy));
  EXPECT_EQ(moved.handle(), h);
  EXPECT_EQ(copy.handle(), nullptr);
  matrix_cuda<double> v = matrix_cuda<double>::view(d);
  EXPECT_TRUE((from_matrix_cuda(v).array() == a.array()).all());
  matrix_cuda<double> z(4, 4);
  z.zero();
  EXPECT_EQ(from_matrix_cuda(z).cwiseAbs().sum(), 0.0);
  matrix_cuda<double> empty(0, 3);
  EXPECT_EQ(empty.size(), 0);
  EXPECT_EQ(from_matrix_cuda(empty).size(), 0);
}

TEST(CudaMatrix, var_round_trip_adjoints) {
  MatrixXd a = MatrixXd::Random(6, 4);
  Eigen::Matrix<var, -1, -1> av = a;
  var_value<matrix_cuda<double>> d = to_matrix_cuda(av);
  EXPECT_EQ(d.rows(), 6);
  EXPECT_EQ(d.cols(), 4);
  var_value<MatrixXd> back = from_matrix_cuda(d);
  EXPECT_TRUE((back.val().array() == a.array()).all());
  MatrixXd w = MatrixXd::Random(6, 4);
  var total = stan::math::sum(stan::math::elt_multiply(back, w));
  total.grad();
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 6; ++i) EXPECT_DOUBLE_EQ(av(i, j).adj(), w(i, j));
  stan::math::set_zero_all_adjoints();
  EXPECT_EQ(from_matrix_cuda(d.adj().to_matrix_cuda()).cwiseAbs().sum(), 0.0);
  stan::math::recover_memory();  // frees the arena-owned device buffers
}

TEST(CudaMatrix, var_value_eigen_and_std_vector) {
  VectorXd a = VectorXd::Random(5);
  var_value<VectorXd> av(a);
  var_value<matrix_cuda<double>> d = to_matrix_cuda(av);
  std::vector<var> sv{1.0, 2.0, 3.0};
  var_value<matrix_cuda<double>> e = to_matrix_cuda(sv);
  EXPECT_EQ(e.rows(), 3);
  var_value<VectorXd> b1 = from_matrix_cuda<VectorXd>(d);
  var_value<VectorXd> b2 = from_matrix_cuda<VectorXd>(e);
  var total = stan::math::sum(b1) * 2.0 + stan::math::sum(b2) * 3.0;
  total.grad();
  for (int i = 0; i < 5; ++i) EXPECT_DOUBLE_EQ(av.adj()[i], 2.0);
  for (int i = 0; i < 3; ++i) EXPECT_DOUBLE_EQ(sv[i].adj(), 3.0);
  stan::math::recover_memory();
}

// SURVEY.md 8(b) "Threading" / 8(f)4: several chains evaluate the same GLM
// concurrently on ONE shared, read-only device x.  The C ABI keeps its stream and
// workspace per host thread, so the calls are re-entrant; every thread must get
// exactly what a lone call gets (the reductions are fixed-order).
#include <thread>
TEST(CudaMatrix, concurrent_chains_share_one_design_matrix) {
  const int N = 20011, K = 96, T = 4, REPS = 8;
  srand(11);
  MatrixXd x = MatrixXd::Random(N, K);
  std::vector<int> y(N);
  for (int i = 0; i < N; ++i) y[i] = (i * 13) % 2;
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> y_d(y);
  std::vector<VectorXd> betas(T);
  std::vector<double> want(T);
  for (int t = 0; t < T; ++t) {
    betas[t] = VectorXd::Random(K) / std::sqrt(K);
    want[t] = stan::math::bernoulli_logit_glm_lpmf(y_d, x_d, 0.1 * t, betas[t]);
    EXPECT_NEAR(want[t], stan::math::bernoulli_logit_glm_lpmf(y, x, 0.1 * t, betas[t]),
                1e-10 * std::fabs(want[t]));
  }
  {
    // first touch from all threads at once: the caches that live on the shared
    // matrices (TMA descriptor of x, range of y) are filled in under a lock
    matrix_cuda<double> x_cold(x);
    matrix_cuda<int> y_cold(y);
    std::vector<double> got(T, 0.0);
    std::vector<std::thread> first;
    for (int t = 0; t < T; ++t) {
      first.emplace_back([&, t]() {
        got[t] = stan::math::bernoulli_logit_glm_lpmf(y_cold, x_cold, 0.1 * t, betas[t]);
      });
    }
    for (auto& th : first) th.join();
    for (int t = 0; t < T; ++t) EXPECT_EQ(got[t], want[t]) << "cold thread " << t;
  }
  std::vector<int> mismatches(T, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < T; ++t) {
    pool.emplace_back([&, t]() {
      for (int r = 0; r < REPS; ++r) {
        const double got = stan::math::bernoulli_logit_glm_lpmf(y_d, x_d, 0.1 * t, betas[t]);
        if (got != want[t]) ++mismatches[t];  // bit-identical, not merely close
      }
    });
  }
  for (auto& th : pool) th.join();
  for (int t = 0; t < T; ++t) EXPECT_EQ(mismatches[t], 0) << "thread " << t;
}
