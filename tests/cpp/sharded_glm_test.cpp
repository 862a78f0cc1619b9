// Multi-GPU behind the boundary: the SAME overloads on row-sharded device matrices
// (matrix_cuda_sharded / to_matrix_cuda_sharded) against the reference's prim
// implementation for every prim / var combination, and against the single-GPU
// overload.  The shard set is one shard per GPU when the box has at least two (NCCL
// all-reduce of the packed partials), else three shards on GPU 0 (host-side sum) --
// the sharding logic, the per-shard constant terms and the sharded N-vector / N x K
// partials are exercised either way.  Pattern matched:
// stan/math/prim/functor/mpi_parallel_call.hpp L332-392, L408-449.
#include "cuda_test_util.hpp"

using Eigen::Dynamic;
using Eigen::Matrix;
using Eigen::MatrixXd;
using Eigen::VectorXd;
using stan::math::matrix_cuda;
using stan::math::matrix_cuda_sharded;
using stan::math::var;
using std::vector;
using namespace cuda_test;  // NOLINT

namespace {

class CudaShardedGLM : public ::testing::Test {
 protected:
  static void SetUpTestSuite() {
    int n_dev = 0;
    ASSERT_EQ(smc_device_count(&n_dev), SMC_OK) << smc_last_error();
    if (n_dev >= 2) {
      ASSERT_EQ(smc_shard_init(n_dev, nullptr), SMC_OK) << smc_last_error();
    } else {
      const int dev[3] = {0, 0, 0};
      ASSERT_EQ(smc_shard_init(3, dev), SMC_OK) << smc_last_error();
    }
    std::printf("shard set: %d shards, reduction: %s\n", stan::math::cuda_shard_count(),
                smc_shard_reduce_mode());
  }
  static void TearDownTestSuite() { stan::math::cuda_shard_shutdown(); }
};

vector<int> counts(int N, int mod) {
  vector<int> y(N);
  for (int i = 0; i < N; ++i) y[i] = (i * 7 + 3) % mod;
  return y;
}

}  // namespace

TEST_F(CudaShardedGLM, bernoulli) {
  auto f = [](const auto& y, const auto& x, const auto& alpha, const auto& beta) {
    return stan::math::bernoulli_logit_glm_lpmf(y, x, alpha, beta);
  };
  auto fp = [](const auto& y, const auto& x, const auto& alpha, const auto& beta) {
    return stan::math::bernoulli_logit_glm_lpmf<true>(y, x, alpha, beta);
  };
  srand(3);
  for (int N : {2, 153, 4099}) {
    const int M = N < 100 ? 2 : 71;
    vector<int> y = counts(N, 2);
    MatrixXd x = MatrixXd::Random(N, M);
    VectorXd beta = VectorXd::Random(M);
    VectorXd alpha_vec = VectorXd::Random(N);
    double alpha = 0.3;
    compare_cpu_cuda_prim_rev(f, std::make_tuple(SHARD, SHARD, HOST, HOST), y, x, alpha, beta);
    compare_cpu_cuda_prim_rev(fp, std::make_tuple(SHARD, SHARD, HOST, HOST), y, x, alpha, beta);
    // vector alpha on the device (sharded, also as an autodiff variable) and on the host
    compare_cpu_cuda_prim_rev(f, std::make_tuple(SHARD, SHARD, SHARD, HOST), y, x, alpha_vec,
                              beta);
    compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, SHARD, HOST, HOST), y, x, alpha_vec,
                              beta);
  }
}

TEST_F(CudaShardedGLM, poisson_and_broadcast_y) {
  auto f = [](const auto& y, const auto& x, const auto& alpha, const auto& beta) {
    return stan::math::poisson_log_glm_lpmf(y, x, alpha, beta);
  };
  srand(4);
  const int N = 1531, M = 64;
  vector<int> y = counts(N, 6);
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M) / 8.0;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(SHARD, SHARD, HOST, HOST), y, x, 0.1, beta);
  // lgamma(y + 1) of a broadcast scalar y enters once per call, not once per shard
  compare_cpu_cuda_prim_rev(f, std::make_tuple(HOST, SHARD, HOST, HOST), 3, x, 0.1, beta);
}

TEST_F(CudaShardedGLM, normal_id) {
  auto f = [](const auto& y, const auto& x, const auto& alpha, const auto& beta,
              const auto& sigma) {
    return stan::math::normal_id_glm_lpdf(y, x, alpha, beta, sigma);
  };
  srand(5);
  const int N = 777, M = 33;
  VectorXd y = VectorXd::Random(N);
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M);
  VectorXd sigma_vec = VectorXd::Random(N).array() + 1.5;
  compare_cpu_cuda_prim_rev(f, std::make_tuple(SHARD, SHARD, HOST, HOST, HOST), y, x, 0.3, beta,
                            1.3);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(SHARD, SHARD, HOST, HOST, SHARD), y, x, 0.3,
                            beta, sigma_vec);
}

TEST_F(CudaShardedGLM, neg_binomial_ordered_binomial) {
  srand(6);
  const int N = 1000, M = 17;
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M);
  {
    auto f = [](const auto& y, const auto& x, const auto& alpha, const auto& beta,
                const auto& phi) {
      return stan::math::neg_binomial_2_log_glm_lpmf(y, x, alpha, beta, phi);
    };
    compare_cpu_cuda_prim_rev(f, std::make_tuple(SHARD, SHARD, HOST, HOST, HOST), counts(N, 9), x,
                              0.2, beta, 2.5);
  }
  {
    auto f = [](const auto& y, const auto& x, const auto& beta, const auto& cuts) {
      return stan::math::ordered_logistic_glm_lpmf(y, x, beta, cuts);
    };
    vector<int> y = counts(N, 5);
    for (int& v : y) v += 1;
    VectorXd cuts(4);
    cuts << -1.1, -0.2, 0.4, 1.7;
    compare_cpu_cuda_prim_rev(f, std::make_tuple(SHARD, SHARD, HOST, HOST), y, x, beta, cuts);
  }
  {
    auto f = [](const auto& n, const auto& N_trials, const auto& x, const auto& alpha,
                const auto& beta) {
      return stan::math::binomial_logit_glm_lpmf(n, N_trials, x, alpha, beta);
    };
    vector<int> trials = counts(N, 40), n(N);
    for (int i = 0; i < N; ++i) n[i] = trials[i] / 3;
    compare_cpu_cuda_prim_rev(f, std::make_tuple(SHARD, SHARD, SHARD, HOST, HOST), n, trials, x,
                              0.1, beta);
  }
}

TEST_F(CudaShardedGLM, categorical) {
  auto f = [](const auto& y, const auto& x, const auto& alpha, const auto& beta) {
    return stan::math::categorical_logit_glm_lpmf(y, x, alpha, beta);
  };
  srand(7);
  const int N = 600, M = 20, C = 5;
  vector<int> y = counts(N, C);
  for (int& v : y) v += 1;
  MatrixXd x = MatrixXd::Random(N, M);
  MatrixXd beta = MatrixXd::Random(M, C);
  VectorXd alpha = VectorXd::Random(C);
  compare_cpu_cuda_prim_rev(f, std::make_tuple(SHARD, SHARD, HOST, HOST), y, x, alpha, beta);
}

TEST_F(CudaShardedGLM, autodiff_design_matrix_stays_sharded) {
  // x as var_value<matrix_cuda> over the shard set: value and N x K adjoint sharded, the
  // reverse sweep x.adj() += lp.adj() * d beta^T runs shard by shard, no exchange
  srand(8);
  const int N = 2000, M = 40;
  vector<int> y = counts(N, 7);
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta0 = VectorXd::Random(M) / 6.0;
  matrix_cuda_sharded<double> x_sh(x);
  matrix_cuda_sharded<int> y_sh(y);
  EXPECT_EQ(x_sh.shard_count(), stan::math::cuda_shard_count());

  stan::math::var_value<matrix_cuda<double>> xv(x_sh);
  Matrix<var, Dynamic, 1> beta = beta0;
  var alpha = 0.1, phi = 2.5;
  var lp = stan::math::neg_binomial_2_log_glm_lpmf(y_sh, xv, alpha, beta, phi);
  Matrix<var, Dynamic, Dynamic> x_host = x;
  Matrix<var, Dynamic, 1> beta_h = beta0;
  var alpha_h = 0.1, phi_h = 2.5;
  var lp_h = stan::math::neg_binomial_2_log_glm_lpmf(y, x_host, alpha_h, beta_h, phi_h);
  expect_close("log density", lp.val(), lp_h.val(), kRelLogp, 0.0);
  (2.0 * lp + lp_h).grad();
  EXPECT_EQ(xv.adj().handle() ? smc_matrix_shard_count(xv.adj().handle()) : -1,
            stan::math::cuda_shard_count());
  const MatrixXd gx = stan::math::from_matrix_cuda(xv.adj().to_matrix_cuda());
  const MatrixXd gx_h = x_host.adj();
  const double scale = gx_h.cwiseAbs().maxCoeff();
  for (int j = 0; j < M; ++j)
    for (int i = 0; i < N; ++i)
      expect_close("x adjoint", gx(i, j), 2.0 * gx_h(i, j), kRelGrad, scale);
  for (int k = 0; k < M; ++k)
    expect_close("beta adjoint", beta[k].adj(), 2.0 * beta_h[k].adj(), kRelGrad, 0.0);
  expect_close("phi adjoint", phi.adj(), 2.0 * phi_h.adj(), kRelGrad, 0.0);
  stan::math::recover_memory();
}

TEST_F(CudaShardedGLM, same_result_as_one_gpu_and_errors) {
  srand(9);
  const int N = 30011, M = 100;
  vector<int> y = counts(N, 2);
  MatrixXd x = MatrixXd::Random(N, M);
  VectorXd beta = VectorXd::Random(M) / 10.0;
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> y_d(y);
  matrix_cuda_sharded<double> x_sh(x);
  matrix_cuda_sharded<int> y_sh(y);
  const double one = stan::math::bernoulli_logit_glm_lpmf(y_d, x_d, 0.1, beta);
  const double many = stan::math::bernoulli_logit_glm_lpmf(y_sh, x_sh, 0.1, beta);
  EXPECT_NEAR(many, one, 1e-12 * std::fabs(one));
  // operands must be sharded alike
  EXPECT_THROW(stan::math::bernoulli_logit_glm_lpmf(y_d, x_sh, 0.1, beta),
               std::invalid_argument);
  EXPECT_THROW(stan::math::bernoulli_logit_glm_lpmf(y_sh, x_d, 0.1, beta),
               std::invalid_argument);
  // value errors are found on whichever shard holds them
  vector<int> y_bad = y;
  y_bad[N - 1] = 2;
  EXPECT_THROW(stan::math::bernoulli_logit_glm_lpmf(matrix_cuda_sharded<int>(y_bad), x_sh, 0.1,
                                                    beta),
               std::domain_error);
  MatrixXd x_bad = x;
  x_bad(N - 2, 3) = NAN;
  EXPECT_THROW(stan::math::bernoulli_logit_glm_lpmf(y_sh, matrix_cuda_sharded<double>(x_bad), 0.1,
                                                    beta),
               std::domain_error);
  // a host vector operand is scattered like x on the way in
  EXPECT_NEAR(stan::math::bernoulli_logit_glm_lpmf(y, x_sh, 0.1, beta), one,
              1e-12 * std::fabs(one));
}
