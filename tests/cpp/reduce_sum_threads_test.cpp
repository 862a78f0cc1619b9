// reduce_sum over TBB worker threads with the device GLM as the slice functor
// (SURVEY.md 8(f)4, 8(b) "Threading"; compiled with -DSTAN_THREADS).  Every
// worker evaluates bernoulli_logit_glm_lpmf on a row block VIEW of the one shared
// device design matrix, on its own stream and workspace (the C ABI keeps them per
// host thread) and on its own thread_local autodiff tape (rev/core/
// autodiffstackstorage.hpp L12-24); reduce_sum stitches values and adjoints
// together (rev/functor/reduce_sum.hpp).  The result must agree with one fused
// call over all rows and with the reference's prim implementation on the host.
#include "cuda_test_util.hpp"

using Eigen::Dynamic;
using Eigen::Matrix;
using Eigen::MatrixXd;
using Eigen::VectorXd;
using stan::math::matrix_cuda;
using stan::math::var;
using namespace cuda_test;  // NOLINT

namespace {
const matrix_cuda<double>* g_x = nullptr;

// slice functor: y_slice holds the responses of rows [start, end]
struct glm_slice {
  template <typename T_alpha, typename T_beta>
  auto operator()(const std::vector<int>& y_slice, std::size_t start, std::size_t end,
                  std::ostream* /*msgs*/, const T_alpha& alpha, const T_beta& beta) const {
    const int64_t n = static_cast<int64_t>(end - start + 1);
    const matrix_cuda<double>& x = *g_x;
    matrix_cuda<double> x_block = matrix_cuda<double>::view(
        static_cast<double*>(smc_matrix_data(x.handle())) + start, n, x.cols(),
        smc_matrix_ld(x.handle()));
    return stan::math::bernoulli_logit_glm_lpmf(y_slice, x_block, alpha, beta);
  }
};
}  // namespace

TEST(CudaReduceSum, device_glm_as_slice_functor_over_tbb_threads) {
  const int N = 40000, K = 48;
  stan::math::init_threadpool_tbb(4);
  srand(21);
  MatrixXd x = MatrixXd::Random(N, K);
  VectorXd beta_v = VectorXd::Random(K) / std::sqrt(K);
  std::vector<int> y(N);
  for (int i = 0; i < N; ++i) y[i] = (i * 17) % 2;
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> y_d(y);
  g_x = &x_d;

  // data only
  const double cpu = stan::math::bernoulli_logit_glm_lpmf(y, x, 0.2, beta_v);
  const double sliced = stan::math::reduce_sum<glm_slice>(y, 2048, nullptr, 0.2, beta_v);
  expect_close("value (prim)", sliced, cpu, kRelLogp, 0);

  // autodiff alpha and beta: the sliced sum, one fused device call, the host GLM
  var a1 = 0.2, a2 = 0.2, a3 = 0.2;
  Matrix<var, Dynamic, 1> b1 = beta_v, b2 = beta_v, b3 = beta_v;
  var lp_sliced = stan::math::reduce_sum<glm_slice>(y, 2048, nullptr, a1, b1);
  var lp_fused = stan::math::bernoulli_logit_glm_lpmf(y_d, x_d, a2, b2);
  var lp_cpu = stan::math::bernoulli_logit_glm_lpmf(y, x, a3, b3);
  (lp_sliced + lp_fused + lp_cpu).grad();
  expect_close("value vs fused", lp_sliced.val(), lp_fused.val(), kRelLogp, 0);
  expect_close("value vs host", lp_sliced.val(), lp_cpu.val(), kRelLogp, 0);
  expect_close("d_alpha", a1.adj(), a3.adj(), kRelGrad, std::fabs(a3.adj()));
  const double scale = b3.adj().cwiseAbs().maxCoeff();
  for (int k = 0; k < K; ++k) {
    expect_close("d_beta vs host", b1[k].adj(), b3[k].adj(), kRelGrad, scale);
    expect_close("d_beta vs fused", b1[k].adj(), b2[k].adj(), kRelGrad, scale);
  }
  stan::math::recover_memory();
  g_x = nullptr;
}
