// End-to-end timing of the C++ drop-in itself: the call a Stan model makes,
//   var lp = stan::math::bernoulli_logit_glm_lpmf(y_cuda, x_cuda, alpha, beta);
//   lp.grad();  ... read the adjoints ...  recover_memory();
// with alpha a var and beta an Eigen vector of var on the reference's own tape,
// x / y resident on the device (uploaded once, as in the model constructor).
// Every iteration includes: host parameters -> kernel arguments, the fused
// kernel, packed result -> pinned host memory, make_partials_propagator edges,
// the reverse sweep over the tape and the arena reset.  Wall-clock timed.
// usage: glm_bench N K steps warmup       -> one JSON line on stdout
// (bench.py runs it on rank 0 and reports it as e2e.cpp_drop_in.)
#include <stan/math.hpp>
#include <stan/math/cuda.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
  using stan::math::matrix_cuda;
  using stan::math::var;
  const long N = argc > 1 ? atol(argv[1]) : 10000000;
  const int K = argc > 2 ? atoi(argv[2]) : 256;
  const int steps = argc > 3 ? atoi(argv[3]) : 30;
  const int warmup = argc > 4 ? atoi(argv[4]) : 5;
  try {
    matrix_cuda<double> x(N, K);
    matrix_cuda<int> y(N, 1);
    // the synthetic inputs of bench.py (counter-based hash, generated in place)
    stan::math::check_cuda_status(
        "glm_bench", smc_matrix_fill_synthetic(x.handle(), 12345, 0, 0, 1.0, 0, 0));
    stan::math::check_cuda_status(
        "glm_bench", smc_matrix_fill_synthetic(y.handle(), 12346, 0, 1, 1.0, 0, 1));
    Eigen::VectorXd beta0(K);
    srand(12345);
    beta0 = Eigen::VectorXd::Random(K) / std::sqrt(static_cast<double>(K));
    double lp_val = 0, g_alpha = 0, g_beta0 = 0;
    auto step = [&]() {
      var alpha = 0.1;
      Eigen::Matrix<var, Eigen::Dynamic, 1> beta = beta0;
      var lp = stan::math::bernoulli_logit_glm_lpmf(y, x, alpha, beta);
      lp.grad();
      lp_val = lp.val();
      g_alpha = alpha.adj();
      g_beta0 = beta[0].adj();
      stan::math::recover_memory();
    };
    for (int i = 0; i < warmup; ++i) step();
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; ++i) step();
    const double sec
        = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf(
        "{\"call\": \"stan::math::bernoulli_logit_glm_lpmf(matrix_cuda y, matrix_cuda x, var "
        "alpha, Matrix<var> beta) + grad() + recover_memory()\", \"rows\": %ld, \"cols\": %d, "
        "\"steps\": %d, \"ms_per_eval\": %.6f, \"evals_per_s\": %.4f, \"logp_per_row\": %.12f, "
        "\"d_alpha\": %.9e, \"d_beta0\": %.9e}\n",
        N, K, steps, sec / steps * 1e3, steps / sec, lp_val / N, g_alpha, g_beta0);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "glm_bench: %s\n", e.what());
    return 1;
  }
  return 0;
}
