// End-to-end timing of the C++ drop-in itself: the call a Stan model makes,
//   var lp = stan::math::bernoulli_logit_glm_lpmf(y_cuda, x_cuda, alpha, beta);
//   lp.grad();  ... read the adjoints ...  recover_memory();
// with alpha a var and beta an Eigen vector of var on the reference's own tape,
// x / y resident on the device (uploaded once, as in the model constructor).
// Every iteration includes: host parameters -> kernel arguments, the fused
// kernel, packed result -> pinned host memory, make_partials_propagator edges,
// the reverse sweep over the tape and the arena reset.  Wall-clock timed.
// usage: glm_bench N K steps warmup [mode]  -> one JSON line on stdout
//   mode bernoulli (default)  BASELINE config 2, the call above
//   mode normal               config 1: normal_id_glm_lpdf, alpha / beta / sigma var
//   mode negbin_xvar          config 4: neg_binomial_2_log_glm_lpmf with x a
//        var_value<matrix_cuda> (value and N x K adjoint resident in HBM), beta and
//        phi var; every iteration runs the forward sweep, grad() -- whose reverse
//        sweep x.adj() += lp.adj() * d beta^T touches the N x K adjoint once -- reads
//        one row of that adjoint back, and recover_memory()
// (bench.py runs it on rank 0 and reports it as e2e.cpp_drop_in.)
#include <stan/math.hpp>
#include <stan/math/cuda.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>

int main(int argc, char** argv) {
  using stan::math::matrix_cuda;
  using stan::math::var;
  const long N = argc > 1 ? atol(argv[1]) : 10000000;
  const int K = argc > 2 ? atoi(argv[2]) : 256;
  const int steps = argc > 3 ? atoi(argv[3]) : 30;
  const int warmup = argc > 4 ? atoi(argv[4]) : 5;
  const std::string mode = argc > 5 ? argv[5] : "bernoulli";
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
    double lp_val = 0, g_alpha = 0, g_beta0 = 0, g_x00 = 0;
    const char* call
        = "stan::math::bernoulli_logit_glm_lpmf(matrix_cuda y, matrix_cuda x, var "
          "alpha, Matrix<var> beta) + grad() + recover_memory()";
    std::function<void()> step = [&]() {
      var alpha = 0.1;
      Eigen::Matrix<var, Eigen::Dynamic, 1> beta = beta0;
      var lp = stan::math::bernoulli_logit_glm_lpmf(y, x, alpha, beta);
      lp.grad();
      lp_val = lp.val();
      g_alpha = alpha.adj();
      g_beta0 = beta[0].adj();
      stan::math::recover_memory();
    };
    matrix_cuda<double> y_real;
    if (mode == "normal") {
      y_real = matrix_cuda<double>(N, 1);
      stan::math::check_cuda_status(
          "glm_bench", smc_matrix_fill_synthetic(y_real.handle(), 5, 0, 0, 2.0, 0, 0));
      call = "stan::math::normal_id_glm_lpdf(matrix_cuda y, matrix_cuda x, var alpha, "
             "Matrix<var> beta, var sigma) + grad() + recover_memory()";
      step = [&]() {
        var alpha = 0.1, sigma = 1.3;
        Eigen::Matrix<var, Eigen::Dynamic, 1> beta = beta0;
        var lp = stan::math::normal_id_glm_lpdf(y_real, x, alpha, beta, sigma);
        lp.grad();
        lp_val = lp.val();
        g_alpha = alpha.adj();
        g_beta0 = beta[0].adj();
        stan::math::recover_memory();
      };
    } else if (mode == "negbin_xvar") {
      stan::math::check_cuda_status(
          "glm_bench", smc_matrix_fill_synthetic(y.handle(), 777, 0, 1, 1.0, 0, 4));
      call = "stan::math::neg_binomial_2_log_glm_lpmf(matrix_cuda y, var_value<matrix_cuda> "
             "x, var alpha, Matrix<var> beta, var phi) + grad() + one row of x.adj() read "
             "back + recover_memory()";
      step = [&]() {
        stan::math::var_value<matrix_cuda<double>> xv(x);  // view: no copy
        var alpha = 0.1, phi = 2.5;
        Eigen::Matrix<var, Eigen::Dynamic, 1> beta = beta0;
        var lp = stan::math::neg_binomial_2_log_glm_lpmf(y, xv, alpha, beta, phi);
        lp.grad();
        lp_val = lp.val();
        g_alpha = alpha.adj();
        g_beta0 = beta[0].adj();
        Eigen::VectorXd row0(K);  // waits for the reverse sweep on the device
        stan::math::check_cuda_status(
            "glm_bench", smc_matrix_download_rows(xv.adj().handle(), 0, 1, row0.data(), 1));
        g_x00 = row0[0];
        stan::math::recover_memory();
      };
    } else if (mode != "bernoulli") {
      std::fprintf(stderr, "glm_bench: unknown mode %s\n", mode.c_str());
      return 2;
    }
    for (int i = 0; i < warmup; ++i) step();
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < steps; ++i) step();
    const double sec
        = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf(
        "{\"call\": \"%s\", \"mode\": \"%s\", \"rows\": %ld, \"cols\": %d, "
        "\"steps\": %d, \"ms_per_eval\": %.6f, \"evals_per_s\": %.4f, \"logp_per_row\": %.12f, "
        "\"d_alpha\": %.9e, \"d_beta0\": %.9e, \"d_x00\": %.9e}\n",
        call, mode.c_str(), N, K, steps, sec / steps * 1e3, steps / sec, lp_val / N, g_alpha,
        g_beta0, g_x00);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "glm_bench: %s\n", e.what());
    return 1;
  }
  return 0;
}
