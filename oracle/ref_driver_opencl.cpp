// oracle/_ref/libstan_ref_cl.so -- the UNMODIFIED reference compiled with
// -DSTAN_OPENCL: its own OpenCL device path (stan/math/opencl/prim/
// bernoulli_logit_glm_lpmf.hpp L52-167 and siblings; flags make/compiler_flags
// L238-250) timed on the B200 through whatever OpenCL ICD the box has.  Test / bench
// infrastructure, never the product: bench.py's `cpu_baseline.opencl` leg only.
// This file only CALLS stan::math; nothing of the reference is copied.
//
// x and y are uploaded once (matrix_cl), as a Stan model does in its constructor;
// every timed iteration builds var alpha / beta, evaluates the GLM on the device,
// runs grad() and reads the beta adjoints back.
#include <stan/math.hpp>

#include <chrono>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

using Eigen::Map;
using Eigen::MatrixXd;
using Eigen::VectorXd;
using stan::math::matrix_cl;
using stan::math::var;

namespace {
std::string g_error;
}

extern "C" {

const char* ref_cl_last_error(void) { return g_error.c_str(); }

// Name of the OpenCL device the reference selected; 0 on success.
int ref_cl_device_name(char* buf, int len) {
  try {
    const std::string name
        = stan::math::opencl_context.device()[0].getInfo<CL_DEVICE_NAME>() + " / "
          + stan::math::opencl_context.platform()[0].getInfo<CL_PLATFORM_NAME>();
    std::strncpy(buf, name.c_str(), len - 1);
    buf[len - 1] = 0;
    return 0;
  } catch (const std::exception& e) {
    g_error = e.what();
    return 1;
  }
}

// family: 1 bernoulli_logit, 2 poisson_log.  Seconds per lpdf+grad evaluation (best
// of reps), or a negative number on failure (see ref_cl_last_error).
double ref_time_glm_opencl(int family, long N, long K, const int* y, const double* x,
                           double alpha, const double* beta, int reps, double* logp_out,
                           double* d_beta_out) {
  try {
    Map<const MatrixXd> xm(x, N, K);
    Map<const VectorXd> bm(beta, K);
    std::vector<int> yi(y, y + N);
    matrix_cl<double> x_cl(xm);
    matrix_cl<int> y_cl(yi);
    x_cl.wait_for_read_write_events();
    double best = 1e300;
    for (int r = 0; r < reps; ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      {
        Eigen::Matrix<var, Eigen::Dynamic, 1> b = bm.cast<var>();
        var a = alpha;
        auto b_cl = stan::math::to_matrix_cl(b);
        var lp = family == 1 ? stan::math::bernoulli_logit_glm_lpmf(y_cl, x_cl, a, b_cl)
                             : stan::math::poisson_log_glm_lpmf(y_cl, x_cl, a, b_cl);
        lp.grad();
        if (logp_out) *logp_out = lp.val();
        for (long k = 0; k < K; ++k) {
          const double g = b(k).adj();
          if (d_beta_out) d_beta_out[k] = g;
        }
      }
      stan::math::recover_memory();
      const double dt
          = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (dt < best) best = dt;
    }
    return best;
  } catch (const std::exception& e) {
    g_error = e.what();
    return -1.0;
  }
}

}  // extern "C"
