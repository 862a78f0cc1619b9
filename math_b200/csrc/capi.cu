// extern "C" entry points of the GLM families: argument validation in the
// reference's order (sizes -> std::invalid_argument, values -> std::domain_error),
// the include_summand / size_zero early returns, launch, host synchronisation,
// lazy finiteness checks, and unpacking of the packed device result.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "smc_internal.h"

using namespace smc;

namespace {

bool host_all_finite(const double* p, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    if (!std::isfinite(p[i])) return false;
  return true;
}

// Shape checks shared by every family (check_consistent_size in the reference).
int check_shapes(const char* fn, const smc_matrix* x, const smc_matrix* y,
                 int y_dtype, const smc_matrix* v1, const smc_matrix* v2) {
  if (!x || x->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: x must be an f64 device matrix", fn);
  const int64_t N = x->rows;
  if (y) {
    if (y->dtype != y_dtype)
      return fail(SMC_ERR_INVALID_ARGUMENT, "%s: y has the wrong dtype", fn);
    if (y->rows * y->cols != N || !vec_contiguous(y))
      return fail(SMC_ERR_INVALID_ARGUMENT,
                  "%s: size of y (%lld) does not match rows of x (%lld)", fn,
                  (long long)(y->rows * y->cols), (long long)N);
  }
  const smc_matrix* vs[2] = {v1, v2};
  for (const smc_matrix* v : vs) {
    if (!v) continue;
    if (v->dtype != SMC_F64 || v->rows * v->cols != N || !vec_contiguous(v))
      return fail(SMC_ERR_INVALID_ARGUMENT,
                  "%s: size of a per-row vector (%lld) does not match rows of x "
                  "(%lld)",
                  fn, (long long)(v->rows * v->cols), (long long)N);
  }
  return SMC_OK;
}

int check_out_vec(const char* fn, const smc_matrix* m, int64_t N) {
  if (m && (m->dtype != SMC_F64 || m->rows * m->cols != N || !vec_contiguous(m)))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: per-row output has the wrong size",
                fn);
  return SMC_OK;
}

int check_dx(const char* fn, unsigned flags, const smc_matrix* x,
             const smc_matrix* d_x) {
  if (!(flags & SMC_VAR_X)) return SMC_OK;
  if (flags & SMC_DX_FACTORED) {
    if (!d_x || d_x->dtype != SMC_F64 || d_x->rows * d_x->cols != x->rows
        || !vec_contiguous(d_x))
      return fail(SMC_ERR_INVALID_ARGUMENT,
                  "%s: SMC_DX_FACTORED needs d_x to be an f64 vector of rows(x) doubles",
                  fn);
    return SMC_OK;
  }
  if (!d_x || d_x->dtype != SMC_F64 || d_x->rows != x->rows
      || d_x->cols != x->cols)
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "%s: SMC_VAR_X needs d_x with the shape of x", fn);
  return SMC_OK;
}

}  // namespace

namespace smc {
int64_t spin_budget_us() {
  static const int64_t us = [] {
    const char* e = getenv("SMC_SPIN_US");
    return e ? (int64_t)atoll(e) : (int64_t)60;
  }();
  return us;
}

// Runs the call with the packed result in pinned host memory; returns it.
int run_sync(GlmCall& c, int n_out, const double** out) {
  if (is_sharded(c.x)) return run_sharded(c, n_out, out);
  for (const smc_matrix* m : {c.y, c.alpha_vec, c.aux_vec, (const smc_matrix*)c.d_alpha_vec,
                              (const smc_matrix*)c.d_aux_vec, (const smc_matrix*)c.d_y_vec,
                              (const smc_matrix*)c.d_x})
    if (is_sharded(m))
      return fail(SMC_ERR_INVALID_ARGUMENT,
                  "a per-row operand is sharded but x is not (shard x too)");
  Context& cx = ctx();
  if (int rc = ensure_out(sizeof(double) * ((size_t)n_out + 1))) return rc;
  c.out = cx.out_host;
  // Completion flag behind the packed result.  The fused kernel stores it last;
  // polling it for a short while saves the wake-up latency of a stream
  // synchronise, which is most of the call for small N (a whole evaluation of
  // N = 1e4, K = 100 takes ~10 us on the GPU).  Long kernels fall through to the
  // blocking synchronise, so a waiting chain does not burn a core for milliseconds
  // (SMC_SPIN_US, default 60; polling for the whole of a 2.9 ms evaluation was measured
  // and buys nothing, profiles/r02/r02_spin_budget.txt).
  volatile unsigned long long* flag
      = reinterpret_cast<volatile unsigned long long*>(cx.out_host + n_out);
  const unsigned long long seq = ++cx.sync_seq;
  *flag = 0;
  c.done_flag = const_cast<unsigned long long*>(flag);
  c.done_val = seq;
  cx.flag_armed = false;
  if (int rc = launch_glm(c)) return rc;
  bool done = false;
  if (cx.flag_armed) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int spin = 0; !done; ++spin) {
      done = *flag == seq;
      if (!done && (spin & 63) == 63
          && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_budget_us()))
        break;
    }
  }
  if (!done) SMC_CUDA(cudaStreamSynchronize(cx.stream));
  *out = cx.out_host;
  return SMC_OK;
}
}  // namespace smc

namespace {

int y_bounds(const char* fn, const smc_matrix* y, double y_scalar, int lo, int hi,
             bool check_hi) {
  int mn, mx;
  if (y) {
    if (y->rows * y->cols == 0) return SMC_OK;
    if (int rc = y_range(y, &mn, &mx)) return rc;
  } else {
    mn = mx = (int)y_scalar;
  }
  if (mn < lo || (check_hi && mx > hi))
    return fail(SMC_ERR_DOMAIN, "%s: dependent variable out of range [%d, %d]",
                fn, mn, mx);
  return SMC_OK;
}

}  // namespace

extern "C" {

int smc_bernoulli_logit_glm(const smc_matrix* y, int y_scalar,
                            const smc_matrix* x, const smc_matrix* alpha_vec,
                            double alpha, const double* beta, unsigned flags,
                            double* logp, double* d_alpha,
                            smc_matrix* d_alpha_vec, double* d_beta,
                            smc_matrix* d_x) {
  static const char* fn = "bernoulli_logit_glm_lpmf";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = check_shapes(fn, x, y, SMC_I32, alpha_vec, nullptr)) return rc;
  const int64_t N = x->rows, K = x->cols;
  if ((K > 0 && !beta) || !logp)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL beta or logp", fn);
  if (int rc = check_out_vec(fn, d_alpha_vec, N)) return rc;
  if (int rc = check_dx(fn, flags, x, d_x)) return rc;
  *logp = 0.0;
  if (N == 0) return SMC_OK;  // size_zero(y), L80-82
  if (int rc = y_bounds(fn, y, y_scalar, 0, 1, true)) return rc;  // L85
  if ((flags & SMC_PROPTO)
      && !(flags & (SMC_VAR_X | SMC_VAR_ALPHA | SMC_VAR_BETA)))
    return SMC_OK;  // include_summand, L87-89
  GlmCall c;
  c.family = kBernoulli;
  c.x = x;
  c.y = y;
  c.y_scalar = y_scalar;
  c.alpha_vec = alpha_vec;
  c.alpha = alpha;
  c.beta_host = beta;
  c.flags = flags;
  c.d_alpha_vec = (flags & SMC_VAR_ALPHA) ? d_alpha_vec : nullptr;
  c.d_x = d_x;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER + (int)K, &o)) return rc;
  if (!std::isfinite(o[SMC_OUT_LOGP])) {  // lazy checks, L128-132
    if (!host_all_finite(beta, K))
      return fail(SMC_ERR_DOMAIN, "%s: Weight vector is not finite", fn);
    if (!alpha_vec && !std::isfinite(alpha))
      return fail(SMC_ERR_DOMAIN, "%s: Intercept is not finite", fn);
    if (o[SMC_OUT_NONFINITE] > 0)
      return fail(SMC_ERR_DOMAIN,
                  "%s: Matrix of independent variables is not finite", fn);
  }
  *logp = o[SMC_OUT_LOGP];
  if (d_alpha && (flags & SMC_VAR_ALPHA)) *d_alpha = o[SMC_OUT_SUM_D];
  if (d_beta && (flags & SMC_VAR_BETA))
    memcpy(d_beta, o + SMC_OUT_HEADER, sizeof(double) * K);
  return SMC_OK;
}

int smc_binomial_logit_glm(const smc_matrix* n, int n_scalar,
                           const smc_matrix* trials, int trials_scalar,
                           const smc_matrix* x, const smc_matrix* alpha_vec,
                           double alpha, const double* beta, unsigned flags,
                           double* logp, double* d_alpha,
                           smc_matrix* d_alpha_vec, double* d_beta,
                           smc_matrix* d_x) {
  static const char* fn = "binomial_logit_glm_lpmf";
  if (int rc = ensure_ctx()) return rc;
  if (!x || x->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: x must be an f64 device matrix", fn);
  if (!logp) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL logp", fn);
  const int64_t N = x->rows, K = x->cols;
  *logp = 0.0;
  // size_zero(n, N, alpha, beta, x) comes BEFORE the size checks, L76-78
  if (N == 0 || K == 0 || (n && n->rows * n->cols == 0)
      || (trials && trials->rows * trials->cols == 0)
      || (alpha_vec && alpha_vec->rows * alpha_vec->cols == 0))
    return SMC_OK;
  if ((flags & SMC_PROPTO)
      && !(flags & (SMC_VAR_X | SMC_VAR_ALPHA | SMC_VAR_BETA)))
    return SMC_OK;  // include_summand, L80-82
  if (int rc = check_shapes(fn, x, n, SMC_I32, alpha_vec, nullptr)) return rc;  // L88-93
  if (trials
      && (trials->dtype != SMC_I32 || trials->rows * trials->cols != N
          || !vec_contiguous(trials)))
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "%s: size of the population size parameter (%lld) does not match "
                "rows of x (%lld)",
                fn, (long long)(trials->rows * trials->cols), (long long)N);
  if (!beta) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL beta", fn);
  if (int rc = check_out_vec(fn, d_alpha_vec, N)) return rc;
  if (int rc = check_dx(fn, flags, x, d_x)) return rc;
  {
    // check_bounded(n, 0, N), check_nonnegative(N), L98-99 (n, N are data: cached)
    bool ok = true;
    double bc = 0.0;
    if (int rc = binom_stats(n, n_scalar, trials, trials_scalar,
                             (n || trials) ? N : 1, &ok, &bc))
      return rc;
    if (!ok)
      return fail(SMC_ERR_DOMAIN,
                  "%s: Successes variable must be in the interval [0, N] and the "
                  "population size nonnegative",
                  fn);
  }
  GlmCall c;
  c.family = kBinomial;
  c.x = x;
  c.y = n;
  c.y_scalar = n_scalar;
  c.aux_vec = trials;
  c.aux = trials_scalar;
  c.alpha_vec = alpha_vec;
  c.alpha = alpha;
  c.beta_host = beta;
  c.flags = flags;
  c.d_alpha_vec = (flags & SMC_VAR_ALPHA) ? d_alpha_vec : nullptr;
  c.d_x = d_x;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER + (int)K, &o)) return rc;
  if (!std::isfinite(o[SMC_OUT_LOGP])) {  // lazy checks, L118-122
    if (!host_all_finite(beta, K))
      return fail(SMC_ERR_DOMAIN, "%s: Weight vector is not finite", fn);
    if (!alpha_vec && !std::isfinite(alpha))
      return fail(SMC_ERR_DOMAIN, "%s: Intercept is not finite", fn);
    std::vector<double> keep(o, o + SMC_OUT_HEADER + K);
    int ok = 1;
    if (alpha_vec) {
      if (int rc = smc_matrix_all_finite(alpha_vec, &ok)) return rc;
      if (!ok) return fail(SMC_ERR_DOMAIN, "%s: Intercept is not finite", fn);
    }
    if (int rc = smc_matrix_all_finite(x, &ok)) return rc;
    if (!ok)
      return fail(SMC_ERR_DOMAIN,
                  "%s: Matrix of independent variables is not finite", fn);
    *logp = keep[SMC_OUT_LOGP];
    if (d_alpha && (flags & SMC_VAR_ALPHA)) *d_alpha = keep[SMC_OUT_SUM_D];
    if (d_beta && (flags & SMC_VAR_BETA))
      memcpy(d_beta, keep.data() + SMC_OUT_HEADER, sizeof(double) * K);
    return SMC_OK;
  }
  *logp = o[SMC_OUT_LOGP];
  if (d_alpha && (flags & SMC_VAR_ALPHA)) *d_alpha = o[SMC_OUT_SUM_D];
  if (d_beta && (flags & SMC_VAR_BETA))
    memcpy(d_beta, o + SMC_OUT_HEADER, sizeof(double) * K);
  return SMC_OK;
}

int smc_poisson_log_glm(const smc_matrix* y, int y_scalar, const smc_matrix* x,
                        const smc_matrix* alpha_vec, double alpha,
                        const double* beta, unsigned flags, double* logp,
                        double* d_alpha, smc_matrix* d_alpha_vec,
                        double* d_beta, smc_matrix* d_x) {
  static const char* fn = "poisson_log_glm_lpmf";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = check_shapes(fn, x, y, SMC_I32, alpha_vec, nullptr)) return rc;
  const int64_t N = x->rows, K = x->cols;
  if ((K > 0 && !beta) || !logp)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL beta or logp", fn);
  if (int rc = check_out_vec(fn, d_alpha_vec, N)) return rc;
  if (int rc = check_dx(fn, flags, x, d_x)) return rc;
  *logp = 0.0;
  if (int rc = y_bounds(fn, y, y_scalar, 0, 0, false)) return rc;  // L84
  if (N == 0) return SMC_OK;                                       // L86-88
  if ((flags & SMC_PROPTO)
      && !(flags & (SMC_VAR_X | SMC_VAR_ALPHA | SMC_VAR_BETA)))
    return SMC_OK;  // L89-91
  GlmCall c;
  c.family = kPoisson;
  c.x = x;
  c.y = y;
  c.y_scalar = y_scalar;
  c.alpha_vec = alpha_vec;
  c.alpha = alpha;
  c.beta_host = beta;
  c.flags = flags;
  c.d_alpha_vec = (flags & SMC_VAR_ALPHA) ? d_alpha_vec : nullptr;
  c.d_x = d_x;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER + (int)K, &o)) return rc;
  // Lazy checks.  prim (L120-124) runs them when the sum of the derivatives is not
  // finite; the reference's device overload also when any linear predictor is
  // (opencl/prim/poisson_log_glm_lpmf.hpp L97-98, L112-121: a -inf entry of x throws there
  // and returns -inf in prim).  A device backend is held to the device overload's tests
  // (test/unit/math/opencl/rev/poisson_log_glm_lpmf_test.cpp L83-85): its rule.
  if (!std::isfinite(o[SMC_OUT_SUM_D]) || o[SMC_OUT_NONFINITE] > 0) {
    if (!host_all_finite(beta, K))
      return fail(SMC_ERR_DOMAIN, "%s: Weight vector is not finite", fn);
    if (!alpha_vec && !std::isfinite(alpha))
      return fail(SMC_ERR_DOMAIN, "%s: Intercept is not finite", fn);
    std::vector<double> keep(o, o + SMC_OUT_HEADER + K);  // (the scans reuse the result buffer)
    int ok = 1;
    if (alpha_vec) {
      if (int rc = smc_matrix_all_finite(alpha_vec, &ok)) return rc;
      if (!ok) return fail(SMC_ERR_DOMAIN, "%s: Intercept is not finite", fn);
    }
    if (int rc = smc_matrix_all_finite(x, &ok)) return rc;
    if (!ok || !std::isfinite(keep[SMC_OUT_SUM_D]))
      return fail(SMC_ERR_DOMAIN,
                  "%s: Matrix of independent variables is not finite", fn);
    *logp = keep[SMC_OUT_LOGP];
    if (d_alpha && (flags & SMC_VAR_ALPHA)) *d_alpha = keep[SMC_OUT_SUM_D];
    if (d_beta && (flags & SMC_VAR_BETA))
      memcpy(d_beta, keep.data() + SMC_OUT_HEADER, sizeof(double) * K);
    return SMC_OK;
  }
  *logp = o[SMC_OUT_LOGP];
  if (d_alpha && (flags & SMC_VAR_ALPHA)) *d_alpha = o[SMC_OUT_SUM_D];
  if (d_beta && (flags & SMC_VAR_BETA))
    memcpy(d_beta, o + SMC_OUT_HEADER, sizeof(double) * K);
  return SMC_OK;
}

int smc_normal_id_glm(const smc_matrix* y, double y_scalar, const smc_matrix* x,
                      const smc_matrix* alpha_vec, double alpha,
                      const double* beta, const smc_matrix* sigma_vec,
                      double sigma, unsigned flags, double* logp,
                      double* d_alpha, smc_matrix* d_alpha_vec, double* d_beta,
                      double* d_sigma, smc_matrix* d_sigma_vec, double* d_y,
                      smc_matrix* d_y_vec, smc_matrix* d_x) {
  static const char* fn = "normal_id_glm_lpdf";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = check_shapes(fn, x, y, SMC_F64, alpha_vec, sigma_vec)) return rc;
  const int64_t N = x->rows, K = x->cols;
  if ((K > 0 && !beta) || !logp)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL beta or logp", fn);
  if (int rc = check_out_vec(fn, d_alpha_vec, N)) return rc;
  if (int rc = check_out_vec(fn, d_sigma_vec, N)) return rc;
  if (int rc = check_out_vec(fn, d_y_vec, N)) return rc;
  if (int rc = check_dx(fn, flags, x, d_x)) return rc;
  *logp = 0.0;
  if (!sigma_vec && (!(sigma > 0.0) || !std::isfinite(sigma)))  // L93
    return fail(SMC_ERR_DOMAIN, "%s: Scale vector is %g, but must be positive finite",
                fn, sigma);
  if (N == 0) return SMC_OK;  // L95-97
  if ((flags & SMC_PROPTO)
      && !(flags
           & (SMC_VAR_X | SMC_VAR_ALPHA | SMC_VAR_BETA | SMC_VAR_AUX | SMC_VAR_Y)))
    return SMC_OK;  // L98-100
  GlmCall c;
  c.family = kNormal;
  c.x = x;
  c.y = y;
  c.y_scalar = y_scalar;
  c.alpha_vec = alpha_vec;
  c.alpha = alpha;
  c.aux_vec = sigma_vec;
  c.aux = sigma;
  c.beta_host = beta;
  c.flags = flags;
  c.d_alpha_vec = (flags & SMC_VAR_ALPHA) ? d_alpha_vec : nullptr;
  c.d_aux_vec = (flags & SMC_VAR_AUX) ? d_sigma_vec : nullptr;
  c.d_y_vec = (flags & SMC_VAR_Y) ? d_y_vec : nullptr;
  c.d_x = d_x;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER + (int)K, &o)) return rc;
  if (o[SMC_OUT_NONFINITE] > 0)
    return fail(SMC_ERR_DOMAIN, "%s: Scale vector must be positive finite", fn);
  if (!std::isfinite(o[SMC_OUT_AUX2]))  // L192-198: the last check always throws
    return fail(SMC_ERR_DOMAIN,
                "%s: y, alpha, beta or the matrix of independent variables is "
                "not finite",
                fn);
  *logp = o[SMC_OUT_LOGP];
  if (d_alpha && (flags & SMC_VAR_ALPHA)) *d_alpha = o[SMC_OUT_SUM_D];
  if (d_y && (flags & SMC_VAR_Y)) *d_y = -o[SMC_OUT_SUM_D];  // L141-147
  if (d_sigma && (flags & SMC_VAR_AUX) && !sigma_vec)
    *d_sigma = (o[SMC_OUT_AUX2] - (double)N) * (1.0 / sigma);  // L179-183
  if (d_beta && (flags & SMC_VAR_BETA))
    memcpy(d_beta, o + SMC_OUT_HEADER, sizeof(double) * K);
  return SMC_OK;
}

int smc_neg_binomial_2_log_glm(const smc_matrix* y, int y_scalar,
                               const smc_matrix* x, const smc_matrix* alpha_vec,
                               double alpha, const double* beta,
                               const smc_matrix* phi_vec, double phi,
                               unsigned flags, double* logp, double* d_alpha,
                               smc_matrix* d_alpha_vec, double* d_beta,
                               double* d_phi, smc_matrix* d_phi_vec,
                               smc_matrix* d_x) {
  static const char* fn = "neg_binomial_2_log_glm_lpmf";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = check_shapes(fn, x, y, SMC_I32, alpha_vec, phi_vec)) return rc;
  const int64_t N = x->rows, K = x->cols;
  if ((K > 0 && !beta) || !logp)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL beta or logp", fn);
  if (int rc = check_out_vec(fn, d_alpha_vec, N)) return rc;
  if (int rc = check_out_vec(fn, d_phi_vec, N)) return rc;
  if (int rc = check_dx(fn, flags, x, d_x)) return rc;
  *logp = 0.0;
  if (!host_all_finite(beta, K))  // L114
    return fail(SMC_ERR_DOMAIN, "%s: Weight vector is not finite", fn);
  if (!alpha_vec && !std::isfinite(alpha))  // L115
    return fail(SMC_ERR_DOMAIN, "%s: Intercept is not finite", fn);
  if (N == 0) return SMC_OK;                                       // L117-119
  if (int rc = y_bounds(fn, y, y_scalar, 0, 0, false)) return rc;  // L129
  if (!phi_vec && (!(phi > 0.0) || !std::isfinite(phi)))           // L130
    return fail(SMC_ERR_DOMAIN,
                "%s: Precision parameter is %g, but must be positive finite", fn,
                phi);
  if ((flags & SMC_PROPTO)
      && !(flags & (SMC_VAR_X | SMC_VAR_ALPHA | SMC_VAR_BETA | SMC_VAR_AUX)))
    return SMC_OK;  // L132-134
  GlmCall c;
  c.family = kNegBinomial;
  c.x = x;
  c.y = y;
  c.y_scalar = y_scalar;
  c.alpha_vec = alpha_vec;
  c.alpha = alpha;
  c.aux_vec = phi_vec;
  c.aux = phi;
  c.beta_host = beta;
  c.flags = flags;
  c.d_alpha_vec = (flags & SMC_VAR_ALPHA) ? d_alpha_vec : nullptr;
  c.d_aux_vec = (flags & SMC_VAR_AUX) ? d_phi_vec : nullptr;
  c.d_x = d_x;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER + (int)K, &o)) return rc;
  if (o[SMC_OUT_NONFINITE] > 0)  // L152 (and L115 / L130 for vector alpha / phi)
    return fail(SMC_ERR_DOMAIN,
                "%s: Matrix of independent variables, intercept or precision is "
                "not finite",
                fn);
  *logp = o[SMC_OUT_LOGP];
  if (d_alpha && (flags & SMC_VAR_ALPHA)) *d_alpha = o[SMC_OUT_SUM_D];
  if (d_phi && (flags & SMC_VAR_AUX) && !phi_vec) *d_phi = o[SMC_OUT_AUX];
  if (d_beta && (flags & SMC_VAR_BETA))
    memcpy(d_beta, o + SMC_OUT_HEADER, sizeof(double) * K);
  return SMC_OK;
}

int smc_ordered_logistic_glm(const smc_matrix* y, int y_scalar,
                             const smc_matrix* x, const double* beta,
                             const double* cuts, int64_t ncuts, unsigned flags,
                             double* logp, double* d_beta, double* d_cuts,
                             smc_matrix* d_x) {
  static const char* fn = "ordered_logistic_glm_lpmf";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = check_shapes(fn, x, y, SMC_I32, nullptr, nullptr)) return rc;
  const int64_t N = x->rows, K = x->cols;
  if ((K > 0 && !beta) || !logp || ncuts < 0 || (ncuts > 0 && !cuts))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL beta, cuts or logp", fn);
  if (int rc = check_dx(fn, flags, x, d_x)) return rc;
  *logp = 0.0;
  const int64_t C = ncuts + 1;
  if (int rc = y_bounds(fn, y, y_scalar, 1, (int)C, true)) return rc;  // L82
  for (int64_t i = 1; i < ncuts; ++i)  // check_ordered, L83
    if (!(cuts[i] > cuts[i - 1]))
      return fail(SMC_ERR_DOMAIN, "%s: Cut-points are not a valid ordered vector",
                  fn);
  if (ncuts == 1 && std::isnan(cuts[0]))
    return fail(SMC_ERR_DOMAIN, "%s: Cut-points are not a valid ordered vector", fn);
  if (C > 1) {  // L84-89
    if (C > 2 && !std::isfinite(cuts[C - 2]))
      return fail(SMC_ERR_DOMAIN, "%s: Final cut-point is not finite", fn);
    if (!std::isfinite(cuts[0]))
      return fail(SMC_ERR_DOMAIN, "%s: First cut-point is not finite", fn);
  }
  if (N == 0 || ncuts == 0) return SMC_OK;  // size_zero(y, cuts), L91-93
  if ((flags & SMC_PROPTO) && !(flags & (SMC_VAR_X | SMC_VAR_BETA | SMC_VAR_AUX)))
    return SMC_OK;  // L94-96
  GlmCall c;
  c.family = kOrdered;
  c.x = x;
  c.y = y;
  c.y_scalar = y_scalar;
  c.beta_host = beta;
  c.cuts_host = cuts;
  c.ncuts = ncuts;
  c.flags = flags;
  c.d_x = d_x;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER + (int)K + (int)ncuts, &o)) return rc;
  if (!std::isfinite(o[SMC_OUT_AUX2])) {  // L124-127
    if (!host_all_finite(beta, K))
      return fail(SMC_ERR_DOMAIN, "%s: Weight vector is not finite", fn);
    int ok = 1;
    // the packed result lives in the buffer the check reuses: copy it out first
    std::vector<double> keep(o, o + SMC_OUT_HEADER + K + ncuts);
    if (int rc = smc_matrix_all_finite(x, &ok)) return rc;
    if (!ok)
      return fail(SMC_ERR_DOMAIN,
                  "%s: Matrix of independent variables is not finite", fn);
    *logp = keep[SMC_OUT_LOGP];
    if (d_beta && (flags & SMC_VAR_BETA))
      memcpy(d_beta, keep.data() + SMC_OUT_HEADER, sizeof(double) * K);
    if (d_cuts && (flags & SMC_VAR_AUX))
      memcpy(d_cuts, keep.data() + SMC_OUT_HEADER + K, sizeof(double) * ncuts);
    return SMC_OK;
  }
  *logp = o[SMC_OUT_LOGP];
  if (d_beta && (flags & SMC_VAR_BETA))
    memcpy(d_beta, o + SMC_OUT_HEADER, sizeof(double) * K);
  if (d_cuts && (flags & SMC_VAR_AUX))
    memcpy(d_cuts, o + SMC_OUT_HEADER + K, sizeof(double) * ncuts);
  return SMC_OK;
}

int smc_glm_eval_device(int family, const smc_matrix* y, double y_scalar,
                        const smc_matrix* x, const smc_matrix* alpha_vec,
                        double alpha, const smc_matrix* aux_vec, double aux,
                        const double* params_dev, int64_t ncuts, unsigned flags,
                        double* out_dev, smc_matrix* d_alpha_vec,
                        smc_matrix* d_aux_vec, smc_matrix* d_y_vec,
                        smc_matrix* d_x) {
  static const char* fn = "smc_glm_eval_device";
  if (int rc = ensure_ctx()) return rc;
  if (family < kNormal || family > kBinomial)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: unknown family %d", fn, family);
  if (int rc = check_shapes(fn, x, y, family == kNormal ? SMC_F64 : SMC_I32,
                            alpha_vec, family == kBinomial ? nullptr : aux_vec))
    return rc;
  if (family == kBinomial && aux_vec
      && (aux_vec->dtype != SMC_I32 || aux_vec->rows * aux_vec->cols != x->rows
          || !vec_contiguous(aux_vec)))
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "%s: binomial trials must be an i32 vector with one entry per row", fn);
  if (!params_dev || !out_dev)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL params_dev or out_dev", fn);
  if (int rc = check_dx(fn, flags, x, d_x)) return rc;
  if (x->rows == 0)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: empty shard", fn);
  GlmCall c;
  c.family = family;
  c.x = x;
  c.y = y;
  c.y_scalar = y_scalar;
  c.alpha_vec = alpha_vec;
  c.alpha = alpha;
  c.aux_vec = aux_vec;
  c.aux = aux;
  c.params_dev = params_dev;
  c.ncuts = family == kOrdered ? ncuts : 0;
  c.flags = flags;
  c.out = out_dev;
  c.d_alpha_vec = d_alpha_vec;
  c.d_aux_vec = d_aux_vec;
  c.d_y_vec = d_y_vec;
  c.d_x = d_x;
  return launch_glm(c);
}

}  // extern "C"
