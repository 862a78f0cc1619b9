// The step either side of the fused GLMs: models that add
// terms to the linear predictor before the likelihood cannot call a *_glm_*
// function; they form theta on the device and call the un-fused density on it.
//
//   smc_linear_predictor          theta = x beta (+ alpha)       one sweep over x
//   smc_linear_predictor_adjoint  x^T v, sum v                   one sweep over x
//   smc_<family>_lpmf             value + d/dtheta on an N-vector theta
//
// The products run through the fused TMA kernel (family kLinear), so they stream
// x at the same rate as the GLMs; the densities run the GLM link functions with
// K = 0 (glm_link.cuh follows the same formulas as prim/prob/<family>_lpmf.hpp:
// bernoulli_logit_lpmf.hpp L60-95, poisson_log_lpmf.hpp L49-96,
// neg_binomial_2_log_lpmf.hpp L50-130, ordered_logistic_lpmf.hpp L93-200).
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "glm_link.cuh"
#include "smc_internal.h"

using namespace smc;

namespace {

// An N x 0 design matrix: the GLM machinery with no columns.
smc_matrix empty_design(int64_t N) {
  smc_matrix x;
  x.rows = N;
  x.cols = 0;
  x.ld = N > 0 ? N : 1;
  x.dtype = SMC_F64;
  return x;
}

bool is_vec(const smc_matrix* m) { return vec_contiguous(m); }

// theta must be an f64 vector; y (if a vector) and the other per-row operands
// must have its length.
int theta_shapes(const char* fn, const smc_matrix* theta, const smc_matrix* y,
                 const smc_matrix* v1, smc_matrix* d_theta, smc_matrix* d_v1,
                 int64_t* N) {
  if (!theta || theta->dtype != SMC_F64 || !is_vec(theta))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: theta must be an f64 device vector", fn);
  *N = theta->rows * theta->cols;
  if (y && (y->dtype != SMC_I32 || y->rows * y->cols != *N || !is_vec(y)))
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "%s: size of the random variable (%lld) does not match the size of "
                "the parameter (%lld)",
                fn, (long long)(y->rows * y->cols), (long long)*N);
  const smc_matrix* vs[3] = {v1, d_theta, d_v1};
  for (const smc_matrix* v : vs)
    if (v && (v->dtype != SMC_F64 || v->rows * v->cols != *N || !is_vec(v)))
      return fail(SMC_ERR_INVALID_ARGUMENT,
                  "%s: size of a per-row vector (%lld) does not match the size of "
                  "the parameter (%lld)",
                  fn, (long long)(v->rows * v->cols), (long long)*N);
  return SMC_OK;
}

int int_bounds(const char* fn, const smc_matrix* y, int y_scalar, int lo, int hi,
               bool check_hi) {
  int mn = y_scalar, mx = y_scalar;
  if (y) {
    if (y->rows * y->cols == 0) return SMC_OK;
    if (int rc = y_range(y, &mn, &mx)) return rc;
  }
  if (mn < lo || (check_hi && mx > hi))
    return fail(SMC_ERR_DOMAIN, "%s: Random variable is out of range", fn);
  return SMC_OK;
}

// The lazy value checks behind a non-finite result: which kind of non-finite value
// does a device vector hold?  bit 0: a NaN; bit 1: +inf; bit 2: -inf in a row whose
// count n_i is not 0 (poisson_log_lpmf.hpp L56-66).  One scan on the device; only
// the three flags come back.
constexpr int kHasNan = 1, kHasPosInf = 2, kHasNegInfNonzeroCount = 4;
__global__ void classify_kernel(const double* __restrict__ v, const int* __restrict__ n,
                                int n_scalar, int64_t N, int* __restrict__ out) {
  int f = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double t = v[i];
    if (t != t) f |= kHasNan;
    if (t == INFINITY) f |= kHasPosInf;
    if (t == -INFINITY && (n ? n[i] : n_scalar) != 0) f |= kHasNegInfNonzeroCount;
  }
  f = __reduce_or_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && f) atomicOr(out, f);
}

int classify(const smc_matrix* v, const smc_matrix* n, int n_scalar, int* flags_out) {
  Context& c = ctx();
  const int64_t N = v->rows * v->cols;
  *flags_out = 0;
  if (N == 0) return SMC_OK;
  if (int rc = ensure_out(4096)) return rc;
  int* flag = reinterpret_cast<int*>(c.out_host);
  *flag = 0;
  int64_t blocks = (N + 255) / 256;
  if (blocks > c.sm_count * 16) blocks = c.sm_count * 16;
  classify_kernel<<<(int)blocks, 256, 0, c.stream>>>(
      static_cast<const double*>(v->data), n ? static_cast<const int*>(n->data) : nullptr,
      n_scalar, N, flag);
  SMC_CUDA(cudaGetLastError());
  c.launches += 1;
  SMC_CUDA(cudaStreamSynchronize(c.stream));
  *flags_out = *flag;
  return SMC_OK;
}

}  // namespace

extern "C" {

int smc_linear_predictor(const smc_matrix* x, const double* beta,
                         const smc_matrix* alpha_vec, double alpha,
                         smc_matrix* theta_out) {
  static const char* fn = "linear_predictor";
  if (int rc = ensure_ctx()) return rc;
  if (!x || x->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: x must be an f64 device matrix", fn);
  const int64_t N = x->rows, K = x->cols;
  if (K > 0 && !beta) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL beta", fn);
  if (!theta_out || theta_out->dtype != SMC_F64 || theta_out->rows * theta_out->cols != N
      || !is_vec(theta_out))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: theta_out must hold rows(x) doubles", fn);
  if (alpha_vec && (alpha_vec->dtype != SMC_F64 || alpha_vec->rows * alpha_vec->cols != N
                    || !is_vec(alpha_vec)))
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "%s: size of alpha (%lld) does not match rows of x (%lld)", fn,
                (long long)(alpha_vec->rows * alpha_vec->cols), (long long)N);
  if (N == 0) return SMC_OK;
  GlmCall c;
  c.family = kLinear;
  c.x = x;
  c.alpha_vec = alpha_vec;
  c.alpha = alpha;
  c.beta_host = beta;
  c.flags = 0;
  c.d_alpha_vec = theta_out;
  theta_out->version++;
  const double* o;
  return run_sync(c, SMC_OUT_HEADER + (int)K, &o);
}

int smc_linear_predictor_adjoint(const smc_matrix* x, const smc_matrix* v,
                                 double* xt_v, double* sum_v) {
  static const char* fn = "linear_predictor_adjoint";
  if (int rc = ensure_ctx()) return rc;
  if (!x || x->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: x must be an f64 device matrix", fn);
  const int64_t N = x->rows, K = x->cols;
  if (!v || v->dtype != SMC_F64 || v->rows * v->cols != N || !is_vec(v))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: v must hold rows(x) doubles", fn);
  if (sum_v) *sum_v = 0.0;
  if (xt_v) memset(xt_v, 0, sizeof(double) * K);
  if (N == 0 || v->zero_pending) return SMC_OK;  // x^T 0
  std::vector<double> zeros((size_t)K, 0.0);
  GlmCall c;
  c.family = kLinear;
  c.x = x;
  c.aux_vec = v;
  c.beta_host = zeros.data();
  c.flags = SMC_VAR_BETA;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER + (int)K, &o)) return rc;
  if (sum_v) *sum_v = o[SMC_OUT_SUM_D];
  if (xt_v) memcpy(xt_v, o + SMC_OUT_HEADER, sizeof(double) * K);
  return SMC_OK;
}

int smc_vector_sum(const smc_matrix* v, double* sum) {
  static const char* fn = "vector_sum";
  if (int rc = ensure_ctx()) return rc;
  if (!v || v->dtype != SMC_F64 || !is_vec(v) || !sum)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: need an f64 device vector", fn);
  smc_matrix x0 = empty_design(v->rows * v->cols);
  return smc_linear_predictor_adjoint(&x0, v, nullptr, sum);
}

int smc_bernoulli_logit_lpmf(const smc_matrix* n, int n_scalar,
                             const smc_matrix* theta, unsigned flags, double* logp,
                             smc_matrix* d_theta) {
  static const char* fn = "bernoulli_logit_lpmf";
  if (int rc = ensure_ctx()) return rc;
  int64_t N;
  if (int rc = theta_shapes(fn, theta, n, nullptr, d_theta, nullptr, &N)) return rc;
  if (!logp) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL logp", fn);
  *logp = 0.0;
  if (N == 0) return SMC_OK;                                          // L52-54
  if (int rc = int_bounds(fn, n, n_scalar, 0, 1, true)) return rc;    // L47
  if ((flags & SMC_PROPTO) && !(flags & SMC_VAR_ALPHA)) return SMC_OK;  // L55-57
  smc_matrix x0 = empty_design(N);
  GlmCall c;
  c.family = kBernoulli;
  c.unfused = true;
  c.x = &x0;
  c.y = n;
  c.y_scalar = n_scalar;
  c.alpha_vec = theta;
  c.flags = flags & (SMC_PROPTO | SMC_VAR_ALPHA);
  c.d_alpha_vec = (flags & SMC_VAR_ALPHA) ? d_theta : nullptr;
  if (c.d_alpha_vec) d_theta->version++;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER, &o)) return rc;
  const double lp = o[SMC_OUT_LOGP];
  if (o[SMC_OUT_NONFINITE] > 0) {
    // check_not_nan(theta), L48-49: +-inf is a legal logit, NaN is not
    int kinds = 0;
    if (int rc = classify(theta, nullptr, 0, &kinds)) return rc;
    if (kinds & kHasNan)
      return fail(SMC_ERR_DOMAIN, "%s: Logit transformed probability parameter is nan",
                  fn);
  }
  *logp = lp;
  return SMC_OK;
}

int smc_poisson_log_lpmf(const smc_matrix* n, int n_scalar, const smc_matrix* alpha,
                         unsigned flags, double* logp, smc_matrix* d_alpha) {
  static const char* fn = "poisson_log_lpmf";
  if (int rc = ensure_ctx()) return rc;
  int64_t N;
  if (int rc = theta_shapes(fn, alpha, n, nullptr, d_alpha, nullptr, &N)) return rc;
  if (!logp) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL logp", fn);
  *logp = 0.0;
  if (int rc = int_bounds(fn, n, n_scalar, 0, 0, false)) return rc;  // L46
  if (N == 0) return SMC_OK;  // (an empty alpha has no NaN to report) L49-51
  const bool skip = (flags & SMC_PROPTO) && !(flags & SMC_VAR_ALPHA);  // L52-54
  smc_matrix x0 = empty_design(N);
  GlmCall c;
  c.family = kPoisson;
  c.unfused = true;
  c.x = &x0;
  c.y = n;
  c.y_scalar = n_scalar;
  c.alpha_vec = alpha;
  c.flags = flags & (SMC_PROPTO | SMC_VAR_ALPHA);
  c.d_alpha_vec = (flags & SMC_VAR_ALPHA) ? d_alpha : nullptr;
  if (c.d_alpha_vec) d_alpha->version++;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER, &o)) return rc;
  const double lp = o[SMC_OUT_LOGP];
  if (o[SMC_OUT_NONFINITE] > 0) {
    // check_not_nan(alpha) L47; then the two log(0) exits, L56-66
    int kinds = 0;
    if (int rc = classify(alpha, n, n_scalar, &kinds)) return rc;
    if (kinds & kHasNan) return fail(SMC_ERR_DOMAIN, "%s: Log rate parameter is nan", fn);
    const bool zero = kinds & (kHasPosInf | kHasNegInfNonzeroCount);
    if (zero && !skip) {
      // LOG_ZERO is returned as a constant: no partials
      if (c.d_alpha_vec) {
        if (int rc = smc_matrix_zero(d_alpha)) return rc;
      }
      *logp = -std::numeric_limits<double>::infinity();
      return SMC_OK;
    }
  }
  if (skip) {
    if (c.d_alpha_vec) return smc_matrix_zero(d_alpha);
    return SMC_OK;
  }
  *logp = lp;
  return SMC_OK;
}

int smc_neg_binomial_2_log_lpmf(const smc_matrix* n, int n_scalar,
                                const smc_matrix* eta, const smc_matrix* phi_vec,
                                double phi, unsigned flags, double* logp,
                                smc_matrix* d_eta, double* d_phi,
                                smc_matrix* d_phi_vec) {
  static const char* fn = "neg_binomial_2_log_lpmf";
  if (int rc = ensure_ctx()) return rc;
  int64_t N;
  if (int rc = theta_shapes(fn, eta, n, phi_vec, d_eta, d_phi_vec, &N)) return rc;
  if (!logp) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL logp", fn);
  *logp = 0.0;
  if (d_phi) *d_phi = 0.0;
  if (int rc = int_bounds(fn, n, n_scalar, 0, 0, false)) return rc;  // L45
  if (!phi_vec && (!(phi > 0.0) || !std::isfinite(phi)))            // L47
    return fail(SMC_ERR_DOMAIN,
                "%s: Precision parameter is %g, but must be positive finite", fn, phi);
  if (N == 0) return SMC_OK;  // L49-51
  const bool skip = (flags & SMC_PROPTO) && !(flags & (SMC_VAR_ALPHA | SMC_VAR_AUX));
  smc_matrix x0 = empty_design(N);
  GlmCall c;
  c.family = kNegBinomial;
  c.unfused = true;
  c.x = &x0;
  c.y = n;
  c.y_scalar = n_scalar;
  c.alpha_vec = eta;
  c.aux_vec = phi_vec;
  c.aux = phi;
  c.flags = flags & (SMC_PROPTO | SMC_VAR_ALPHA | SMC_VAR_AUX);
  c.d_alpha_vec = (flags & SMC_VAR_ALPHA) ? d_eta : nullptr;
  c.d_aux_vec = (flags & SMC_VAR_AUX) ? d_phi_vec : nullptr;
  if (c.d_alpha_vec) d_eta->version++;
  if (c.d_aux_vec) d_phi_vec->version++;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER, &o)) return rc;
  if (o[SMC_OUT_NONFINITE] > 0)  // check_finite(eta) L46, positive finite phi L47
    return fail(SMC_ERR_DOMAIN,
                "%s: Log location parameter or precision parameter is not finite", fn);
  if (skip) return SMC_OK;  // L52-54
  *logp = o[SMC_OUT_LOGP];
  if (d_phi && (flags & SMC_VAR_AUX) && !phi_vec) *d_phi = o[SMC_OUT_AUX];
  return SMC_OK;
}

int smc_normal_lpdf(const smc_matrix* y, double y_scalar, const smc_matrix* mu,
                    double mu_scalar, double sigma, unsigned flags, double* logp,
                    smc_matrix* d_y_vec, double* d_y, smc_matrix* d_mu_vec,
                    double* d_mu, double* d_sigma) {
  static const char* fn = "normal_lpdf";
  if (int rc = ensure_ctx()) return rc;
  if (!y && !mu)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: y or mu must be a device vector", fn);
  const smc_matrix* ref = mu ? mu : y;
  if (ref->dtype != SMC_F64 || !is_vec(ref))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: operands must be f64 device vectors", fn);
  const int64_t N = ref->rows * ref->cols;
  const smc_matrix* vs[4] = {y, mu, d_y_vec, d_mu_vec};
  for (const smc_matrix* v : vs)
    if (v && (v->dtype != SMC_F64 || v->rows * v->cols != N || !is_vec(v)))
      return fail(SMC_ERR_INVALID_ARGUMENT,
                  "%s: sizes of the random variable and the location parameter do not "
                  "match",
                  fn);
  if (!logp) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL logp", fn);
  *logp = 0.0;
  if (d_y) *d_y = 0.0;
  if (d_mu) *d_mu = 0.0;
  if (d_sigma) *d_sigma = 0.0;
  // scalars are checked at once, the vectors lazily below (L60-62)
  if (!y && std::isnan(y_scalar))
    return fail(SMC_ERR_DOMAIN, "%s: Random variable is nan", fn);
  if (!mu && !std::isfinite(mu_scalar))
    return fail(SMC_ERR_DOMAIN, "%s: Location parameter is not finite", fn);
  if (!(sigma > 0.0))
    return fail(SMC_ERR_DOMAIN, "%s: Scale parameter is %g, but must be positive", fn,
                sigma);
  if (N == 0) return SMC_OK;  // L64-66
  const bool skip
      = (flags & SMC_PROPTO) && !(flags & (SMC_VAR_Y | SMC_VAR_ALPHA | SMC_VAR_AUX));
  smc_matrix x0 = empty_design(N);
  GlmCall c;
  c.family = kNormal;
  c.unfused = true;
  c.x = &x0;
  c.y = y;
  c.y_scalar = y_scalar;
  c.alpha_vec = mu;
  c.alpha = mu_scalar;
  c.aux = sigma;
  c.flags = flags & (SMC_PROPTO | SMC_VAR_Y | SMC_VAR_ALPHA | SMC_VAR_AUX);
  c.d_alpha_vec = (flags & SMC_VAR_ALPHA) ? d_mu_vec : nullptr;
  c.d_y_vec = (flags & SMC_VAR_Y) ? d_y_vec : nullptr;
  if (c.d_alpha_vec) d_mu_vec->version++;
  if (c.d_y_vec) d_y_vec->version++;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER, &o)) return rc;
  const double lp = o[SMC_OUT_LOGP], sum_d = o[SMC_OUT_SUM_D], sum_r2 = o[SMC_OUT_AUX2];
  if (!std::isfinite(sum_r2) && std::isfinite(sigma)) {
    // some (y_i - mu_i) / sigma is not finite: check_not_nan(y), check_finite(mu)
    int ok = 1;
    if (mu) {
      if (int rc = smc_matrix_all_finite(mu, &ok)) return rc;
      if (!ok) return fail(SMC_ERR_DOMAIN, "%s: Location parameter is not finite", fn);
    }
    if (y) {
      int kinds = 0;
      if (int rc = classify(y, nullptr, 0, &kinds)) return rc;
      if (kinds & kHasNan) return fail(SMC_ERR_DOMAIN, "%s: Random variable is nan", fn);
    }
  }
  if (skip) return SMC_OK;  // L67-69
  *logp = lp;
  if (d_mu && (flags & SMC_VAR_ALPHA) && !mu) *d_mu = sum_d;
  if (d_y && (flags & SMC_VAR_Y) && !y) *d_y = -sum_d;
  if (d_sigma && (flags & SMC_VAR_AUX)) *d_sigma = (sum_r2 - (double)N) * (1.0 / sigma);
  return SMC_OK;
}

int smc_ordered_logistic_lpmf(const smc_matrix* y, int y_scalar,
                              const smc_matrix* lambda, const double* cuts,
                              int64_t ncuts, unsigned flags, double* logp,
                              smc_matrix* d_lambda, double* d_cuts) {
  static const char* fn = "ordered_logistic_lpmf";
  if (int rc = ensure_ctx()) return rc;
  int64_t N;
  if (int rc = theta_shapes(fn, lambda, y, nullptr, d_lambda, nullptr, &N)) return rc;
  if (!logp || ncuts < 0 || (ncuts > 0 && !cuts))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL cuts or logp", fn);
  *logp = 0.0;
  if (d_cuts && ncuts) memset(d_cuts, 0, sizeof(double) * ncuts);
  if (N > 0) {  // check_finite(lambda), L105: before the size_zero exit
    int ok = 1;
    if (int rc = smc_matrix_all_finite(lambda, &ok)) return rc;
    if (!ok) return fail(SMC_ERR_DOMAIN, "%s: Location parameter is not finite", fn);
  }
  if (N == 0) return SMC_OK;  // L106-108 (one cut vector is always present here)
  const int64_t C = ncuts + 1;
  if (int rc = int_bounds(fn, y, y_scalar, 1, (int)C, true)) return rc;  // L110
  for (int64_t i = 1; i < ncuts; ++i)  // check_ordered, L115
    if (!(cuts[i] > cuts[i - 1]))
      return fail(SMC_ERR_DOMAIN, "%s: Cut-points are not a valid ordered vector", fn);
  if (ncuts == 1 && std::isnan(cuts[0]))
    return fail(SMC_ERR_DOMAIN, "%s: Cut-points are not a valid ordered vector", fn);
  if (C > 1) {  // L116-121
    if (C > 2 && !std::isfinite(cuts[C - 2]))
      return fail(SMC_ERR_DOMAIN, "%s: Final cut-point is not finite", fn);
    if (!std::isfinite(cuts[0]))
      return fail(SMC_ERR_DOMAIN, "%s: First cut-point is not finite", fn);
  }
  if ((flags & SMC_PROPTO) && !(flags & (SMC_VAR_ALPHA | SMC_VAR_AUX)))
    return SMC_OK;  // L123-125
  if (ncuts == 0) {
    // a single class: every term is log(1) = 0 and every partial is 0
    if ((flags & SMC_VAR_ALPHA) && d_lambda) return smc_matrix_zero(d_lambda);
    return SMC_OK;
  }
  smc_matrix x0 = empty_design(N);
  GlmCall c;
  c.family = kOrdered;
  c.unfused = true;
  c.x = &x0;
  c.y = y;
  c.y_scalar = y_scalar;
  c.alpha_vec = lambda;
  c.cuts_host = cuts;
  c.ncuts = ncuts;
  c.flags = flags & (SMC_PROPTO | SMC_VAR_ALPHA | SMC_VAR_AUX);
  c.d_alpha_vec = (flags & SMC_VAR_ALPHA) ? d_lambda : nullptr;
  if (c.d_alpha_vec) d_lambda->version++;
  const double* o;
  if (int rc = run_sync(c, SMC_OUT_HEADER + (int)ncuts, &o)) return rc;
  *logp = o[SMC_OUT_LOGP];
  if (d_cuts && (flags & SMC_VAR_AUX))
    memcpy(d_cuts, o + SMC_OUT_HEADER, sizeof(double) * ncuts);
  return SMC_OK;
}

}  // extern "C"

namespace {

// ordered_logistic_lpmf with ONE CUT-POINT VECTOR PER OUTCOME (prim/prob/
// ordered_logistic_lpmf.hpp L72-200 with a std::vector of cut vectors,
// opencl/prim/ordered_logistic_lpmf.hpp L68-160 with an (C-1) x N matrix_cl): row i
// reads column i of `cuts`.  Lane = outcome.  The row's two cut points, the value
// term and d1 / d2 follow the GLM link (glm_link.cuh, kOrdered) expression by
// expression, with the class constants evaluated per row; the (C-1) x N partial of the
// cut points is written whole (two non-zero entries per column, L189-198), so it
// needs no zero fill.  The column checks of L112-122 (ordered, first and last cut
// finite) ride along: bit 0 not ordered, bit 1 last cut not finite, bit 2 first.
constexpr int kOrdThreads = 256;
constexpr int kCutsUnordered = 1, kCutsLastInf = 2, kCutsFirstInf = 4;

__global__ void __launch_bounds__(kOrdThreads)
    ordered_rows_kernel(int64_t N, int ncuts, const int* __restrict__ y, int y_scalar,
                        const double* __restrict__ lambda,
                        const double* __restrict__ cuts, int64_t ld_cuts,
                        double* __restrict__ d_lambda, double* __restrict__ d_cuts,
                        int64_t ld_dcuts, int check_only,
                        double* __restrict__ block_partials, int* __restrict__ flags_out) {
  __shared__ double sh[kOrdThreads / 32];
  double lp = 0.0;
  int bad = 0;
  const int C = ncuts + 1;
  for (int64_t row = blockIdx.x * (int64_t)kOrdThreads + threadIdx.x; row < N;
       row += (int64_t)gridDim.x * kOrdThreads) {
    const double* col = cuts + row * ld_cuts;
    double prev = col[0];
    if (!isfinite(prev)) bad |= kCutsFirstInf;
    if (prev != prev) bad |= kCutsUnordered;  // check_ordered rejects a NaN
    for (int j = 1; j < ncuts; ++j) {
      const double cur = col[j];
      if (!(cur > prev)) bad |= kCutsUnordered;
      prev = cur;
    }
    if (C > 2 && !isfinite(prev)) bad |= kCutsLastInf;
    if (check_only) continue;
    const int c = y ? y[row] : y_scalar;
    double ce[4];
    ordered_class_entry(col, ncuts, c, ce);
    const double loc = lambda[row];
    const double cut2 = loc - ce[1], cut1 = loc - ce[0];
    const double e1 = exp_nonpos(-fabs(cut1)), e2 = exp_nonpos(-fabs(cut2));
    const double d1 = div_or_zero(cut2 > 0.0 ? e2 : 1.0, 1.0 + e2) - ce[2];
    const double d2 = ce[3] - div_or_zero(cut1 > 0.0 ? e1 : 1.0, 1.0 + e1);
    if (d_lambda) d_lambda[row] = d1 - d2;
    if (d_cuts) {
      double* dc = d_cuts + row * ld_dcuts;
      for (int j = 0; j < ncuts; ++j) dc[j] = j == c - 1 ? d2 : (j == c - 2 ? -d1 : 0.0);
    }
    const double A = (cut1 > 0.0 ? -cut1 : 0.0) - log1p_or_zero(e1);
    const double B = (cut2 <= 0.0 ? cut2 : 0.0) - log1p_or_zero(e2);
    if (c == 1)
      lp += A;
    else if (c == C)
      lp += B;
    else
      lp += B + log1m_exp(cut1 - cut2) + A;
  }
  // fixed-order block sum: butterfly inside the warp, warp totals in index order
#pragma unroll
  for (int o = 16; o; o >>= 1) lp += __shfl_xor_sync(0xffffffffu, lp, o);
  bad = __reduce_or_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0) {
    sh[threadIdx.x >> 5] = lp;
    if (bad) atomicOr(flags_out, bad);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kOrdThreads / 32; ++w) t += sh[w];
    block_partials[blockIdx.x] = t;
  }
}

__global__ void ordered_rows_final_kernel(const double* __restrict__ block_partials,
                                          int nblocks, double* __restrict__ out) {
  double v = 0.0;
  for (int b = 0; b < nblocks; ++b) v += block_partials[b];
  *out = v;
}

}  // namespace

extern "C" {

int smc_ordered_logistic_lpmf_rows(const smc_matrix* y, int y_scalar,
                                   const smc_matrix* lambda, const smc_matrix* cuts,
                                   unsigned flags, double* logp, smc_matrix* d_lambda,
                                   smc_matrix* d_cuts) {
  static const char* fn = "ordered_logistic_lpmf";
  if (int rc = ensure_ctx()) return rc;
  int64_t N;
  if (int rc = theta_shapes(fn, lambda, y, nullptr, d_lambda, nullptr, &N)) return rc;
  if (!logp || !cuts || cuts->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: cuts must be an f64 device matrix", fn);
  if (is_sharded(cuts) || is_sharded(d_cuts))
    return fail(SMC_ERR_UNSUPPORTED, "%s: the cut points are not sharded", fn);
  const int64_t ncuts = cuts->rows, M = cuts->cols;
  if (M > 1 && ncuts > 0)  // (the one-vector form below takes a sharded lambda)
    for (const smc_matrix* m : {y, lambda, (const smc_matrix*)d_lambda})
      if (is_sharded(m))
        return fail(SMC_ERR_UNSUPPORTED,
                    "%s: per-outcome cut points with a sharded location vector", fn);
  const bool var_cuts = flags & SMC_VAR_AUX;
  if (var_cuts
      && (!d_cuts || d_cuts->dtype != SMC_F64 || d_cuts->rows != ncuts || d_cuts->cols != M))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: d_cuts must have the shape of cuts", fn);
  *logp = 0.0;
  if (M > 1 && M != N)  // L102-105
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "%s: Length of location variables (%lld) does not match the number of "
                "cutpoint vectors (%lld)",
                fn, (long long)N, (long long)M);
  if (M <= 1 || ncuts == 0 || N == 0) {
    // one cut-point vector for every outcome (or nothing to index): the host-cuts
    // entry point, C - 1 doubles each way
    std::vector<double> c_host((size_t)(ncuts > 0 ? ncuts : 1)), dc_host(c_host.size());
    if (M == 1 && ncuts > 0)
      if (int rc = smc_matrix_download(cuts, c_host.data(), ncuts)) return rc;
    if (M == 0) {  // no cut-point vector at all: L107-109 after the finiteness check
      if (N > 0) {
        int ok = 1;
        if (int rc = smc_matrix_all_finite(lambda, &ok)) return rc;
        if (!ok) return fail(SMC_ERR_DOMAIN, "%s: Location parameter is not finite", fn);
      }
      return SMC_OK;
    }
    if (int rc = smc_ordered_logistic_lpmf(y, y_scalar, lambda, c_host.data(), ncuts, flags,
                                           logp, d_lambda, dc_host.data()))
      return rc;
    if (var_cuts && ncuts > 0 && M > 0) {
      if (M == 1) {
        if (int rc = smc_matrix_upload(d_cuts, dc_host.data(), ncuts)) return rc;
      } else {
        if (int rc = smc_matrix_zero(d_cuts)) return rc;
      }
    }
    return SMC_OK;
  }
  {  // check_finite(lambda), L106
    int ok = 1;
    if (int rc = smc_matrix_all_finite(lambda, &ok)) return rc;
    if (!ok) return fail(SMC_ERR_DOMAIN, "%s: Location parameter is not finite", fn);
  }
  const int64_t C = ncuts + 1;
  if (int rc = int_bounds(fn, y, y_scalar, 1, (int)C, true)) return rc;  // L111
  for (const smc_matrix* m : {y, lambda, cuts})
    if (int rc = realize(m)) return rc;
  Context& cx = ctx();
  const bool skip = (flags & SMC_PROPTO) && !(flags & (SMC_VAR_ALPHA | SMC_VAR_AUX));
  int grid = (int)((N + kOrdThreads - 1) / kOrdThreads);
  if (grid > cx.sm_count * 8) grid = cx.sm_count * 8;
  if (int rc = ensure_scratch(sizeof(double) * (size_t)grid)) return rc;
  if (int rc = ensure_out(4096)) return rc;
  int* fl = reinterpret_cast<int*>(cx.out_host + 1);
  *fl = 0;
  cx.out_host[0] = 0.0;
  double* dl = (flags & SMC_VAR_ALPHA) && d_lambda ? static_cast<double*>(d_lambda->data)
                                                   : nullptr;
  double* dc = var_cuts ? static_cast<double*>(d_cuts->data) : nullptr;
  if (dl) {
    d_lambda->zero_pending = false;
    d_lambda->version++;
  }
  if (dc) {
    d_cuts->zero_pending = false;
    d_cuts->version++;
  }
  ordered_rows_kernel<<<grid, kOrdThreads, 0, cx.stream>>>(
      N, (int)ncuts, y ? static_cast<const int*>(y->data) : nullptr, y_scalar,
      static_cast<const double*>(lambda->data), static_cast<const double*>(cuts->data),
      cuts->ld, dl, dc, dc ? d_cuts->ld : 0, skip ? 1 : 0, cx.scratch, fl);
  SMC_CUDA(cudaGetLastError());
  ordered_rows_final_kernel<<<1, 1, 0, cx.stream>>>(cx.scratch, grid, cx.out_host);
  SMC_CUDA(cudaGetLastError());
  cx.launches += 2;
  SMC_CUDA(cudaStreamSynchronize(cx.stream));
  const int bad = *fl;
  if (bad & kCutsUnordered)  // L115
    return fail(SMC_ERR_DOMAIN, "%s: Cut-points are not a valid ordered vector", fn);
  if (bad & kCutsLastInf) return fail(SMC_ERR_DOMAIN, "%s: Final cut-point is not finite", fn);
  if (bad & kCutsFirstInf) return fail(SMC_ERR_DOMAIN, "%s: First cut-point is not finite", fn);
  if (!skip) *logp = cx.out_host[0];
  return SMC_OK;
}

}  // extern "C"
