// Small deterministic reductions over the integer response vector y.  y is
// DATA, so both results are cached on the matrix handle and only recomputed
// after an upload: min/max for the reference's eager range checks
// (check_bounded / check_nonnegative) and sum_i lgamma(y_i + 1), the
// parameter-free term of the poisson / neg-binomial log density
// (poisson_log_glm_lpmf.hpp L126-128, neg_binomial_2_log_glm_lpmf.hpp L163-169).
#include <climits>
#include <cmath>

#include "smc_internal.h"

namespace smc {

constexpr int kRedThreads = 256;

__global__ void y_stats_kernel(const int* __restrict__ y, int64_t n,
                               int* __restrict__ lo_out, int* __restrict__ hi_out,
                               double* __restrict__ lg_out) {
  __shared__ int s_lo[kRedThreads / 32], s_hi[kRedThreads / 32];
  __shared__ double s_lg[kRedThreads / 32];
  int lo = INT_MAX, hi = INT_MIN;
  double lg = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int v = y[i];
    lo = min(lo, v);
    hi = max(hi, v);
    lg += lgamma((double)v + 1.0);
  }
  for (int o = 16; o; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    lg += __shfl_xor_sync(0xffffffffu, lg, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    s_lo[w] = lo;
    s_hi[w] = hi;
    s_lg[w] = lg;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 1; j < kRedThreads / 32; ++j) {
      lo = min(lo, s_lo[j]);
      hi = max(hi, s_hi[j]);
      lg += s_lg[j];
    }
    lo_out[blockIdx.x] = lo;
    hi_out[blockIdx.x] = hi;
    lg_out[blockIdx.x] = lg;
  }
}

static int compute_stats(const smc_matrix* yc) {
  smc_matrix* y = const_cast<smc_matrix*>(yc);
  std::lock_guard<std::mutex> lock(cache_mutex());
  if (y->range_valid && y->lgamma_valid) return SMC_OK;
  if (int rc = ensure_ctx()) return rc;
  if (int rc = realize(y)) return rc;
  Context& c = ctx();
  const int64_t n = y->rows * y->cols;
  if (y->cols > 1 && y->ld != y->rows)
    return fail(SMC_ERR_UNSUPPORTED, "y statistics: vectors only");
  int grid = (int)((n + kRedThreads - 1) / kRedThreads);
  if (grid > c.sm_count * 8) grid = c.sm_count * 8;
  if (grid < 1) grid = 1;
  const size_t bytes = (size_t)grid * 16;
  if (int rc = ensure_scratch(bytes)) return rc;
  if (int rc = ensure_out(bytes)) return rc;
  int* lo_d = reinterpret_cast<int*>(c.scratch);
  int* hi_d = lo_d + grid;
  double* lg_d = reinterpret_cast<double*>(c.scratch) + grid;
  y_stats_kernel<<<grid, kRedThreads, 0, c.stream>>>(
      static_cast<const int*>(y->data), n, lo_d, hi_d, lg_d);
  SMC_CUDA(cudaGetLastError());
  SMC_CUDA(cudaMemcpyAsync(c.out_host, c.scratch, bytes, cudaMemcpyDeviceToHost,
                           c.stream));
  SMC_CUDA(cudaStreamSynchronize(c.stream));
  const int* lo_h = reinterpret_cast<const int*>(c.out_host);
  const int* hi_h = lo_h + grid;
  const double* lg_h = c.out_host + grid;
  int lo = INT_MAX, hi = INT_MIN;
  double lg = 0.0;
  for (int b = 0; b < grid; ++b) {
    lo = lo_h[b] < lo ? lo_h[b] : lo;
    hi = hi_h[b] > hi ? hi_h[b] : hi;
    lg += lg_h[b];
  }
  y->imin = lo;
  y->imax = hi;
  y->lgamma_sum = lg;
  y->range_valid = true;
  y->lgamma_valid = true;
  return SMC_OK;
}

// ------------------------------------------------- binomial pair statistics
// Device restatement of binomial_coefficient_log(N, n) for integer 0 <= n <= N
// (prim/fun/binomial_coefficient_log.hpp L79-115) and what it calls: lbeta
// (lbeta.hpp L64-118), lgamma_stirling_diff (lgamma_stirling_diff.hpp L44-76),
// lgamma_stirling (lgamma_stirling.hpp L27-29).  Same branches, same constants.
constexpr double kHalfLogTwoPi = 0.91893853320467274178032973640561764;
constexpr double kStirlingUseful = 10.0;

__host__ __device__ inline double lgamma_stirling_diff_dev(double x) {
  if (x == 0.0) return (double)INFINITY;
  if (x < kStirlingUseful) return lgamma(x) - (kHalfLogTwoPi + (x - 0.5) * log(x) - x);
  const double series[6]
      = {0.0833333333333333333333333,   -0.00277777777777777777777778,
         0.000793650793650793650793651, -0.000595238095238095238095238,
         0.000841750841750841750841751, -0.00191752691752691752691753};
  double result = 0.0;
  double multiplier = 1.0 / x;
  const double inv_x_squared = multiplier * multiplier;
#pragma unroll
  for (int n = 0; n < 6; ++n) {
    if (n > 0) multiplier *= inv_x_squared;
    result += series[n] * multiplier;
  }
  return result;
}

__host__ __device__ inline double lbeta_dev(double a, double b) {
  const double x = a < b ? a : b, y = a < b ? b : a;  // x is the smaller
  if (x == 0.0) return (double)INFINITY;
  if (y < kStirlingUseful) return lgamma(x) + lgamma(y) - lgamma(x + y);
  const double x_over_xy = x / (x + y);
  if (x < kStirlingUseful) {
    const double sd = lgamma_stirling_diff_dev(y) - lgamma_stirling_diff_dev(x + y);
    const double st = (y - 0.5) * log1p(-x_over_xy) + x * (1.0 - log(x + y));
    return st + lgamma(x) + sd;
  }
  const double sd = lgamma_stirling_diff_dev(x) + lgamma_stirling_diff_dev(y)
                    - lgamma_stirling_diff_dev(x + y);
  const double st = (x - 0.5) * log(x_over_xy) + y * log1p(-x_over_xy) + kHalfLogTwoPi
                    - 0.5 * log(y);
  return st + sd;
}

__host__ __device__ inline double binomial_coefficient_log_dev(double n, double k) {
  if (k > n / 2.0 + 1e-8) k = n - k;  // the more stable symmetric branch, L89-91
  const double n_plus_1 = n + 1.0;
  const double n_plus_1_mk = n_plus_1 - k;
  if (k == 0.0) return 0.0;
  if (n_plus_1 < kStirlingUseful)
    return lgamma(n_plus_1) - lgamma(k + 1.0) - lgamma(n_plus_1_mk);
  return -lbeta_dev(n_plus_1_mk, k + 1.0) - log1p(n);
}

__global__ void binom_stats_kernel(const int* __restrict__ n, int n_scalar,
                                   const int* __restrict__ trials, int trials_scalar,
                                   int64_t count, int* __restrict__ bad_out,
                                   double* __restrict__ sum_out) {
  __shared__ int s_bad[kRedThreads / 32];
  __shared__ double s_sum[kRedThreads / 32];
  int bad = 0;
  double sum = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int ni = n ? n[i] : n_scalar;
    const int Ni = trials ? trials[i] : trials_scalar;
    if (ni < 0 || ni > Ni || Ni < 0)  // check_bounded / check_nonnegative, L98-99
      bad = 1;
    else
      sum += binomial_coefficient_log_dev((double)Ni, (double)ni);
  }
  for (int o = 16; o; o >>= 1) {
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    s_bad[w] = bad;
    s_sum[w] = sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 1; j < kRedThreads / 32; ++j) {
      bad |= s_bad[j];
      sum += s_sum[j];
    }
    bad_out[blockIdx.x] = bad;
    sum_out[blockIdx.x] = sum;
  }
}

int binom_stats(const smc_matrix* n, int n_scalar, const smc_matrix* trials,
                int trials_scalar, int64_t count, bool* in_support,
                double* coef_sum) {
  *in_support = true;
  *coef_sum = 0.0;
  if (count <= 0) return SMC_OK;
  if (is_sharded(n) || is_sharded(trials)) {
    const smc_matrix* ref = is_sharded(n) ? n : trials;
    if ((n && !is_sharded(n)) || (trials && !is_sharded(trials))
        || (n && trials && !same_partition(n, trials)))
      return fail(SMC_ERR_INVALID_ARGUMENT,
                  "binomial: successes and trials are not sharded alike");
    return for_each_shard(ref, [&](int g, smc_matrix* p, int64_t) {
      bool ok = true;
      double s = 0.0;
      if (int rc = binom_stats(n ? n->shards[g] : nullptr, n_scalar,
                               trials ? trials->shards[g] : nullptr, trials_scalar, p->rows,
                               &ok, &s))
        return rc;
      *in_support = *in_support && ok;
      *coef_sum += s;
      return (int)SMC_OK;
    });
  }
  // the cache lives on whichever operand is a matrix (n first)
  smc_matrix* owner = const_cast<smc_matrix*>(n ? n : trials);
  if (!owner) {  // two broadcast scalars: one pair, evaluated on the host
    *in_support = !(n_scalar < 0 || n_scalar > trials_scalar || trials_scalar < 0);
    if (*in_support)
      *coef_sum = binomial_coefficient_log_dev((double)trials_scalar, (double)n_scalar);
    return SMC_OK;
  }
  const smc_matrix* partner = n ? trials : nullptr;
  const int partner_scalar = n ? trials_scalar : n_scalar;
  std::lock_guard<std::mutex> lock(cache_mutex());
  if (owner && owner->binom_valid && owner->binom_self_version == owner->version
      && owner->binom_partner_id == (partner ? partner->id : 0)
      && (partner ? owner->binom_partner_version == partner->version
                  : owner->binom_partner_scalar == partner_scalar)) {
    *in_support = owner->binom_in_support;
    *coef_sum = owner->binom_coef_sum;
    return SMC_OK;
  }
  if (int rc = ensure_ctx()) return rc;
  if (int rc = realize(n)) return rc;
  if (int rc = realize(trials)) return rc;
  Context& c = ctx();
  int grid = (int)((count + kRedThreads - 1) / kRedThreads);
  if (grid > c.sm_count * 8) grid = c.sm_count * 8;
  const size_t bytes = (size_t)grid * 16;
  if (int rc = ensure_scratch(bytes)) return rc;
  if (int rc = ensure_out(bytes)) return rc;
  double* sum_d = c.scratch;
  int* bad_d = reinterpret_cast<int*>(c.scratch + grid);
  binom_stats_kernel<<<grid, kRedThreads, 0, c.stream>>>(
      n ? static_cast<const int*>(n->data) : nullptr, n_scalar,
      trials ? static_cast<const int*>(trials->data) : nullptr, trials_scalar, count,
      bad_d, sum_d);
  SMC_CUDA(cudaGetLastError());
  SMC_CUDA(cudaMemcpyAsync(c.out_host, c.scratch, bytes, cudaMemcpyDeviceToHost,
                           c.stream));
  SMC_CUDA(cudaStreamSynchronize(c.stream));
  const int* bad_h = reinterpret_cast<const int*>(c.out_host + grid);
  bool ok = true;
  double sum = 0.0;
  for (int b = 0; b < grid; ++b) {
    ok = ok && bad_h[b] == 0;
    sum += c.out_host[b];
  }
  *in_support = ok;
  *coef_sum = sum;
  if (owner) {
    owner->binom_valid = true;
    owner->binom_self_version = owner->version;
    owner->binom_partner_id = partner ? partner->id : 0;
    owner->binom_partner_version = partner ? partner->version : 0;
    owner->binom_partner_scalar = partner_scalar;
    owner->binom_in_support = ok;
    owner->binom_coef_sum = sum;
  }
  return SMC_OK;
}

int y_range(const smc_matrix* y, int* lo, int* hi) {
  if (y->rows * y->cols == 0) {
    *lo = INT_MAX;
    *hi = INT_MIN;
    return SMC_OK;
  }
  if (is_sharded(y)) {
    *lo = INT_MAX;
    *hi = INT_MIN;
    return for_each_shard(y, [&](int, smc_matrix* p, int64_t) {
      int l, h;
      if (int rc = y_range(p, &l, &h)) return rc;
      *lo = l < *lo ? l : *lo;
      *hi = h > *hi ? h : *hi;
      return (int)SMC_OK;
    });
  }
  if (int rc = compute_stats(y)) return rc;
  *lo = y->imin;
  *hi = y->imax;
  return SMC_OK;
}

int y_lgamma_sum(const smc_matrix* y, double* sum) {
  if (y->rows * y->cols == 0) {
    *sum = 0.0;
    return SMC_OK;
  }
  if (int rc = compute_stats(y)) return rc;
  *sum = y->lgamma_sum;
  return SMC_OK;
}

}  // namespace smc
