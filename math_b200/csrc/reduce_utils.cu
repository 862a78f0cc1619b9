// Small deterministic reductions over the integer response vector y.  y is
// DATA, so both results are cached on the matrix handle and only recomputed
// after an upload: min/max for the reference's eager range checks
// (check_bounded / check_nonnegative) and sum_i lgamma(y_i + 1), the
// parameter-free term of the poisson / neg-binomial log density
// (poisson_log_glm_lpmf.hpp L126-128, neg_binomial_2_log_glm_lpmf.hpp L163-169).
#include <climits>

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
  if (y->range_valid && y->lgamma_valid) return SMC_OK;
  if (int rc = ensure_ctx()) return rc;
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

int y_range(const smc_matrix* y, int* lo, int* hi) {
  if (y->rows * y->cols == 0) {
    *lo = INT_MAX;
    *hi = INT_MIN;
    return SMC_OK;
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
