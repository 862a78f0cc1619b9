// Device-side multi-indexing of a small host vector by a resident index vector,
// and its reverse sweep: the hierarchical-intercept step in front of the GLMs
// (alpha_i = z[group_i]).  Without it the N-vector intercept of a
// hierarchical model crosses PCIe on every evaluation (N * 8 bytes up, N * 8 bytes
// of partials down); with it G doubles go up and G doubles come down.
//
// Stands where opencl/kernel_generator/indexing.hpp (`indexing(mat, idx)`) and
// opencl/indexing_rev.hpp L24-60 stand.  The reference's reverse kernels add
// with atomics; this one is deterministic: every warp owns a contiguous row
// range and a private accumulator per group in shared memory, collisions inside
// a 32-row step are resolved in lane order, warps / CTAs are combined in index
// order.
#include <cub/device/device_radix_sort.cuh>

#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "smc_internal.h"

using namespace smc;

namespace {

constexpr int kIdxThreads = 256;
constexpr int kIdxWarps = kIdxThreads / 32;
constexpr int64_t kMaxGroupsFast = 2048;  // 8 warps x 2048 doubles = 128 KB of accumulators

__global__ void indexing_kernel(const double* __restrict__ z, const int* __restrict__ idx,
                                int64_t n, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = z[idx[i]];
}

__global__ void __launch_bounds__(kIdxThreads)
    indexing_rev_kernel(const int* __restrict__ idx, const double* __restrict__ v,
                        int64_t n, int G, int64_t rows_per_warp,
                        double* __restrict__ partials) {
  extern __shared__ double idx_smem[];
  double* acc = idx_smem;                                  // [warps][G]
  double* stage_v = acc + (size_t)kIdxWarps * G;           // [warps][32]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = threadIdx.x; j < kIdxWarps * G; j += kIdxThreads) acc[j] = 0.0;
  __syncthreads();
  double* my_acc = acc + (size_t)warp * G;
  double* my_stage = stage_v + warp * 32;
  const int64_t w = (int64_t)blockIdx.x * kIdxWarps + warp;
  const int64_t lo = w * rows_per_warp;
  int64_t hi = lo + rows_per_warp;
  if (hi > n) hi = n;
  for (int64_t r = lo; r < hi; r += 32) {
    const int64_t i = r + lane;
    const bool live = i < hi;
    const int g = live ? idx[i] : -1;
    my_stage[lane] = live ? v[i] : 0.0;
    __syncwarp();
    const unsigned peers = __match_any_sync(0xffffffffu, g);
    if (live && lane == __ffs(peers) - 1) {
      // the lowest lane of each group adds its peers' values in lane order
      double s = 0.0;
      for (unsigned m = peers; m; m &= m - 1) s += my_stage[__ffs(m) - 1];
      my_acc[g] += s;
    }
    __syncwarp();
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += kIdxThreads) {
    double s = 0.0;
    for (int ww = 0; ww < kIdxWarps; ++ww) s += acc[(size_t)ww * G + g];  // fixed order
    partials[(size_t)blockIdx.x * G + g] = s;
  }
}

__global__ void indexing_rev_final_kernel(const double* __restrict__ partials, int nblocks,
                                          int G, double* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * G + g];  // fixed order
  out[g] = s;
}

// More groups than the per-warp accumulators hold: the rows of every group are
// summed from a cached, stably sorted row list (idx is data, so the list is built
// once per upload, on the device): one warp per group, lanes stride the group's rows, fixed-order
// combine -- deterministic, and nothing but G doubles crosses PCIe per evaluation.
__global__ void __launch_bounds__(kIdxThreads)
    indexing_rev_sorted_kernel(const int* __restrict__ perm, const int* __restrict__ off,
                               const double* __restrict__ v, int G, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int64_t g = (int64_t)blockIdx.x * kIdxWarps + (threadIdx.x >> 5); g < G;
       g += (int64_t)gridDim.x * kIdxWarps) {
    const int lo = off[g], hi = off[g + 1];
    double s = 0.0;
    for (int j = lo + lane; j < hi; j += 32) s += v[perm[j]];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[g] = s;
  }
}

__global__ void iota_kernel(int* __restrict__ v, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    v[i] = i;
}

// off[g] = first position of group g in the sorted keys (n for groups past the
// last one; an empty group gets the position of the next non-empty one).
__global__ void group_offsets_kernel(const int* __restrict__ keys, int n, int G,
                                     int* __restrict__ off) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int cur = keys[j], prev = j ? keys[j - 1] : -1;
    for (int g = prev + 1; g <= cur; ++g) off[g] = j;
    if (j == n - 1)
      for (int g = cur + 1; g <= G; ++g) off[g] = n;
  }
}

// Builds (or finds) the sorted row list of idx for G groups: a stable radix sort of
// (group, row) pairs on the device (cub), then the start of every group.  idx is
// data, so this runs once per upload.
int group_index(const char* fn, const smc_matrix* idxc, int64_t G) {
  smc_matrix* idx = const_cast<smc_matrix*>(idxc);
  std::lock_guard<std::mutex> lock(cache_mutex());
  if (idx->grp_perm && idx->grp_version == idx->version && idx->grp_G == G) return SMC_OK;
  const int64_t n64 = idx->rows * idx->cols;
  if (n64 > 0x7fffffff || G > 0x7ffffffe)
    return fail(SMC_ERR_UNSUPPORTED, "%s: more than 2^31 rows or groups", fn);
  const int n = (int)n64;
  Context& c = ctx();
  const size_t pb = sizeof(int) * (size_t)n, ob = sizeof(int) * ((size_t)G + 1);
  if (idx->grp_perm && idx->grp_perm_bytes != pb) {
    cache_free(idx->grp_perm, idx->grp_perm_bytes);
    idx->grp_perm = nullptr;
  }
  if (idx->grp_off && idx->grp_off_bytes != ob) {
    cache_free(idx->grp_off, idx->grp_off_bytes);
    idx->grp_off = nullptr;
  }
  if (!idx->grp_perm) {
    void* p = nullptr;
    if (int rc = cache_alloc(&p, pb)) return rc;
    idx->grp_perm = static_cast<int*>(p);
    idx->grp_perm_bytes = pb;
  }
  if (!idx->grp_off) {
    void* p = nullptr;
    if (int rc = cache_alloc(&p, ob)) return rc;
    idx->grp_off = static_cast<int*>(p);
    idx->grp_off_bytes = ob;
  }
  int end_bit = 1;
  while (end_bit < 31 && (1ll << end_bit) < G) ++end_bit;
  size_t temp_bytes = 0;
  const int* keys_in = static_cast<const int*>(idx->data);
  SMC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys_in, (int*)nullptr,
                                           (const int*)nullptr, idx->grp_perm, n, 0, end_bit,
                                           c.stream));
  // scratch: sorted keys, the row numbers 0..n-1, cub's workspace
  const size_t off_iota = (pb + 255) & ~(size_t)255, off_tmp = 2 * off_iota;
  if (int rc = ensure_scratch(off_tmp + temp_bytes)) return rc;
  char* scratch = reinterpret_cast<char*>(c.scratch);
  int* keys_sorted = reinterpret_cast<int*>(scratch);
  int* rows_in = reinterpret_cast<int*>(scratch + off_iota);
  int blocks = (n + 255) / 256;
  if (blocks > c.sm_count * 16) blocks = c.sm_count * 16;
  iota_kernel<<<blocks, 256, 0, c.stream>>>(rows_in, n);
  SMC_CUDA(cudaGetLastError());
  SMC_CUDA(cub::DeviceRadixSort::SortPairs(scratch + off_tmp, temp_bytes, keys_in, keys_sorted,
                                           rows_in, idx->grp_perm, n, 0, end_bit, c.stream));
  group_offsets_kernel<<<blocks, 256, 0, c.stream>>>(keys_sorted, n, (int)G, idx->grp_off);
  SMC_CUDA(cudaGetLastError());
  c.launches += 3;
  // other host threads (their own streams) may pick the list up as soon as it is
  // marked valid
  SMC_CUDA(cudaStreamSynchronize(c.stream));
  idx->grp_version = idx->version;
  idx->grp_G = G;
  return SMC_OK;
}

int check_index(const char* fn, const smc_matrix* idx, int64_t G) {
  if (!idx || idx->dtype != SMC_I32 || !vec_contiguous(idx))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: idx must be an i32 device vector", fn);
  if (idx->rows * idx->cols == 0) return SMC_OK;
  int lo, hi;
  if (int rc = y_range(idx, &lo, &hi)) return rc;  // idx is data: cached
  if (lo < 0 || hi >= G)
    return fail(SMC_ERR_DOMAIN, "%s: index out of range [%d, %d] for %lld elements", fn,
                lo, hi, (long long)G);
  return SMC_OK;
}

}  // namespace

extern "C" {

int smc_indexing(const double* z, int64_t G, const smc_matrix* idx, smc_matrix* out) {
  static const char* fn = "indexing";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = refuse_sharded(fn, {idx, out})) return rc;
  if (int rc = check_index(fn, idx, G)) return rc;
  const int64_t n = idx->rows * idx->cols;
  if (!out || out->dtype != SMC_F64 || out->rows * out->cols != n || !vec_contiguous(out))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: out must hold size(idx) doubles", fn);
  if (n == 0) return SMC_OK;
  out->zero_pending = false;  // overwritten
  if (!z || G <= 0) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: empty source vector", fn);
  Context& c = ctx();
  if (int rc = ensure_params(sizeof(double) * (size_t)G)) return rc;
  SMC_CUDA(cudaMemcpyAsync(c.params_dev, z, sizeof(double) * (size_t)G,
                           cudaMemcpyHostToDevice, c.stream));
  int grid = (int)((n + kIdxThreads - 1) / kIdxThreads);
  if (grid > c.sm_count * 16) grid = c.sm_count * 16;
  out->version++;
  indexing_kernel<<<grid, kIdxThreads, 0, c.stream>>>(
      c.params_dev, static_cast<const int*>(idx->data), n, static_cast<double*>(out->data));
  SMC_CUDA(cudaGetLastError());
  c.launches += 1;
  return SMC_OK;
}

int smc_indexing_rev(const smc_matrix* idx, const smc_matrix* res_adj, int64_t G,
                     double* adj_z) {
  static const char* fn = "indexing_rev";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = refuse_sharded(fn, {idx, res_adj})) return rc;
  if (int rc = check_index(fn, idx, G)) return rc;
  const int64_t n = idx->rows * idx->cols;
  if (!res_adj || res_adj->dtype != SMC_F64 || res_adj->rows * res_adj->cols != n
      || !vec_contiguous(res_adj))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: res_adj must hold size(idx) doubles", fn);
  if (n == 0 || G <= 0 || res_adj->zero_pending) return SMC_OK;
  if (!adj_z) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL adj_z", fn);
  Context& c = ctx();
  if (G > kMaxGroupsFast) {
    if (int rc = group_index(fn, idx, G)) return rc;
    if (int rc = ensure_out(sizeof(double) * (size_t)G)) return rc;
    int64_t blocks = (G + kIdxWarps - 1) / kIdxWarps;
    if (blocks > c.sm_count * 16) blocks = c.sm_count * 16;
    indexing_rev_sorted_kernel<<<(int)blocks, kIdxThreads, 0, c.stream>>>(
        idx->grp_perm, idx->grp_off, static_cast<const double*>(res_adj->data), (int)G,
        c.out_host);
    SMC_CUDA(cudaGetLastError());
    c.launches += 1;
    SMC_CUDA(cudaStreamSynchronize(c.stream));
    for (int64_t g = 0; g < G; ++g) adj_z[g] += c.out_host[g];
    return SMC_OK;
  }
  int grid = c.sm_count;
  int64_t rows_per_warp = (n + (int64_t)grid * kIdxWarps - 1) / ((int64_t)grid * kIdxWarps);
  rows_per_warp = (rows_per_warp + 31) / 32 * 32;
  grid = (int)((n + rows_per_warp * kIdxWarps - 1) / (rows_per_warp * kIdxWarps));
  const size_t smem = sizeof(double) * ((size_t)kIdxWarps * G + kIdxWarps * 32);
  static std::atomic<size_t> attr[16];
  if (attr[c.device & 15] < smem) {
    SMC_CUDA(cudaFuncSetAttribute(indexing_rev_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[c.device & 15] = smem;
  }
  if (int rc = ensure_partials(sizeof(double) * (size_t)grid * G)) return rc;
  if (int rc = ensure_out(sizeof(double) * (size_t)G)) return rc;
  indexing_rev_kernel<<<grid, kIdxThreads, smem, c.stream>>>(
      static_cast<const int*>(idx->data), static_cast<const double*>(res_adj->data), n,
      (int)G, rows_per_warp, c.partials);
  SMC_CUDA(cudaGetLastError());
  indexing_rev_final_kernel<<<(int)((G + 127) / 128), 128, 0, c.stream>>>(c.partials, grid,
                                                                          (int)G, c.out_host);
  SMC_CUDA(cudaGetLastError());
  c.launches += 2;
  SMC_CUDA(cudaStreamSynchronize(c.stream));
  for (int64_t g = 0; g < G; ++g) adj_z[g] += c.out_host[g];
  return SMC_OK;
}

}  // extern "C"
