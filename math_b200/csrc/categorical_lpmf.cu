// categorical_logit_lpmf on a device-resident N x C matrix of log odds, one row
// per outcome (the last un-fused comparator):
//
//   logp = sum_i ( lin[i, y_i - 1] - log_sum_exp(lin[i, :]) )
//   d lin[i, c] = [c == y_i - 1] - softmax(lin[i, :])[c]
//
// i.e. sum_i categorical_logit_lpmf(y_i, lin.row(i)^T) of
// prim/prob/categorical_logit_lpmf.hpp L16-32 (value: L30-31; log_sum_exp as in
// prim/fun/log_sum_exp.hpp L81-93: max + log(sum(exp(v - max)))), which a Stan
// model that adds terms to x * beta writes as a loop over the rows.  One sweep:
// lane = row, the C log odds of a row are staged in shared memory (C <= 32)
// between the max, the sum and the derivative, so lin is read once and d_lin
// written once (2 * N * C * 8 bytes); wider rows use two lanes per row or re-read
// through L1/L2 (see the launch logic).  Deterministic:
// static row -> thread schedule, fixed-order block and grid sums, no atomics.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <limits>

#include "smc_internal.h"
#include "tma_utils.cuh"

using namespace smc;


namespace {

constexpr int kCatThreads = 256;
constexpr int kCatWarps = kCatThreads / 32;

__device__ __forceinline__ void prefetch_l2(const double* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// kStaged (C <= 32): every thread parks the C log odds of its row in shared memory
// ([class][thread], conflict-free) between the max, the sum and the derivative, so
// lin is read once and exp evaluated once per entry, in short rolled loops (a fully
// unrolled register-array version is 112 KB of SASS -- 64 inlined exp bodies -- and
// runs at half the rate).  While a thread works on a row it prefetches its next
// one into L2.  Variants measured at N=1e7, C=32 (ncu): load-then-compute 1.18 ms
// (38 % of the issue slots stalled on the loads); next row through cp.async into a
// second shared-memory slot 1.29 ms (half the warps: the FP64 chains are exposed).
// !kStaged: any C, re-reads through L1/L2.
template <bool kStaged>
__global__ void __launch_bounds__(kCatThreads)
    cat_lpmf_kernel(const double* __restrict__ lin, int64_t ld, int64_t N, int C,
                    const int* __restrict__ y, int y_scalar, double* __restrict__ d_lin,
                    int64_t d_ld, double* __restrict__ partials) {
  extern __shared__ double cat_stage[];  // [C][blockDim.x] when kStaged
  __shared__ double s_lp[kCatWarps], s_bad[kCatWarps];
  const int T = blockDim.x;
  const int64_t stride = (int64_t)gridDim.x * T;
  double* mine = cat_stage + threadIdx.x;
  double lp = 0.0, bad = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)T + threadIdx.x; i < N; i += stride) {
    const double* row = lin + i;
    const int yi = (y ? y[i] : y_scalar) - 1;
    double m = -INFINITY, vy = 0.0, s = 0.0;
    bool finite = true;
#pragma unroll 8
    for (int c = 0; c < C; ++c) {
      const double v = row[(int64_t)c * ld];
      if constexpr (kStaged) mine[c * T] = v;
      finite = finite && isfinite(v);
      m = fmax(m, v);
      vy = c == yi ? v : vy;
    }
    if constexpr (kStaged) {
      // one lane per 128-byte line of the next row block
      if ((threadIdx.x & 15) == 0 && i + stride < N) {
#pragma unroll 8
        for (int c = 0; c < C; ++c) prefetch_l2(row + stride + (int64_t)c * ld);
      }
    }
#pragma unroll 4
    for (int c = 0; c < C; ++c) {
      if constexpr (kStaged) {
        const double e = exp(mine[c * T] - m);
        mine[c * T] = e;
        s += e;
      } else {
        s += exp(row[(int64_t)c * ld] - m);
      }
    }
    if (d_lin) {
      const double inv = 1.0 / s;
      double* drow = d_lin + i;
#pragma unroll 4
      for (int c = 0; c < C; ++c) {
        double e;
        if constexpr (kStaged) e = mine[c * T];
        else e = exp(row[(int64_t)c * ld] - m);
        drow[(int64_t)c * d_ld] = (c == yi ? 1.0 : 0.0) - e * inv;
      }
    }
    if (finite) lp += vy - (m + log(s));
    else bad += 1.0;
  }
  for (int o = 16; o; o >>= 1) {
    lp += __shfl_xor_sync(0xffffffffu, lp, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    s_lp[w] = lp;
    s_bad[w] = bad;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int j = 1; j < (T >> 5); ++j) {  // fixed order
      lp += s_lp[j];
      bad += s_bad[j];
    }
    partials[2 * (size_t)blockIdx.x] = lp;
    partials[2 * (size_t)blockIdx.x + 1] = bad;
  }
}

// 32 < C <= 64 (instantiated for L = 2): L lanes share a row (lane part q owns the classes
// q, q + L, q + 2 L, ...: at most 32 per lane, staged in shared memory as above), the
// row max / sum / picked log odds are combined with shuffles.  A warp covers 32 / L
// consecutive rows, so a load instruction still touches L full 128- or 64-byte runs.
// Same arithmetic per element as the single-lane kernel; the row sum is associated
// per lane part first (fixed order: deterministic).
template <int L>
__global__ void __launch_bounds__(kCatThreads)
    cat_lpmf_multilane_kernel(const double* __restrict__ lin, int64_t ld, int64_t N, int C,
                              const int* __restrict__ y, int y_scalar,
                              double* __restrict__ d_lin, int64_t d_ld,
                              double* __restrict__ partials) {
  extern __shared__ double cat_stage[];  // [ceil(C / L)][kCatThreads]
  __shared__ double s_lp[kCatWarps], s_bad[kCatWarps];
  constexpr int RW = 32 / L;  // rows per warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane / RW, rl = lane - q * RW;
  const int64_t rows_per_cta = (int64_t)kCatWarps * RW;
  const int64_t stride = (int64_t)gridDim.x * rows_per_cta;
  double* mine = cat_stage + threadIdx.x;
  double lp = 0.0, bad = 0.0;
  for (int64_t base = blockIdx.x * rows_per_cta + (int64_t)warp * RW; base < N;
       base += stride) {  // uniform per warp: every lane takes part in the shuffles
    const int64_t i = base + rl;
    const bool live = i < N;
    const double* row = lin + (live ? i : N - 1);
    const int yi = live ? (y ? y[i] : y_scalar) - 1 : 0;
    double m = -INFINITY, vy = 0.0, s = 0.0;
    bool finite = true;
    int j = 0;
#pragma unroll 8
    for (int c = q; c < C; c += L, ++j) {
      const double v = row[(int64_t)c * ld];
      mine[j * kCatThreads] = v;
      finite = finite && isfinite(v);
      m = fmax(m, v);
      vy = c == yi ? v : vy;
    }
    if (rl == 0 && base + stride < N) {  // one lane per run of the next row block
      for (int c = q; c < C; c += L) prefetch_l2(row + stride + (int64_t)c * ld);
    }
#pragma unroll
    for (int o = RW; o < 32; o <<= 1) {
      m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      vy += __shfl_xor_sync(0xffffffffu, vy, o);
      finite = __shfl_xor_sync(0xffffffffu, (int)finite, o) && finite;
    }
    j = 0;
#pragma unroll 4
    for (int c = q; c < C; c += L, ++j) {
      const double e = exp(mine[j * kCatThreads] - m);
      mine[j * kCatThreads] = e;
      s += e;
    }
#pragma unroll
    for (int o = RW; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (d_lin && live) {
      const double inv = 1.0 / s;
      double* drow = d_lin + i;
      j = 0;
#pragma unroll 4
      for (int c = q; c < C; c += L, ++j)
        drow[(int64_t)c * d_ld] = (c == yi ? 1.0 : 0.0) - mine[j * kCatThreads] * inv;
    }
    if (live && q == 0) {
      if (finite) lp += vy - (m + log(s));
      else bad += 1.0;
    }
  }
  for (int o = 16; o; o >>= 1) {
    lp += __shfl_xor_sync(0xffffffffu, lp, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (lane == 0) {
    s_lp[warp] = lp;
    s_bad[warp] = bad;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kCatWarps; ++w) {  // fixed order
      lp += s_lp[w];
      bad += s_bad[w];
    }
    partials[2 * (size_t)blockIdx.x] = lp;
    partials[2 * (size_t)blockIdx.x + 1] = bad;
  }
}

// The derivative-writing sweep as a TMA pipeline (sm_100a): every WARP owns a small
// ring of shared-memory slots and streams its own row blocks through it -- no CTA-wide
// barrier anywhere.  A slot holds a warp tile of RW = 32 / L consecutive rows by all C
// classes, dense column-major as the tensor map delivers it (element (r, c) at
// c * RW + r: the lanes of a warp read one contiguous 256-byte run per instruction,
// conflict-free for L = 1, 2, 4).  Per tile: lane 0 has requested it `stages - 1`
// iterations ago (cp.async.bulk.tensor, completion on the slot's mbarrier); the warp
// takes the row max, overwrites the slot with exp(v - max), then with the partial
// one-hot - softmax, and hands the slot to ONE bulk tensor store (rows >= N clipped
// by the map); the slot of the previous tile, whose store has finished reading by
// then, is refilled.  No per-element global address arithmetic, no LSU traffic to
// HBM; lin is read once and d_lin written once.  L lanes share a row (class c
// belongs to lane part c mod L), combined with shuffles as in the multilane kernel.
// Same arithmetic per element as the kernels above: identical results.
template <int L, bool kStore>
__global__ void __launch_bounds__(512, 1)
    cat_lpmf_tma_kernel(const __grid_constant__ CUtensorMap tm_in,
                        const __grid_constant__ CUtensorMap tm_out, int64_t N, int C,
                        int stages, int slot_doubles, const int* __restrict__ y,
                        int y_scalar, double* __restrict__ partials) {
  extern __shared__ __align__(128) unsigned char cat_tma_raw[];
  constexpr int RW = 32 / L;
  // unroll of the exp / partial loops: 8 amortises the rematerialised polynomial
  // constants over more elements (C=32: 1.035 -> 1.022 ms, C=8: 0.330 -> 0.324) but costs
  // the two-lane form registers it needs elsewhere (C=128: 1.20 -> 1.25 ms)
  constexpr int kCatlUnroll = L == 1 ? 8 : 4;
  const int nwarps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane / RW, rl = lane - q * RW;
  // [nwarps * stages] mbarriers, then the slots (128-byte aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(cat_tma_raw);
  const int bar_bytes = ((nwarps * stages * 8) + 127) & ~127;
  double* ring = reinterpret_cast<double*>(cat_tma_raw + bar_bytes)
                 + (size_t)warp * stages * slot_doubles;
  uint64_t* mybar = bars + warp * stages;
  __shared__ double s_lp[16], s_bad[16];
  if (lane == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(mybar + s, 1);
    fence_barrier_init();
  }
  __syncwarp();
  const uint32_t tile_bytes = (uint32_t)(RW * C * 8);
  const int64_t ntiles = (N + RW - 1) / RW;
  const int64_t t0 = (int64_t)blockIdx.x * nwarps + warp;
  const int64_t tstride = (int64_t)gridDim.x * nwarps;
  const int64_t mine = t0 < ntiles ? (ntiles - t0 + tstride - 1) / tstride : 0;
  const uint64_t pol = policy_evict_first();
  if (lane == 0) {
    for (int k = 0; k < stages - (kStore ? 1 : 0) && k < mine; ++k) {
      mbar_expect_tx(mybar + k, tile_bytes);
      tma_load_2d(ring + (size_t)k * slot_doubles, &tm_in,
                  (int)((t0 + k * tstride) * RW), 0, mybar + k, pol);
    }
  }
  double lp = 0.0, bad = 0.0;
  int slot = 0;
  uint32_t phase = 0;
  // the outcome of the NEXT tile's row is requested one iteration ahead (its latency
  // would otherwise sit in front of every tile's first pass)
  const auto load_y = [&](int64_t k) {
    const int64_t i = (t0 + k * tstride) * RW + rl;
    return k < mine && i < N ? (y ? y[i] : y_scalar) - 1 : 0;
  };
  int yi_next = load_y(0);
  for (int64_t k = 0; k < mine; ++k) {
    const int64_t row0 = (t0 + k * tstride) * RW;
    const int64_t i = row0 + rl;
    const bool live = i < N;
    const int yi = yi_next;
    yi_next = load_y(k + 1);
    double* tile = ring + (size_t)slot * slot_doubles + rl;
    mbar_wait(mybar + slot, phase);
    double m = -INFINITY, vy = 0.0, s = 0.0;
    bool finite = true;
#pragma unroll 8
    for (int c = q; c < C; c += L) {
      const double v = tile[c * RW];
      finite = finite && isfinite(v);
      m = fmax(m, v);
      vy = c == yi ? v : vy;
    }
    if constexpr (L > 1) {
#pragma unroll
      for (int o = RW; o < 32; o <<= 1) {
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        vy += __shfl_xor_sync(0xffffffffu, vy, o);
        finite = __shfl_xor_sync(0xffffffffu, (int)finite, o) && finite;
      }
    }
#pragma unroll kCatlUnroll
    for (int c = q; c < C; c += L) {
      const double e = exp(tile[c * RW] - m);
      if constexpr (kStore) tile[c * RW] = e;
      s += e;
    }
    if constexpr (L > 1) {
#pragma unroll
      for (int o = RW; o < 32; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if constexpr (kStore) {
      const double inv = 1.0 / s;
#pragma unroll kCatlUnroll
      for (int c = q; c < C; c += L)
        tile[c * RW] = (c == yi ? 1.0 : 0.0) - tile[c * RW] * inv;
    }
    if (live && q == 0) {
      if (finite) lp += vy - (m + log(s));
      else bad += 1.0;
    }
    if constexpr (kStore) fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if constexpr (kStore) {
        tma_store_2d(&tm_out, (int)row0, 0, ring + (size_t)slot * slot_doubles, pol);
        bulk_commit();
      }
      // kStore: the slot of tile k - 1 (read by its store, committed one iteration ago)
      // takes tile k + stages - 1; data log odds (no store): this tile's own slot is free
      // again and takes tile k + stages
      const int64_t kn = kStore ? k + stages - 1 : k + stages;
      if (kn < mine) {
        if constexpr (kStore) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        const int sn = !kStore ? slot : slot == 0 ? stages - 1 : slot - 1;
        mbar_expect_tx(mybar + sn, tile_bytes);
        tma_load_2d(ring + (size_t)sn * slot_doubles, &tm_in,
                    (int)((t0 + kn * tstride) * RW), 0, mybar + sn, pol);
      }
    }
    __syncwarp();
    if (++slot == stages) {
      slot = 0;
      phase ^= 1;
    }
  }
  if (kStore && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  for (int o = 16; o; o >>= 1) {
    lp += __shfl_xor_sync(0xffffffffu, lp, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (lane == 0) {
    s_lp[warp] = lp;
    s_bad[warp] = bad;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nwarps; ++w) {  // fixed order
      lp += s_lp[w];
      bad += s_bad[w];
    }
    partials[2 * (size_t)blockIdx.x] = lp;
    partials[2 * (size_t)blockIdx.x + 1] = bad;
  }
}

// out[0] = sum of the per-CTA log densities, out[1] = number of rows with a
// non-finite entry; one warp, lane-strided partial sums combined in lane order.
__global__ void cat_lpmf_final_kernel(const double* __restrict__ partials, int nblocks,
                                      double* __restrict__ out) {
  double lp = 0.0, bad = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 32) {
    lp += partials[2 * (size_t)b];
    bad += partials[2 * (size_t)b + 1];
  }
  for (int o = 16; o; o >>= 1) {
    lp += __shfl_xor_sync(0xffffffffu, lp, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (threadIdx.x == 0) {
    out[0] = lp;
    out[1] = bad;
  }
}

}  // namespace

extern "C" int smc_categorical_logit_lpmf(const smc_matrix* y, int y_scalar,
                                          const smc_matrix* lin, unsigned flags,
                                          double* logp, smc_matrix* d_lin) {
  static const char* fn = "categorical_logit_lpmf";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = refuse_sharded(fn, {y, lin, d_lin})) return rc;
  if (!lin || lin->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: lin must be an f64 device matrix", fn);
  const int64_t N = lin->rows, C = lin->cols;
  if (y && (y->dtype != SMC_I32 || y->rows * y->cols != N
            || !vec_contiguous(y)))
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "%s: size of the random variable (%lld) does not match the rows of the "
                "log odds (%lld)",
                fn, (long long)(y->rows * y->cols), (long long)N);
  const bool lin_var = (flags & SMC_VAR_ALPHA) != 0;
  if (lin_var && (!d_lin || d_lin->dtype != SMC_F64 || d_lin->rows != N || d_lin->cols != C))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: d_lin must be an f64 %lld x %lld matrix", fn,
                (long long)N, (long long)C);
  if (!logp) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL logp", fn);
  *logp = 0.0;
  if (N == 0) return SMC_OK;  // L49-51
  {                           // check_bounded(n, 1, size(beta)), L19 / L38
    int mn = y_scalar, mx = y_scalar;
    if (y) {
      if (int rc = y_range(y, &mn, &mx)) return rc;
    }
    if (mn < 1 || mx > C)
      return fail(SMC_ERR_DOMAIN, "%s: categorical outcome out of support [1, %lld]", fn,
                  (long long)C);
  }
  Context& c = ctx();
  // The LSU kernels, for what the TMA pipeline below cannot take (a layout TMA cannot
  // address, more than 128 classes).  C <= 32: lane = row, C doubles of shared memory
  // per thread (64 KB per CTA at C = 32, three CTAs per SM).  C <= 64 with the
  // derivative wanted: two lanes per row, at most 32 doubles per thread again (N=4e6,
  // C=64: 1.44 -> 1.12 ms).  Everything else re-reads through L1/L2: without a
  // derivative pass (data log odds) exp is evaluated once per entry anyway and staging
  // is pure overhead (N=1e7, C=32: 0.80 vs 0.93 ms; N=4e6, C=64: 0.65 vs 0.81 ms), and
  // four lanes per row (C <= 128) lose to it either way (1.76 vs 2.03 ms).
  const int lanes = !lin_var ? 0 : C <= 32 ? 1 : C <= 64 ? 2 : 0;
  const int threads = kCatThreads;
  const int64_t rows_per_cta = lanes ? threads / lanes : threads;
  const size_t smem = lanes ? sizeof(double) * (size_t)((C + lanes - 1) / lanes) * threads : 0;
  int grid = (int)((N + rows_per_cta - 1) / rows_per_cta);
  const int cap = c.sm_count * 8;  // grid-stride rows: more CTAs than fit just queue
  if (grid > cap) grid = cap;
  if (int rc = ensure_partials(sizeof(double) * 2 * (size_t)grid)) return rc;
  if (int rc = ensure_out(sizeof(double) * 2)) return rc;
  // when lin is data and propto drops the value, only check_finite is left: L22 / L41
  double* d = lin_var ? static_cast<double*>(d_lin->data) : nullptr;
  const int64_t d_ld = lin_var ? d_lin->ld : 0;
  if (d) {
    d_lin->version++;
    d_lin->zero_pending = false;  // overwritten
  }
  if (int rc = realize(lin)) return rc;
  if (int rc = realize(y)) return rc;
  const double* l = static_cast<const double*>(lin->data);
  const int* yp = y ? static_cast<const int*>(y->data) : nullptr;
  // A layout TMA can address (16-byte aligned base and column stride) and at most 128
  // classes: the per-warp TMA pipeline, with the in-place store stream when the
  // derivative is wanted.
  const auto tma_ok = [](const smc_matrix* m) {
    return (reinterpret_cast<uintptr_t>(m->data) & 15) == 0 && (m->ld & 1) == 0;
  };
  static const int knob_mode = [] {
    const char* e = getenv("SMC_CATL_TMA");  // A/B: 0 = the LSU kernels
    return e ? atoi(e) : 1;
  }();
  // (data log odds: only 16 <= C <= 32 gains, N=1e7 C=32 0.80 -> 0.76 ms; narrower and
  // wider rows are faster through the re-reading LSU kernel,
  // profiles/r02/r02_time_categorical_lpmf_tma4.txt)
  if (C <= 128 && knob_mode && N < (1ll << 31) - 64 && tma_ok(lin)
      && (lin_var ? tma_ok(d_lin) : (C >= 16 && C <= 32)) && get_encode()) {
    // Lanes per row, warps and ring depth, measured on B200 (profiles/r02/
    // r02_time_categorical_lpmf_tma.txt): the sweep is co-limited by the FP64 exp, so
    // resident warps matter more than ring depth -- 12 warps x 2 slots beat 8 x 3 at
    // C = 32 (1.03 vs 1.17 ms at N = 1e7); rows wider than 32 classes take two lanes
    // per row (16-row tiles: 128-byte box rows; four lanes -- 64-byte box rows -- halve
    // the TMA rate).  SMC_CATL_L / _W / _S override for A/B.
    static const int knob_l = [] { const char* e = getenv("SMC_CATL_L"); return e ? atoi(e) : 0; }();
    static const int knob_w = [] { const char* e = getenv("SMC_CATL_W"); return e ? atoi(e) : 0; }();
    static const int knob_s = [] { const char* e = getenv("SMC_CATL_S"); return e ? atoi(e) : 0; }();
    const int L = knob_l ? knob_l : C <= 32 ? 1 : 2;
    const int RW = 32 / L;
    const int slot_doubles = (int)(((size_t)RW * C * 8 + 127) / 128 * 16);
    const size_t slot_bytes = (size_t)slot_doubles * 8, budget = 200 * 1024;
    int S = knob_s ? knob_s : L == 1 ? (slot_bytes > 4096 ? 2 : 4) : 3;
    int W = knob_w ? knob_w : L == 1 && slot_bytes > 4096 ? 12 : 16;
    while (W > 1 && (size_t)W * S * slot_bytes > budget) --W;
    const size_t smem_tma = (((size_t)W * S * 8 + 127) & ~(size_t)127)
                            + (size_t)W * S * slot_doubles * 8;
    CUtensorMap tm_in, tm_out;
    if (encode_tmap_f64(&tm_in, lin->data, N, C, lin->ld, RW, (int)C) != CUDA_SUCCESS
        || (lin_var
            && encode_tmap_f64(&tm_out, d_lin->data, N, C, d_lin->ld, RW, (int)C)
                   != CUDA_SUCCESS))
      return fail(SMC_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed", fn);
    if (!lin_var) tm_out = tm_in;  // unused
    const int64_t ntiles = (N + RW - 1) / RW;
    int g = (int)((ntiles + W - 1) / W);
    if (g > c.sm_count) g = c.sm_count;
    if (int rc = ensure_partials(sizeof(double) * 2 * (size_t)g)) return rc;
    auto launch = [&](auto kern) -> int {
      SMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_tma));
      kern<<<g, W * 32, smem_tma, c.stream>>>(tm_in, tm_out, N, (int)C, S, slot_doubles, yp,
                                              y_scalar, c.partials);
      return SMC_OK;
    };
    int rc = lin_var ? (L == 1   ? launch(cat_lpmf_tma_kernel<1, true>)
                        : L == 2 ? launch(cat_lpmf_tma_kernel<2, true>)
                                 : launch(cat_lpmf_tma_kernel<4, true>))
                     : (L == 1   ? launch(cat_lpmf_tma_kernel<1, false>)
                        : L == 2 ? launch(cat_lpmf_tma_kernel<2, false>)
                                 : launch(cat_lpmf_tma_kernel<4, false>));
    if (rc) return rc;
    SMC_CUDA(cudaGetLastError());
    grid = g;
  } else if (lanes) {
    static std::atomic<bool> attr_set[16];
    if (!attr_set[c.device & 15]) {
      SMC_CUDA(cudaFuncSetAttribute(cat_lpmf_kernel<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      SMC_CUDA(cudaFuncSetAttribute(cat_lpmf_multilane_kernel<2>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      attr_set[c.device & 15] = true;
    }
    if (lanes == 1)
      cat_lpmf_kernel<true><<<grid, threads, smem, c.stream>>>(l, lin->ld, N, (int)C, yp,
                                                              y_scalar, d, d_ld, c.partials);
    else
      cat_lpmf_multilane_kernel<2><<<grid, threads, smem, c.stream>>>(
          l, lin->ld, N, (int)C, yp, y_scalar, d, d_ld, c.partials);
  } else {
    cat_lpmf_kernel<false><<<grid, threads, 0, c.stream>>>(l, lin->ld, N, (int)C, yp,
                                                          y_scalar, d, d_ld, c.partials);
  }
  SMC_CUDA(cudaGetLastError());
  cat_lpmf_final_kernel<<<1, 32, 0, c.stream>>>(c.partials, grid, c.out_host);
  SMC_CUDA(cudaGetLastError());
  c.launches += 2;
  SMC_CUDA(cudaStreamSynchronize(c.stream));
  if (c.out_host[1] > 0)  // check_finite(beta), L22 / L41
    return fail(SMC_ERR_DOMAIN, "%s: log odds parameter is not finite", fn);
  if ((flags & SMC_PROPTO) && !lin_var) return SMC_OK;  // L24-26 / L43-45
  *logp = c.out_host[0];
  return SMC_OK;
}
