// categorical_logit_glm_lpmf on sm_100a: the one GEMM-shaped GLM.
//
// reference: stan/math/prim/prob/categorical_logit_glm_lpmf.hpp L43-195
//   lin = x beta + alpha^T                 (N x K . K x C)           L95-96
//   softmax with row-max subtraction, logp                           L97-117
//   T   = -softmax(lin) + onehot(y)                                  L158-159, L165-185
//   d_alpha = colsum(T)                                              L160-169
//   d_beta  = x^T T                        (K x N . N x C)           L171-190
//   d_x     = T beta^T                     (N x C . C x K)           L142-150
// (the reference's "+1 at column y_i" scatters are folded into T).
//
// 4 N K C flops against 2 N K 8 bytes puts this family above the FP64 ridge for
// C >= ~16, so both contractions run on the FP64 tensor pipe (DMMA,
// mma.sync.m8n8k4.f64 -- tcgen05 has no FP64 kind) and x is swept twice:
//   pass 1  cat_lin_kernel   lin tiles (32 rows x C per warp) accumulated in
//           registers, softmax / logp / T fused into the epilogue
//   pass 2  cat_dbeta_kernel K x C accumulators of a CTA live in registers while it
//           streams its row slice; per-CTA partials, fixed-order final sum
// Operand fragments are loaded straight from global memory in the DMMA lane
// layout (a column-major 32-row tile cannot be read conflict-free from shared
// memory in that layout: the k stride is a multiple of all 32 banks); beta sits
// in shared memory with a +4 padded stride, which makes its fragment reads
// conflict-free.  All reductions are fixed-order (no floating-point atomics).
#include <cmath>
#include <cstring>
#include <vector>

#include <limits>

#include "smc_internal.h"
#include "tma_utils.cuh"

namespace smc {

constexpr int kCatThreads = 256;
constexpr int kCatWarps = 8;
constexpr size_t kCatSmemBudget = 200 * 1024;

struct CatArgs {
  int64_t N;
  int K, C, C8;
  const double* x;
  int64_t ldx;
  const int* y;
  int y_scalar;
  const double* beta;   // device, column-major K x C
  const double* alpha;  // device, C
  double* T;            // device, column-major N x C8 (ld = ldT)
  int64_t ldT;
  double* partials;  // pass 1: [grid][1 + 1 + C8]; pass 2: [grid][K*C8]
  double* out;       // pass 1 finalize: [0]=logp, [1]=nonfinite, [2..2+C) d_alpha
  double* d_beta;    // pass 2 finalize: device K x C
  double* d_x;
  int64_t ld_dx;
  int kchunk;   // k columns of beta resident in shared memory at a time (mult of 4)
  int kstride;  // kchunk + 4
  int rows_per_cta;  // pass 2 row slice (multiple of 4)
  int lin_stages;    // pass 1 (TMA): depth of the x / beta^T ring
  const double* beta_t;  // pass 1 (TMA): beta^T, [K][C8 + 4], zero padded
  // pass 1 as a plain matrix product (smc_linear_predictor_matrix): the epilogue
  // stores lin = x beta + alpha^T into T (classes < C only) and stops
  int lin_only;
  int dx_acc;  // cat_dx_kernel adds to d_x (class blocks after the first of a wide C)
};

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ double quad_max(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// ----------------------------------------------------------------- pass 1
template <int NT>
__global__ void __launch_bounds__(kCatThreads, 1)
    cat_lin_kernel(const __grid_constant__ CatArgs a) {
  extern __shared__ __align__(16) unsigned char cat_smem[];
  double* beta_s = reinterpret_cast<double*>(cat_smem);  // [C8][kstride]
  double* alpha_s = beta_s + (size_t)a.C8 * a.kstride;    // [C8]
  double* red_s = alpha_s + a.C8;                         // [kCatWarps][2 + C8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tig = lane & 3;
  const int nchunks = (a.K + a.kchunk - 1) / a.kchunk;
  const int64_t ntiles = (a.N + 31) / 32;
  const int64_t tiles_per_round = (int64_t)gridDim.x * kCatWarps;
  const int64_t nrounds = (ntiles + tiles_per_round - 1) / tiles_per_round;

  for (int c = tid; c < a.C8; c += kCatThreads) alpha_s[c] = c < a.C ? a.alpha[c] : 0.0;

  double lp_acc = 0.0, bad_acc = 0.0;
  double dal[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) dal[nt][0] = dal[nt][1] = 0.0;

  for (int64_t round = 0; round < nrounds; ++round) {
    const int64_t tile = round * tiles_per_round + (int64_t)blockIdx.x * kCatWarps + warp;
    const int64_t r0 = tile * 32;
    double acc[4][NT][2];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    for (int ch = 0; ch < nchunks; ++ch) {
      const int kb = ch * a.kchunk;
      const int kc = min(a.kchunk, ((a.K - kb) + 3) & ~3);
      if (nchunks > 1 || round == 0) {
        // (re)load the beta chunk: beta_s[c][k - kb], zero padded
        __syncthreads();
        for (int idx = tid; idx < a.C8 * kc; idx += kCatThreads) {
          const int c = idx / kc, k = idx - c * kc;
          beta_s[(size_t)c * a.kstride + k]
              = (c < a.C && kb + k < a.K) ? a.beta[(size_t)c * a.K + kb + k] : 0.0;
        }
        __syncthreads();
      }
      if (r0 < a.N) {
        // A fragment addresses: lane holds x[r0 + 8 mt + grp][k + tig]
        const double* xa = a.x + (size_t)(kb + tig) * a.ldx + r0 + grp;
        bool rok[4];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) rok[mt] = r0 + 8 * mt + grp < a.N;
        const double* bfrag = beta_s + (size_t)grp * a.kstride + tig;
#pragma unroll 4
        for (int k = 0; k < kc; k += 4) {
          const bool kok = kb + k + tig < a.K;
          double af[4];
#pragma unroll
          for (int mt = 0; mt < 4; ++mt)
            af[mt] = (rok[mt] && kok) ? __ldg(xa + (size_t)k * a.ldx + 8 * mt) : 0.0;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const double bf = bfrag[(size_t)(8 * nt) * a.kstride + k];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
              dmma(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf);
          }
        }
      }
    }
    if (r0 >= a.N) continue;

    // ---- epilogue: softmax over the C classes of each row (held by a quad)
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int64_t row = r0 + 8 * mt + grp;
      const bool valid = row < a.N;
      if (a.lin_only) {
        if (valid) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int c = 8 * nt + 2 * tig + j;
              if (c < a.C) a.T[(size_t)c * a.ldT + row] = acc[mt][nt][j] + alpha_s[c];
            }
        }
        continue;
      }
      const int yc = valid ? (a.y ? a.y[row] : a.y_scalar) - 1 : 0;
      double m = -INFINITY;
      double lin_y = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = 8 * nt + 2 * tig + j;
          double v = acc[mt][nt][j] + alpha_s[c];
          acc[mt][nt][j] = v;
          if (c < a.C) {
            m = fmax(m, v);
            if (c == yc) lin_y = v;
          }
        }
      m = quad_max(m);
      double se = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = 8 * nt + 2 * tig + j;
          const double e = c < a.C ? exp(acc[mt][nt][j] - m) : 0.0;
          acc[mt][nt][j] = e;
          se += e;
        }
      se = quad_sum(se);
      lin_y = quad_sum(lin_y);
      const double inv = 1.0 / se;
      if (valid) {
        if (tig == 0) {
          const double t = (log(inv) - m) + lin_y;  // L106, L110-116
          lp_acc += t;
          bad_acc += isfinite(t) ? 0.0 : 1.0;
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = 8 * nt + 2 * tig + j;
            if (c < a.C8) {
              double t = acc[mt][nt][j] * -inv;  // neg_softmax_lin, L158-159
              if (c == yc) t += 1.0;             // the "+1 at class y_i" scatters
              if (c >= a.C) t = 0.0;
              a.T[(size_t)c * a.ldT + row] = t;
              dal[nt][j] += t;
            }
          }
      }
    }
  }

  // ---- CTA reduction of logp / nonfinite / d_alpha partials (fixed order)
  // lanes sharing tig hold the same columns: add over grp (xor 4, 8, 16)
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double v = dal[nt][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      dal[nt][j] = v;
    }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lp_acc += __shfl_xor_sync(0xffffffffu, lp_acc, o);
    bad_acc += __shfl_xor_sync(0xffffffffu, bad_acc, o);
  }
  __syncthreads();
  const int rs = 2 + a.C8;
  if (lane == 0) {
    red_s[warp * rs + 0] = lp_acc;
    red_s[warp * rs + 1] = bad_acc;
  }
  if (grp == 0) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) red_s[warp * rs + 2 + 8 * nt + 2 * tig + j] = dal[nt][j];
  }
  __syncthreads();
  for (int j = tid; j < rs; j += kCatThreads) {
    double v = 0.0;
    for (int w = 0; w < kCatWarps; ++w) v += red_s[w * rs + j];
    a.partials[(size_t)blockIdx.x * rs + j] = v;
  }
}

__global__ void cat_lin_finalize_kernel(const double* __restrict__ partials, int nb,
                                        int rs, int C, double* __restrict__ out) {
  for (int j = threadIdx.x; j < rs; j += blockDim.x) {
    double v = 0.0;
    for (int b = 0; b < nb; ++b) v += partials[(size_t)b * rs + j];
    if (j < 2 || j - 2 < C) out[j] = v;
  }
}

// ----------------------------------------------------------------- pass 2
// CTA (bx, by): K-chunk by (MT*8*8 attributes), row slice bx.  Warp w owns the
// attributes [kbase + 64 w', ...): MT blocks of 8.  D[k][c] += x[i][k] T[i][c].
template <int NT, int MT>
__global__ void __launch_bounds__(kCatThreads, 1)
    cat_dbeta_kernel(const __grid_constant__ CatArgs a) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tig = lane & 3;
  const int kbase = blockIdx.y * (kCatWarps * MT * 8) + warp * (MT * 8);
  const int64_t i_begin = (int64_t)blockIdx.x * a.rows_per_cta;
  const int64_t i_end = min(a.N, i_begin + a.rows_per_cta);
  double acc[MT][NT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  if (kbase < a.K) {
    bool kok[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) kok[mt] = kbase + 8 * mt + grp < a.K;
    // A[m = attribute][kk = row]: lane holds x[i + tig][kbase + 8 mt + grp]
    const double* xa = a.x + (size_t)(kbase + grp) * a.ldx + tig;
    // B[kk = row][n = class]: lane holds T[i + tig][8 nt + grp]
    const double* tb = a.T + (size_t)grp * a.ldT + tig;
#pragma unroll 2
    for (int64_t i = i_begin; i < i_end; i += 4) {
      const bool iok = i + tig < a.N;
      double af[MT], bf[NT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        af[mt] = (kok[mt] && iok) ? __ldg(xa + (size_t)(8 * mt) * a.ldx + i) : 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        bf[nt] = iok ? __ldg(tb + (size_t)(8 * nt) * a.ldT + i) : 0.0;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          dmma(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
  }
  // per-CTA partial, layout [k][c8] of this K chunk; every element written
  double* part = a.partials
                 + ((size_t)blockIdx.y * gridDim.x + blockIdx.x)
                       * ((size_t)kCatWarps * MT * 8 * a.C8);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int kl = warp * (MT * 8) + 8 * mt + grp;
        const int c = 8 * nt + 2 * tig + j;
        if (c < a.C8) part[(size_t)kl * a.C8 + c] = acc[mt][nt][j];
      }
}

__global__ void cat_dbeta_finalize_kernel(const double* __restrict__ partials,
                                          int nbx, int kchunk_sz, int K, int C,
                                          int C8, double* __restrict__ d_beta) {
  const int64_t total = (int64_t)K * C;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx / K), k = (int)(idx - (int64_t)c * K);
    const int by = k / kchunk_sz, kl = k - by * kchunk_sz;
    const double* p = partials + (size_t)by * nbx * ((size_t)kchunk_sz * C8)
                      + (size_t)kl * C8 + c;
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    int b = 0;
    const size_t stride = (size_t)kchunk_sz * C8;
    for (; b + 3 < nbx; b += 4) {
      v0 += p[(size_t)(b + 0) * stride];
      v1 += p[(size_t)(b + 1) * stride];
      v2 += p[(size_t)(b + 2) * stride];
      v3 += p[(size_t)(b + 3) * stride];
    }
    for (; b < nbx; ++b) v0 += p[(size_t)b * stride];
    d_beta[(size_t)c * K + k] = (v0 + v1) + (v2 + v3);
  }
}

// ======================================================================= TMA path
// The two kernels above load DMMA fragments straight from global memory and
// stall on those loads (ncu r01: DMMA pipe 59 % / 29 % busy).  The kernels below
// keep the same fragment / accumulator layout but stage x (and T) tiles in a
// shared-memory ring filled by TMA, so the consumer warps only ever wait on an
// mbarrier and read fragments with LDS.  Out-of-range rows / columns of a box
// are zero-filled by TMA and contribute nothing.
//
// pass 1  cat_lin_tma_kernel<NT>   CTA row block = 256 rows (8 warps x 32 rows);
//         ring of {256 rows x 8 cols} x tiles (16 KB, 2 KB contiguous per column);
//         the whole beta (C8 x (K + 4) doubles) resident in shared memory.
// pass 2  cat_dbeta_tma_kernel<NT, MT>   CTA = contiguous row range x one chunk of
//         64 MT attributes; ring of {16 rows x 64 MT cols} x tiles + {16 rows x C8}
//         T tiles; rows of an 8-row block are paired (2 tig, 2 tig + 1) so each
//         lane feeds two DMMAs from one 16-byte LDS.
constexpr int kLinRows = 256;   // rows per CTA row block in pass 1
constexpr int kDbRows = 16;     // rows per stage in pass 2
constexpr int kDbStages = 3;

// beta^T with the class index contiguous and a pitch of C8 + 4 doubles, so that
// the DMMA B fragments of pass 1 read shared memory without bank conflicts
__global__ void cat_beta_transpose_kernel(const double* __restrict__ beta, int K,
                                          int C, int C8p, double* __restrict__ bt) {
  const int64_t total = (int64_t)K * C8p;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / C8p), c = (int)(i - (int64_t)k * C8p);
    bt[i] = c < C ? beta[(size_t)c * K + k] : 0.0;
  }
}

constexpr int kLinWarps = 16;           // pass 1 (TMA): 16 warps x 16 rows
constexpr int kLinThreads = 32 * kLinWarps;
#ifndef SMC_LIN_BOX
#define SMC_LIN_BOX 132
#endif
constexpr int kLinBox = SMC_LIN_BOX;    // rows per TMA box: 128 used + 4 of padding, so
                                        // the column pitch is 4 (mod 16) doubles and the
                                        // 16-byte DMMA A-fragment reads are conflict-free

// Fragment layout of pass 1.  mma.m8n8k4 wants A[m = grp][k = tig] and
// B[k = tig][n = grp] per lane.  Rows and classes are PERMUTED so that the two
// row tiles of a warp, and two class tiles of a pair, sit next to each other in
// shared memory and one 16-byte LDS feeds two DMMAs:
//   row of   (mt, grp)      = r0 + 2 grp + mt
//   class of (nt, n index g) = 16 (nt / 2) + 2 g + (nt & 1)     (paired tiles)
//                            = 16 (NT / 2) + g                  (last tile of an odd NT)
// The accumulator element (mt, nt, j) of a lane therefore belongs to row
// r0 + 2 grp + mt and class lin_class<NT>(nt, 2 tig + j).
template <int NT>
__device__ __forceinline__ int lin_class(int nt, int g) {
  return nt < (NT & ~1) ? 16 * (nt >> 1) + 2 * g + (nt & 1) : 16 * (NT >> 1) + g;
}

template <int NT, int KS>
__global__ void __launch_bounds__(kLinThreads, 1)
    cat_lin_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                       const __grid_constant__ CUtensorMap tmb,
                       const __grid_constant__ CatArgs a) {
  extern __shared__ __align__(1024) unsigned char cat_smem[];
  // stage = two x boxes [KS cols][kLinBox rows] (rows 0-127 and 128-255 of the row
  // block, each with a few rows of padding) followed by the matching slice of
  // beta^T [KS attributes][C8 + 4]; beta streams from L2 with x, so no K x C block
  // has to stay resident and the whole shared memory is ring
  const int kLinStages = a.lin_stages;
  const int C8p = a.C8 + 4;
  constexpr int xbox = kLinBox * KS;  // doubles per x box
  constexpr int NP = NT / 2;          // class-tile pairs
  const uint32_t stage_bytes = (uint32_t)(2 * xbox + C8p * KS) * 8u;
  double* ring = reinterpret_cast<double*>(cat_smem);
  double* alpha_s = ring + (size_t)kLinStages * (stage_bytes / 8);  // [C8]
  double* red_s = alpha_s + a.C8;                                 // [warps][2 + C8]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(red_s + kLinWarps * (2 + a.C8));
  int* rel_cnt = reinterpret_cast<int*>(full_bar + kLinStages);  // warps done per slot
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tig = lane & 3;

  const int64_t nblocks = (a.N + kLinRows - 1) / kLinRows;
  const int ksteps = (a.K + KS - 1) / KS;  // stages per row block
  // this CTA's row blocks: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int64_t my_blocks
      = nblocks > blockIdx.x ? (nblocks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t total = my_blocks * ksteps;  // stage loads of this CTA

  for (int c = tid; c < a.C8; c += kLinThreads) alpha_s[c] = c < a.C ? a.alpha[c] : 0.0;
  if (tid == 0) {
    for (int st = 0; st < kLinStages; ++st) {
      mbar_init(&full_bar[st], 1);
      rel_cnt[st] = 0;
    }
    fence_barrier_init();
  }
  __syncthreads();

  // The ring is refilled by whichever warp releases a slot LAST (shared-memory
  // counter): no warp ever blocks on an "empty" barrier and the refill is issued
  // the moment the slot becomes free.
  const uint64_t pol = policy_evict_first();
  const uint64_t pol_keep = policy_evict_last();  // beta is re-read by every row block
  // stage load (first row `row` of its row block, k-step ks) into slot st
  auto issue = [&](int row, int ks, int st) {
    double* dst = ring + (size_t)st * (stage_bytes / 8);
    mbar_expect_tx(&full_bar[st], stage_bytes);
    tma_load_2d(dst, &tmx, row, ks * KS, &full_bar[st], pol);
    tma_load_2d(dst + xbox, &tmx, row + 128, ks * KS, &full_bar[st], pol);
    tma_load_2d(dst + 2 * xbox, &tmb, 0, ks * KS, &full_bar[st], pol_keep);
  };
  // Every warp tracks (row block, k-step) of the stage load that refills the slot it is
  // about to leave -- stage q + kLinStages -- by increments: whichever warp leaves last
  // issues it, and that warp is the one every other warp is waiting for (two 64-bit
  // divisions on its path showed up as ~8 % of the sweep, profiles/r02/r02_cat_hack.jsonl).
  int n_row = blockIdx.x * kLinRows, n_ks = 0;
  const int row_step = gridDim.x * kLinRows;
  for (int q0 = 0; q0 < kLinStages; ++q0) {
    if (tid == 0 && q0 < total) issue(n_row, n_ks, q0);
    if (++n_ks == ksteps) {
      n_ks = 0;
      n_row += row_step;
    }
  }

  double lp_acc = 0.0, bad_acc = 0.0;
  double dal[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) dal[nt][0] = dal[nt][1] = 0.0;

  // this warp's 16 rows inside the stage: box (warp / 8), local row 16 (warp % 8)
  const int xoff = (warp >> 3) * xbox + 16 * (warp & 7) + 2 * grp;
  int64_t q = 0;
  int st = 0;       // ring slot and mbarrier phase of stage load q, advanced by hand:
  uint32_t ph = 0;  // q % stages and q / stages would be two integer divisions per step
  for (int64_t b = 0; b < my_blocks; ++b) {
    const int64_t r0 = (blockIdx.x + b * gridDim.x) * kLinRows + 16 * warp;
    double acc[2][NT][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

    for (int ks = 0; ks < ksteps; ++ks, ++q) {
#if defined(SMC_CAT_HACK) && (SMC_CAT_HACK & 2)
      if (q < kLinStages)
#endif
      mbar_wait(&full_bar[st], ph);
      // A: x[r0 + 2 grp + {0, 1}][KS ks + 4 h + tig]      one double2
      // B: beta[KS ks + 4 h + tig][16 p + 2 grp + {0, 1}]  one double2 per pair
      const double* stg = ring + (size_t)st * (stage_bytes / 8);
      const double* xa = stg + xoff + tig * kLinBox;
      const double* bfrag = stg + 2 * xbox + tig * C8p + 2 * grp;
#pragma unroll
      for (int h = 0; h < KS / 4; ++h) {
        const double2 af = *reinterpret_cast<const double2*>(xa + 4 * h * kLinBox);
        double2 bf[NP > 0 ? NP : 1];
#pragma unroll
        for (int pr = 0; pr < NP; ++pr)
          bf[pr] = *reinterpret_cast<const double2*>(bfrag + 4 * h * C8p + 16 * pr);
#pragma unroll
        for (int pr = 0; pr < NP; ++pr) {
          dmma(acc[0][2 * pr][0], acc[0][2 * pr][1], af.x, bf[pr].x);
          dmma(acc[1][2 * pr][0], acc[1][2 * pr][1], af.y, bf[pr].x);
          dmma(acc[0][2 * pr + 1][0], acc[0][2 * pr + 1][1], af.x, bf[pr].y);
          dmma(acc[1][2 * pr + 1][0], acc[1][2 * pr + 1][1], af.y, bf[pr].y);
        }
        if constexpr (NT & 1) {
          const double bl = bfrag[4 * h * C8p + 16 * NP - grp];  // class 16 NP + grp
          dmma(acc[0][NT - 1][0], acc[0][NT - 1][1], af.x, bl);
          dmma(acc[1][NT - 1][0], acc[1][NT - 1][1], af.y, bl);
        }
      }
      // release the slot; the last warp out refills it
#if !defined(SMC_CAT_HACK) || !(SMC_CAT_HACK & 4)
      __syncwarp();
      if (lane == 0) {
#if !defined(SMC_CAT_HACK) || !(SMC_CAT_HACK & 8)
        __threadfence_block();
#endif
        if (atomicAdd(&rel_cnt[st], 1) == kLinWarps - 1) {
          rel_cnt[st] = 0;
#if !defined(SMC_CAT_HACK) || !(SMC_CAT_HACK & 8)
          __threadfence_block();
#endif
#if defined(SMC_CAT_HACK) && (SMC_CAT_HACK & 2)
          if (false)
#endif
          if (q + kLinStages < total) issue(n_row, n_ks, st);
        }
      }
#endif
      if (++n_ks == ksteps) {
        n_ks = 0;
        n_row += row_step;
      }
      if (++st == kLinStages) {
        st = 0;
        ph ^= 1u;
      }
    }
    if (r0 >= a.N) continue;
#if defined(SMC_CAT_HACK) && (SMC_CAT_HACK & 1)
    {
      double sacc = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) sacc += acc[0][nt][0] + acc[0][nt][1] + acc[1][nt][0] + acc[1][nt][1];
      if (sacc == 1.2345e-300) a.T[0] = sacc;
      lp_acc += sacc;
      continue;
    }
#endif

    if (a.lin_only) {
      // rows 2 grp, 2 grp + 1 of a class are adjacent: one 16-byte store
      const int64_t row = r0 + 2 * grp;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = lin_class<NT>(nt, 2 * tig + j);
          if (c < a.C) {
            double* tp = a.T + (size_t)c * a.ldT + row;
            const double v0 = acc[0][nt][j] + alpha_s[c], v1 = acc[1][nt][j] + alpha_s[c];
            if (row + 1 < a.N)
              *reinterpret_cast<double2*>(tp) = make_double2(v0, v1);
            else if (row < a.N)
              *tp = v0;
          }
        }
      continue;
    }

    // ---- epilogue: softmax over the C classes of each row (held by a quad)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int64_t row = r0 + 2 * grp + mt;
      const bool valid = row < a.N;
      const int yc = valid ? (a.y ? a.y[row] : a.y_scalar) - 1 : 0;
      double m = -INFINITY;
      double lin_y = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = lin_class<NT>(nt, 2 * tig + j);
          double v = acc[mt][nt][j] + alpha_s[c];
          acc[mt][nt][j] = v;
          if (c < a.C) {
            m = fmax(m, v);
            if (c == yc) lin_y = v;
          }
        }
      m = quad_max(m);
      double se = 0.0;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = lin_class<NT>(nt, 2 * tig + j);
          const double e = c < a.C ? exp(acc[mt][nt][j] - m) : 0.0;
          acc[mt][nt][j] = e;
          se += e;
        }
      se = quad_sum(se);
      lin_y = quad_sum(lin_y);
      const double inv = 1.0 / se;
      if (valid && tig == 0) {
        const double t = (log(inv) - m) + lin_y;  // L106, L110-116
        lp_acc += t;
        bad_acc += isfinite(t) ? 0.0 : 1.0;
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = lin_class<NT>(nt, 2 * tig + j);
          double t = acc[mt][nt][j] * -inv;  // neg_softmax_lin, L158-159
          if (c == yc) t += 1.0;             // the "+1 at class y_i" scatters
          if (c >= a.C || !valid) t = 0.0;
          acc[mt][nt][j] = t;
          dal[nt][j] += t;
        }
    }
    // T rows 2 grp, 2 grp + 1 of a class are adjacent: one 16-byte store
    {
      const int64_t row = r0 + 2 * grp;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = lin_class<NT>(nt, 2 * tig + j);
          double* tp = a.T + (size_t)c * a.ldT + row;
          if (row + 1 < a.N)
            *reinterpret_cast<double2*>(tp) = make_double2(acc[0][nt][j], acc[1][nt][j]);
          else if (row < a.N)
            *tp = acc[0][nt][j];
        }
    }
  }

  // ---- CTA reduction of logp / nonfinite / d_alpha partials (fixed order)
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double v = dal[nt][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      dal[nt][j] = v;
    }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lp_acc += __shfl_xor_sync(0xffffffffu, lp_acc, o);
    bad_acc += __shfl_xor_sync(0xffffffffu, bad_acc, o);
  }
  __syncthreads();
  const int rs = 2 + a.C8;
  if (lane == 0) {
    red_s[warp * rs + 0] = lp_acc;
    red_s[warp * rs + 1] = bad_acc;
  }
  if (grp == 0) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        red_s[warp * rs + 2 + lin_class<NT>(nt, 2 * tig + j)] = dal[nt][j];
  }
  __syncthreads();
  for (int j = tid; j < rs; j += kLinThreads) {
    double v = 0.0;
    for (int w = 0; w < kLinWarps; ++w) v += red_s[w * rs + j];
    a.partials[(size_t)blockIdx.x * rs + j] = v;
  }
}

// pass 2: D[k][c] += x[i][k] T[i][c] over this CTA's rows.
template <int NT, int MT>
__global__ void __launch_bounds__(kCatThreads, 1)
    cat_dbeta_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                         const __grid_constant__ CUtensorMap tmt,
                         const __grid_constant__ CatArgs a) {
  extern __shared__ __align__(1024) unsigned char cat_smem[];
  constexpr int KC = kCatWarps * MT * 8;  // attributes per CTA
  constexpr int XBOX = KC > 256 ? 256 : KC;
  double* xs = reinterpret_cast<double*>(cat_smem);         // [stages][KC][16]
  double* ts = xs + (size_t)kDbStages * KC * kDbRows;         // [stages][C8][16]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ts + (size_t)kDbStages * a.C8 * kDbRows);
  uint64_t* empty_bar = full_bar + kDbStages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tig = lane & 3;
  const int kbase = blockIdx.y * KC;
  const int64_t i_begin = (int64_t)blockIdx.x * a.rows_per_cta;
  const int64_t i_end = min(a.N, i_begin + a.rows_per_cta);
  const int ntiles = i_end > i_begin ? (int)((i_end - i_begin + kDbRows - 1) / kDbRows) : 0;
  const uint32_t stage_bytes = (uint32_t)(KC + a.C8) * kDbRows * 8u;

  if (tid == 0) {
    for (int st = 0; st < kDbStages; ++st) {
      mbar_init(&full_bar[st], 1);
      mbar_init(&empty_bar[st], kCatWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();
  uint64_t pol = 0;
  auto issue = [&](int t) {
    const int st = t % kDbStages;
    const int row = (int)(i_begin + (int64_t)t * kDbRows);
    mbar_expect_tx(&full_bar[st], stage_bytes);
#pragma unroll
    for (int c0 = 0; c0 < KC; c0 += XBOX)
      tma_load_2d(xs + ((size_t)st * KC + c0) * kDbRows, &tmx, row, kbase + c0,
                  &full_bar[st], pol);
    tma_load_2d(ts + (size_t)st * a.C8 * kDbRows, &tmt, row, 0, &full_bar[st], pol);
  };
  if (tid == 0) {
    pol = policy_evict_first();
    for (int t = 0; t < kDbStages && t < ntiles; ++t) issue(t);
  }

  double acc[MT][NT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

  for (int t = 0; t < ntiles; ++t) {
    const int st = t % kDbStages;
    if (tid == 0 && t >= 1 && t - 1 + kDbStages < ntiles) {
      mbar_wait(&empty_bar[(t - 1) % kDbStages], (uint32_t)((t - 1) / kDbStages) & 1u);
      issue(t - 1 + kDbStages);
    }
    mbar_wait(&full_bar[st], (uint32_t)(t / kDbStages) & 1u);
    // rows of an 8-row block are consumed as {0,2,4,6} then {1,3,5,7}: lane tig
    // holds rows 2 tig, 2 tig + 1 in one double2 (A and B use the same order)
    const double* xa = xs + ((size_t)st * KC + warp * (MT * 8) + grp) * kDbRows + 2 * tig;
    const double* tb = ts + ((size_t)st * a.C8 + grp) * kDbRows + 2 * tig;
#pragma unroll
    for (int h = 0; h < kDbRows / 8; ++h) {
      double2 af[MT], bf[NT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        af[mt] = *reinterpret_cast<const double2*>(xa + (size_t)(8 * mt) * kDbRows + 8 * h);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
        bf[nt] = *reinterpret_cast<const double2*>(tb + (size_t)(8 * nt) * kDbRows + 8 * h);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          dmma(acc[mt][nt][0], acc[mt][nt][1], af[mt].x, bf[nt].x);
          dmma(acc[mt][nt][0], acc[mt][nt][1], af[mt].y, bf[nt].y);
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[st]);
  }
  double* part = a.partials
                 + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * ((size_t)KC * a.C8);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int kl = warp * (MT * 8) + 8 * mt + grp;
        const int c = 8 * nt + 2 * tig + j;
        part[(size_t)kl * a.C8 + c] = acc[mt][nt][j];
      }
}

// ----------------------------------------------------------------- d_x = T beta^T
// One thread per row, T row in registers, beta broadcast from global (L1).
template <int CMAX>
__global__ void __launch_bounds__(256)
    cat_dx_kernel(const __grid_constant__ CatArgs a) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= a.N) return;
  double t[CMAX];
#pragma unroll
  for (int c = 0; c < CMAX; ++c) t[c] = c < a.C ? a.T[(size_t)c * a.ldT + row] : 0.0;
  for (int k = 0; k < a.K; ++k) {
    double v = 0.0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < a.C) v = fma(t[c], __ldg(a.beta + (size_t)c * a.K + k), v);
    double* dst = a.d_x + (size_t)k * a.ld_dx + row;
    *dst = a.dx_acc ? *dst + v : v;
  }
}

// d_x = T beta^T on the FP64 tensor pipe, computed transposed so that a lane ends
// up with two consecutive ROWS of one attribute (one 16-byte store):
//   D'[attribute][row] = sum_c beta[attribute][c] T[row][c]
//   A (8 x 4) = beta^T tile from shared memory ([attribute][C8 + 4], conflict-free),
//   B (4 x 8) = T^T tile, held in registers for the warp's 32 rows across every
//   attribute tile of the CTA's chunk.
// CTA (bx, by): row blocks bx, bx + gridDim.x, ... x attribute chunk by.  Written
// once, coalesced in 64-byte runs per attribute; N*K*8 bytes of stores against
// 2 N K C flops: store-bound for C <= ~40, DMMA-bound above.
constexpr int kDxRowsPerWarp = 32;
template <int NT>
__global__ void __launch_bounds__(kCatThreads, 1)
    cat_dx_dmma_kernel(const __grid_constant__ CatArgs a, int kch) {
  extern __shared__ __align__(16) double dx_bs[];  // [kch][C8 + 4]
  constexpr int KS = 2 * NT;                       // k-steps of 4 classes
  const int C8p = a.C8 + 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, tig = lane & 3;
  const int kbase = blockIdx.y * kch;
  for (int idx = tid; idx < kch * C8p; idx += kCatThreads) {
    const int k = idx / C8p;
    dx_bs[idx] = kbase + k < a.K ? a.beta_t[(size_t)kbase * C8p + idx] : 0.0;
  }
  __syncthreads();
  int ktiles = (min(kch, a.K - kbase) + 7) >> 3;
  const int64_t rows_per_cta = (int64_t)kCatWarps * kDxRowsPerWarp;
  for (int64_t rb = blockIdx.x; rb * rows_per_cta < a.N; rb += gridDim.x) {
    const int64_t i0 = rb * rows_per_cta + (int64_t)warp * kDxRowsPerWarp;
    if (i0 >= a.N) continue;
    double tf[4][KS];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
      const int64_t r = i0 + 8 * mt + grp;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int c = 4 * ks + tig;
        tf[mt][ks] = (r < a.N && c < a.C) ? __ldg(a.T + (size_t)c * a.ldT + r) : 0.0;
      }
    }
    for (int kt = 0; kt < ktiles; ++kt) {
      const double* bp = dx_bs + (size_t)(8 * kt + grp) * C8p + tig;
      double af[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) af[ks] = bp[4 * ks];
      double acc[4][2];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) acc[mt][0] = acc[mt][1] = 0.0;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) dmma(acc[mt][0], acc[mt][1], af[ks], tf[mt][ks]);
      const int k = kbase + 8 * kt + grp;
      if (k < a.K) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          const int64_t r = i0 + 8 * mt + 2 * tig;
          double* dst = a.d_x + (size_t)k * a.ld_dx + r;
          if (r + 1 < a.N) {
            double2 v = make_double2(acc[mt][0], acc[mt][1]);
            if (a.dx_acc) {
              const double2 o = *reinterpret_cast<const double2*>(dst);
              v.x += o.x;
              v.y += o.y;
            }
            *reinterpret_cast<double2*>(dst) = v;
          } else if (r < a.N) {
            *dst = a.dx_acc ? *dst + acc[mt][0] : acc[mt][0];
          }
        }
      }
    }
  }
}

template <int NT>
static int run_dx_dmma(const CatArgs& a, dim3 grid, int kch, size_t smem) {
  static size_t attr[16] = {};
  Context& c = ctx();
  if (attr[c.device & 15] < smem) {
    SMC_CUDA(cudaFuncSetAttribute(cat_dx_dmma_kernel<NT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[c.device & 15] = smem;
  }
  cat_dx_dmma_kernel<NT><<<grid, kCatThreads, smem, c.stream>>>(a, kch);
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

// d_x = T beta^T (+= when a.dx_acc).  a.beta_t must point at K x (C8 + 4) doubles of
// device scratch; it is filled here unless `beta_t_ready`.
static int run_cat_dx(const CatArgs& a, bool beta_t_ready) {
  Context& cx = ctx();
  const bool force_fma = knobs().cat_dx_fma;  // A/B switch
  const bool vec_ok = (reinterpret_cast<uintptr_t>(a.d_x) & 15) == 0 && (a.ld_dx & 1) == 0;
  // (8 classes or fewer: two DMMA k-steps per attribute tile do not pay for the
  // fragment loads -- the plain-FMA kernel is faster there, 1.18 vs 1.72 ms at
  // N=4e6, K=128, C=8; from 9 classes on it falls off a cliff, 26.3 vs 2.06 ms at
  // N=2e6, K=512, C=32)
  if (!vec_ok || a.C <= 8 || force_fma) {
    const int g3 = (int)((a.N + 255) / 256);
    if (a.C <= 8)
      cat_dx_kernel<8><<<g3, 256, 0, cx.stream>>>(a);
    else if (a.C <= 16)
      cat_dx_kernel<16><<<g3, 256, 0, cx.stream>>>(a);
    else if (a.C <= 32)
      cat_dx_kernel<32><<<g3, 256, 0, cx.stream>>>(a);
    else
      cat_dx_kernel<64><<<g3, 256, 0, cx.stream>>>(a);
    SMC_CUDA(cudaGetLastError());
    cx.launches += 1;
    return SMC_OK;
  }
  const int C8p = a.C8 + 4;
  if (!beta_t_ready) {
    const int64_t nbt = (int64_t)a.K * C8p;
    cat_beta_transpose_kernel<<<(int)((nbt + 255) / 256), 256, 0, cx.stream>>>(
        a.beta, a.K, a.C, C8p, const_cast<double*>(a.beta_t));
    SMC_CUDA(cudaGetLastError());
    cx.launches += 1;
  }
  int kch = (int)((160 * 1024) / ((size_t)C8p * 8)) & ~7;
  const int k8 = (a.K + 7) & ~7;
  if (kch > k8) kch = k8;
  const int nby = (a.K + kch - 1) / kch;
  const int64_t nrb = (a.N + kCatWarps * kDxRowsPerWarp - 1) / (kCatWarps * kDxRowsPerWarp);
  int nbx = cx.sm_count / nby;
  if (nbx < 1) nbx = 1;
  if (nbx > nrb) nbx = (int)nrb;
  const size_t smem = (size_t)kch * C8p * 8;
  const dim3 grid(nbx, nby);
  int rc;
  switch (a.C8 / 8) {
    case 1: rc = run_dx_dmma<1>(a, grid, kch, smem); break;
    case 2: rc = run_dx_dmma<2>(a, grid, kch, smem); break;
    case 3: rc = run_dx_dmma<3>(a, grid, kch, smem); break;
    case 4: rc = run_dx_dmma<4>(a, grid, kch, smem); break;
    case 5: rc = run_dx_dmma<5>(a, grid, kch, smem); break;
    case 6: rc = run_dx_dmma<6>(a, grid, kch, smem); break;
    case 7: rc = run_dx_dmma<7>(a, grid, kch, smem); break;
    default: rc = run_dx_dmma<8>(a, grid, kch, smem); break;
  }
  if (rc) return rc;
  cx.launches += 1;
  return SMC_OK;
}

template <int NT>
static int run_lin(const CatArgs& a, int grid, size_t smem) {
  static size_t attr[16] = {};
  Context& c = ctx();
  if (attr[c.device & 15] < smem) {
    SMC_CUDA(cudaFuncSetAttribute(cat_lin_kernel<NT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    attr[c.device & 15] = smem;
  }
  cat_lin_kernel<NT><<<grid, kCatThreads, smem, c.stream>>>(a);
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

template <int NT, int MT>
static int run_dbeta(const CatArgs& a, dim3 grid) {
  cat_dbeta_kernel<NT, MT><<<grid, kCatThreads, 0, ctx().stream>>>(a);
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

template <int NT, int KS>
static int run_lin_tma(const CUtensorMap& tmx, const CUtensorMap& tmb, const CatArgs& a,
                       int grid, size_t smem) {
  static size_t attr[16] = {};
  Context& c = ctx();
  if (attr[c.device & 15] < smem) {
    SMC_CUDA(cudaFuncSetAttribute(cat_lin_tma_kernel<NT, KS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    attr[c.device & 15] = smem;
  }
  cat_lin_tma_kernel<NT, KS><<<grid, kLinThreads, smem, c.stream>>>(tmx, tmb, a);
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

template <int NT, int MT>
static int run_dbeta_tma(const CUtensorMap& tmx, const CUtensorMap& tmt,
                         const CatArgs& a, dim3 grid, size_t smem) {
  static size_t attr[16] = {};
  Context& c = ctx();
  if (attr[c.device & 15] < smem) {
    SMC_CUDA(cudaFuncSetAttribute(cat_dbeta_tma_kernel<NT, MT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    attr[c.device & 15] = smem;
  }
  cat_dbeta_tma_kernel<NT, MT><<<grid, kCatThreads, smem, c.stream>>>(tmx, tmt, a);
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

static bool cat_tma_ok(const smc_matrix* x) {
  if (knobs().cat_no_tma) return false;  // A/B switch
  if (!get_encode()) return false;
  if ((reinterpret_cast<uintptr_t>(x->data) & 15) != 0) return false;
  if (x->cols > 1 && (x->ld & 1)) return false;
  return x->rows < 0x7ff00000ll && x->cols >= 1;  // (row coordinates are 32-bit, one sweep of the grid past N)
}

// The two sweeps are also the matrix products either side of an un-fused
// categorical density (smc_linear_predictor_matrix[_adjoint] below):
//   kCatLin  pass 1 alone, epilogue = store lin = x beta + alpha^T into `io` (N x C)
//   kCatAdj  pass 2 alone with T = `io` (N x C adjoints): d_beta = x^T io
enum CatMode { kCatGlm = 0, kCatLin = 1, kCatAdj = 2 };

static bool vec16_ok(const smc_matrix* m) {
  return (reinterpret_cast<uintptr_t>(m->data) & 15) == 0 && (m->ld & 1) == 0;
}

static int launch_cat_impl(CatMode mode, const smc_matrix* y, int y_scalar,
                           const smc_matrix* x, const double* alpha_host,
                           const double* beta_host, int64_t C, unsigned flags,
                           double* logp, double* d_alpha, double* d_beta,
                           smc_matrix* d_x, const smc_matrix* io,
                           const double* params_user = nullptr,
                           double* out_user = nullptr) {
  // params_user / out_user (kCatGlm only): parameters already on the device
  // (beta K x C, then alpha C) and the packed result [logp, #non-finite,
  // d_alpha[C], d_beta[K x C]] left on the device, no host synchronisation --
  // the form the row-sharded multi-GPU driver all-reduces
  Context& cx = ctx();
  for (const smc_matrix* m : {x, y, io})
    if (int rc = realize(m)) return rc;
  if (d_x) d_x->zero_pending = false;  // overwritten
  CatArgs a;
  memset(&a, 0, sizeof(a));
  a.N = x->rows;
  a.K = (int)x->cols;
  a.C = (int)C;
  a.C8 = (a.C + 7) & ~7;
  if (a.C8 > 64)
    return fail(SMC_ERR_UNSUPPORTED,
                "categorical_logit_glm_lpmf: more than 64 classes not supported yet");
  a.lin_only = mode == kCatLin;
  a.x = static_cast<const double*>(x->data);
  a.ldx = x->ld;
  a.y = y ? static_cast<const int*>(y->data) : nullptr;
  a.y_scalar = y_scalar;
  const int NT = a.C8 / 8;

  // device staging: beta (K x C), alpha (C), beta^T (K x (C8 + 4)) for pass 1
  const int C8p = a.C8 + 4;
  const size_t nparam = (size_t)a.K * a.C + a.C;
  const size_t off_bt = (nparam + 15) & ~(size_t)15;  // 128-byte aligned
  if (int rc = ensure_params(sizeof(double) * (off_bt + (size_t)a.K * C8p + 16))) return rc;
  if (params_user) {
    // beta and alpha in one piece (a D2D copy keeps every alignment assumption of
    // the kernels on cx.params_dev)
    SMC_CUDA(cudaMemcpyAsync(cx.params_dev, params_user, sizeof(double) * nparam,
                             cudaMemcpyDefault, cx.stream));
  } else {
    if (a.K && mode != kCatAdj)
      SMC_CUDA(cudaMemcpyAsync(cx.params_dev, beta_host, sizeof(double) * a.K * a.C,
                               cudaMemcpyHostToDevice, cx.stream));
    if (alpha_host)
      SMC_CUDA(cudaMemcpyAsync(cx.params_dev + (size_t)a.K * a.C, alpha_host,
                               sizeof(double) * a.C, cudaMemcpyHostToDevice, cx.stream));
    else
      SMC_CUDA(cudaMemsetAsync(cx.params_dev + (size_t)a.K * a.C, 0, sizeof(double) * a.C,
                               cx.stream));
  }
  a.beta = cx.params_dev;
  a.alpha = cx.params_dev + (size_t)a.K * a.C;
  a.beta_t = cx.params_dev + off_bt;

  // shared-memory plan for beta: whole K if it fits, else chunks
  const size_t fixed = (size_t)a.C8 * 8 + (size_t)kCatWarps * (2 + a.C8) * 8 + 64;
  // kchunk is a multiple of 8 so that kstride = kchunk + 4 is 4 mod 8: the 16
  // lanes of a half-warp (grp 0..3 x tig 0..3) then read 16 distinct banks pairs
  int kchunk = (a.K + 7) & ~7;
  if (kchunk < 8) kchunk = 8;
  while ((size_t)a.C8 * (kchunk + 4) * 8 + fixed > kCatSmemBudget && kchunk > 8)
    kchunk = ((kchunk / 2) + 7) & ~7;
  a.kchunk = kchunk;
  a.kstride = kchunk + 4;
  size_t smem1 = (size_t)a.C8 * a.kstride * 8 + fixed;
  // TMA-staged kernels when x is TMA-addressable and the whole beta fits
  const bool tma = cat_tma_ok(x);
  // pass-1 ring: the widest stage (attributes per stage) that still leaves four
  // stages in ~216 KB
  const size_t lin_fixed = (size_t)a.C8 * 8 + (size_t)kLinWarps * (2 + a.C8) * 8 + 64;
  int ks_lin = 32;
  while (ks_lin > 8
         && 4 * ((size_t)ks_lin * (2 * kLinBox + C8p) * 8 + 16) + lin_fixed > 216 * 1024)
    ks_lin /= 2;
  if (const int v = knobs().cat_ks) {  // tuning knob: attributes per stage
    if (v == 8 || v == 16 || (v == 32 && ks_lin == 32)) ks_lin = v;
  }
  const size_t lin_stage = (size_t)ks_lin * (2 * kLinBox + C8p) * 8;
  a.lin_stages = (int)((216 * 1024 - lin_fixed) / (lin_stage + 16));
  if (a.lin_stages > 12) a.lin_stages = 12;
  const bool lin_tma = tma && a.K >= 1;
  if (lin_tma) smem1 = (size_t)a.lin_stages * (lin_stage + 16) + lin_fixed;

  const bool need_beta = mode == kCatAdj || (mode == kCatGlm && (flags & SMC_VAR_BETA));
  const bool need_dx = mode == kCatGlm && (flags & SMC_VAR_X) && d_x;
  // pass-2 geometry
  const int MT = NT <= 4 ? 8 : 4;
  const int kchunk2 = kCatWarps * MT * 8;
  const int nby = need_beta ? (a.K + kchunk2 - 1) / kchunk2 : 0;
  int nbx = cx.sm_count / (nby > 0 ? nby : 1);
  if (nbx < 1) nbx = 1;
  int64_t rpc = (a.N + nbx - 1) / nbx;
  rpc = (rpc + kDbRows - 1) / kDbRows * kDbRows;  // whole 16-row tiles per CTA
  if (rpc < kDbRows) rpc = kDbRows;
  nbx = (int)((a.N + rpc - 1) / rpc);
  a.rows_per_cta = (int)rpc;

  const int64_t ntiles = (a.N + 31) / 32;
  int grid1 = (int)((ntiles + kCatWarps - 1) / kCatWarps);
  if (grid1 > cx.sm_count) grid1 = cx.sm_count;
  CUtensorMap tm_lin, tm_beta, tm_dbx, tm_dbt;
  if (lin_tma && mode != kCatAdj) {
    if (encode_tmap_f64(&tm_lin, x->data, x->rows, x->cols, x->ld, kLinBox, ks_lin)
            != CUDA_SUCCESS
        || encode_tmap_f64(&tm_beta, a.beta_t, C8p, a.K, C8p, C8p, ks_lin)
               != CUDA_SUCCESS)
      return fail(SMC_ERR_CUDA, "cuTensorMapEncodeTiled failed (categorical pass 1)");
    const int64_t nbt = (int64_t)a.K * C8p;
    cat_beta_transpose_kernel<<<(int)((nbt + 255) / 256), 256, 0, cx.stream>>>(
        a.beta, a.K, a.C, C8p, const_cast<double*>(a.beta_t));
    SMC_CUDA(cudaGetLastError());
    cx.launches += 1;
  }
  const int rs = 2 + a.C8;

  // scratch: T (ldT x C8), out (rs), d_beta_dev (K*C), partials.  The products
  // work in place on the caller's matrix when its layout allows: the TMA pass 1
  // stores row pairs as 16 bytes; pass 2 reads T through a tensor map whose
  // columns >= C are zero-filled (the plain kernel reads C8 columns, so a matrix
  // with C < C8 is copied into the padded scratch first)
  const bool direct = mode == kCatLin   ? (!lin_tma || vec16_ok(io))
                      : mode == kCatAdj ? (tma ? vec16_ok(io) : a.C == a.C8)
                                        : false;
  a.ldT = direct ? io->ld : (a.N + 3) & ~3ll;
  const size_t nT = direct ? 0 : (size_t)a.ldT * a.C8;
  const size_t npart1 = (size_t)grid1 * rs;
  const size_t npart2 = need_beta ? (size_t)nby * nbx * kchunk2 * a.C8 : 0;
  const size_t npart = npart1 > npart2 ? npart1 : npart2;
  const size_t ndb = (size_t)a.K * a.C;
  if (int rc = ensure_scratch(sizeof(double) * (nT + rs + ndb + 8))) return rc;
  if (int rc = ensure_partials(sizeof(double) * (npart + 8))) return rc;
  a.T = direct ? static_cast<double*>(io->data) : cx.scratch;
  double* out_dev = cx.scratch + nT;
  if (mode == kCatAdj && !direct) {
    SMC_CUDA(cudaMemcpy2DAsync(a.T, sizeof(double) * a.ldT, io->data,
                               sizeof(double) * io->ld, sizeof(double) * a.N, a.C,
                               cudaMemcpyDeviceToDevice, cx.stream));
    if (a.C8 > a.C)
      SMC_CUDA(cudaMemsetAsync(a.T + (size_t)a.C * a.ldT, 0,
                               sizeof(double) * a.ldT * (a.C8 - a.C), cx.stream));
  }
  double* d_beta_dev = out_dev + rs;
  a.partials = cx.partials;
  a.out = out_dev;
  a.d_beta = d_beta_dev;
  a.d_x = need_dx ? static_cast<double*>(d_x->data) : nullptr;
  a.ld_dx = need_dx ? d_x->ld : 0;

  int rc = SMC_OK;
  if (mode == kCatAdj) {
    // no pass 1
  } else if (lin_tma) {
#define SMC_LIN_CASE(NTV)                                                         \
  case NTV:                                                                       \
    rc = ks_lin == 32   ? run_lin_tma<NTV, 32>(tm_lin, tm_beta, a, grid1, smem1)  \
         : ks_lin == 16 ? run_lin_tma<NTV, 16>(tm_lin, tm_beta, a, grid1, smem1)  \
                        : run_lin_tma<NTV, 8>(tm_lin, tm_beta, a, grid1, smem1);  \
    break;
    switch (NT) {
      SMC_LIN_CASE(1)
      SMC_LIN_CASE(2)
      SMC_LIN_CASE(3)
      SMC_LIN_CASE(4)
      SMC_LIN_CASE(5)
      SMC_LIN_CASE(6)
      SMC_LIN_CASE(7)
      default:
        rc = ks_lin == 32   ? run_lin_tma<8, 32>(tm_lin, tm_beta, a, grid1, smem1)
             : ks_lin == 16 ? run_lin_tma<8, 16>(tm_lin, tm_beta, a, grid1, smem1)
                            : run_lin_tma<8, 8>(tm_lin, tm_beta, a, grid1, smem1);
        break;
    }
#undef SMC_LIN_CASE
  } else
  switch (NT) {
    case 1: rc = run_lin<1>(a, grid1, smem1); break;
    case 2: rc = run_lin<2>(a, grid1, smem1); break;
    case 3: rc = run_lin<3>(a, grid1, smem1); break;
    case 4: rc = run_lin<4>(a, grid1, smem1); break;
    case 5: rc = run_lin<5>(a, grid1, smem1); break;
    case 6: rc = run_lin<6>(a, grid1, smem1); break;
    case 7: rc = run_lin<7>(a, grid1, smem1); break;
    default: rc = run_lin<8>(a, grid1, smem1); break;
  }
  if (rc) return rc;
  if (mode != kCatAdj) cx.launches += 1;
  if (mode == kCatLin) {
    if (!direct)
      SMC_CUDA(cudaMemcpy2DAsync(io->data, sizeof(double) * io->ld, a.T,
                                 sizeof(double) * a.ldT, sizeof(double) * a.N, a.C,
                                 cudaMemcpyDeviceToDevice, cx.stream));
    return SMC_OK;  // stream-ordered: the consumer runs on the same stream
  }
  if (mode == kCatGlm) {
    cat_lin_finalize_kernel<<<1, 128, 0, cx.stream>>>(cx.partials, grid1, rs, a.C,
                                                      out_dev);
    SMC_CUDA(cudaGetLastError());
    cx.launches += 1;
  }

  if (need_beta && a.K > 0) {
    dim3 g2(nbx, nby);
    const size_t smem2 = (size_t)kDbStages * (kchunk2 + a.C8) * kDbRows * 8
                         + 2 * kDbStages * 8 + 64;
    if (tma) {
      const int xbox = kchunk2 > 256 ? 256 : kchunk2;
      if (encode_tmap_f64(&tm_dbx, x->data, x->rows, x->cols, x->ld, kDbRows, xbox)
              != CUDA_SUCCESS
          || encode_tmap_f64(&tm_dbt, a.T, a.N, direct ? a.C : a.C8, a.ldT, kDbRows,
                             a.C8)
                 != CUDA_SUCCESS)
        return fail(SMC_ERR_CUDA, "cuTensorMapEncodeTiled failed (categorical pass 2)");
      switch (NT) {
        case 1: rc = run_dbeta_tma<1, 8>(tm_dbx, tm_dbt, a, g2, smem2); break;
        case 2: rc = run_dbeta_tma<2, 8>(tm_dbx, tm_dbt, a, g2, smem2); break;
        case 3: rc = run_dbeta_tma<3, 8>(tm_dbx, tm_dbt, a, g2, smem2); break;
        case 4: rc = run_dbeta_tma<4, 8>(tm_dbx, tm_dbt, a, g2, smem2); break;
        case 5: rc = run_dbeta_tma<5, 4>(tm_dbx, tm_dbt, a, g2, smem2); break;
        case 6: rc = run_dbeta_tma<6, 4>(tm_dbx, tm_dbt, a, g2, smem2); break;
        case 7: rc = run_dbeta_tma<7, 4>(tm_dbx, tm_dbt, a, g2, smem2); break;
        default: rc = run_dbeta_tma<8, 4>(tm_dbx, tm_dbt, a, g2, smem2); break;
      }
    } else
    switch (NT) {
      case 1: rc = run_dbeta<1, 8>(a, g2); break;
      case 2: rc = run_dbeta<2, 8>(a, g2); break;
      case 3: rc = run_dbeta<3, 8>(a, g2); break;
      case 4: rc = run_dbeta<4, 8>(a, g2); break;
      case 5: rc = run_dbeta<5, 4>(a, g2); break;
      case 6: rc = run_dbeta<6, 4>(a, g2); break;
      case 7: rc = run_dbeta<7, 4>(a, g2); break;
      default: rc = run_dbeta<8, 4>(a, g2); break;
    }
    if (rc) return rc;
    cx.launches += 1;
    int gf = (int)((ndb + 255) / 256);
    if (gf > cx.sm_count * 4) gf = cx.sm_count * 4;
    cat_dbeta_finalize_kernel<<<gf, 256, 0, cx.stream>>>(cx.partials, nbx, kchunk2,
                                                        a.K, a.C, a.C8, d_beta_dev);
    SMC_CUDA(cudaGetLastError());
    cx.launches += 1;
  }
  if (need_dx && a.K > 0) {
    if (int rc2 = run_cat_dx(a, lin_tma)) return rc2;
  }

  if (out_user) {
    SMC_CUDA(cudaMemcpyAsync(out_user, out_dev, sizeof(double) * (2 + a.C),
                             cudaMemcpyDeviceToDevice, cx.stream));
    if (need_beta && a.K > 0)
      SMC_CUDA(cudaMemcpyAsync(out_user + 2 + a.C, d_beta_dev, sizeof(double) * ndb,
                               cudaMemcpyDeviceToDevice, cx.stream));
    else if (ndb)
      SMC_CUDA(cudaMemsetAsync(out_user + 2 + a.C, 0, sizeof(double) * ndb, cx.stream));
    return SMC_OK;
  }
  // results -> host
  if (int rc2 = ensure_out(sizeof(double) * (rs + ndb + 8))) return rc2;
  if (mode == kCatGlm)
    SMC_CUDA(cudaMemcpyAsync(cx.out_host, out_dev, sizeof(double) * rs,
                             cudaMemcpyDeviceToHost, cx.stream));
  if (need_beta && a.K > 0)
    SMC_CUDA(cudaMemcpyAsync(cx.out_host + rs, d_beta_dev, sizeof(double) * ndb,
                             cudaMemcpyDeviceToHost, cx.stream));
  SMC_CUDA(cudaStreamSynchronize(cx.stream));
  if (mode == kCatGlm) {
    *logp = cx.out_host[0];
    if (d_alpha && (flags & SMC_VAR_ALPHA))
      memcpy(d_alpha, cx.out_host + 2, sizeof(double) * a.C);
  }
  if (d_beta && need_beta) {
    if (a.K > 0)
      memcpy(d_beta, cx.out_host + rs, sizeof(double) * ndb);
  }
  return SMC_OK;
}

int launch_categorical(const smc_matrix* y, int y_scalar, const smc_matrix* x,
                       const double* alpha_host, const double* beta_host,
                       int64_t C, unsigned flags, double* logp, double* d_alpha,
                       double* d_beta, smc_matrix* d_x) {
  return launch_cat_impl(kCatGlm, y, y_scalar, x, alpha_host, beta_host, C, flags, logp,
                         d_alpha, d_beta, d_x, nullptr);
}

// Column sums of an N x C matrix, fixed order: [chunk][c] partials, then one sum
// over the chunks per column.
constexpr int kColsumChunks = 64;
__global__ void colsum_partial_kernel(const double* __restrict__ m, int64_t ld, int64_t N,
                                      double* __restrict__ partials) {
  __shared__ double s_w[8];
  const int c = blockIdx.y;
  const int64_t per = (N + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * per, hi = min(N, lo + per);
  const double* col = m + (size_t)c * ld;
  double v = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) v += col[i];
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v += s_w[w];
    partials[(size_t)c * gridDim.x + blockIdx.x] = v;
  }
}
__global__ void colsum_final_kernel(const double* __restrict__ partials, int nchunks, int C,
                                    double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double v = 0.0;
  for (int j = 0; j < nchunks; ++j) v += partials[(size_t)c * nchunks + j];
  out[c] = v;
}

}  // namespace smc

using namespace smc;

namespace {
// A view of `ncols` columns of m starting at column c0 (same buffer, not owned).
smc_matrix column_block(const smc_matrix* m, int64_t c0, int64_t ncols) {
  smc_matrix v;
  v.data = static_cast<char*>(m->data) + sizeof(double) * (size_t)c0 * m->ld;
  v.rows = m->rows;
  v.cols = ncols;
  v.ld = m->ld;
  v.dtype = SMC_F64;
  v.device = m->device;
  return v;
}
constexpr int64_t kCatBlock = 64;  // classes one launch of the DMMA kernels takes
}  // namespace

extern "C" int smc_linear_predictor_matrix(const smc_matrix* x, const double* beta,
                                           int64_t n_classes, const double* alpha,
                                           smc_matrix* lin_out) {
  static const char* fn = "linear_predictor_matrix";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = refuse_sharded(fn, {x, lin_out})) return rc;
  if (!x || x->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: x must be an f64 device matrix", fn);
  const int64_t N = x->rows, K = x->cols, C = n_classes;
  if (C < 0 || (K > 0 && C > 0 && !beta))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL beta", fn);
  if (!lin_out || lin_out->dtype != SMC_F64 || lin_out->rows != N || lin_out->cols != C)
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "%s: lin_out must be an f64 %lld x %lld device matrix", fn, (long long)N,
                (long long)C);
  if (N == 0 || C == 0) return SMC_OK;
  lin_out->version++;
  for (int64_t c0 = 0; c0 < C; c0 += kCatBlock) {
    const int64_t nc = C - c0 < kCatBlock ? C - c0 : kCatBlock;
    smc_matrix blk = column_block(lin_out, c0, nc);
    if (int rc = launch_cat_impl(kCatLin, nullptr, 1, x, alpha ? alpha + c0 : nullptr,
                                 beta + (size_t)c0 * K, nc, 0, nullptr, nullptr, nullptr,
                                 nullptr, &blk))
      return rc;
  }
  return SMC_OK;
}

extern "C" int smc_linear_predictor_matrix_adjoint(const smc_matrix* x,
                                                   const smc_matrix* adj, double* xt_adj,
                                                   double* colsum) {
  static const char* fn = "linear_predictor_matrix_adjoint";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = refuse_sharded(fn, {x, adj})) return rc;
  if (!x || x->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: x must be an f64 device matrix", fn);
  const int64_t N = x->rows, K = x->cols;
  if (!adj || adj->dtype != SMC_F64 || adj->rows != N)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: adj must be an f64 matrix with rows(x) rows",
                fn);
  const int64_t C = adj->cols;
  if (xt_adj) memset(xt_adj, 0, sizeof(double) * (size_t)K * C);
  if (colsum) memset(colsum, 0, sizeof(double) * (size_t)C);
  if (N == 0 || C == 0 || adj->zero_pending) return SMC_OK;  // x^T 0
  Context& cx = ctx();
  if (xt_adj && K > 0) {
    for (int64_t c0 = 0; c0 < C; c0 += kCatBlock) {
      const int64_t nc = C - c0 < kCatBlock ? C - c0 : kCatBlock;
      smc_matrix blk = column_block(adj, c0, nc);
      if (int rc = launch_cat_impl(kCatAdj, nullptr, 1, x, nullptr, nullptr, nc, 0, nullptr,
                                   nullptr, xt_adj + (size_t)c0 * K, nullptr, &blk))
        return rc;
    }
  }
  if (colsum) {
    const int chunks = N < 65536 ? 1 : kColsumChunks;
    if (int rc = ensure_partials(sizeof(double) * (size_t)C * chunks)) return rc;
    if (int rc = ensure_out(sizeof(double) * (size_t)C)) return rc;
    colsum_partial_kernel<<<dim3(chunks, (unsigned)C), 256, 0, cx.stream>>>(
        static_cast<const double*>(adj->data), adj->ld, N, cx.partials);
    SMC_CUDA(cudaGetLastError());
    colsum_final_kernel<<<(int)((C + 127) / 128), 128, 0, cx.stream>>>(cx.partials, chunks,
                                                                      (int)C, cx.out_host);
    SMC_CUDA(cudaGetLastError());
    cx.launches += 2;
    SMC_CUDA(cudaStreamSynchronize(cx.stream));
    memcpy(colsum, cx.out_host, sizeof(double) * (size_t)C);
  }
  return SMC_OK;
}

// More classes than one launch of the DMMA kernels holds (C8 > 64): the same
// evaluation composed from the stand-alone pieces -- lin = x beta + alpha^T in
// class blocks of 64, the row-wise categorical density on lin (value and
// T = one-hot - softmax), d_beta = x^T T and d_alpha = column sums of T in class
// blocks, d_x = T beta^T accumulated over the blocks.  Two sweeps over x per block.
static int categorical_wide(const char* fn, const smc_matrix* y, int y_scalar,
                            const smc_matrix* x, const double* alpha, const double* beta,
                            int64_t C, unsigned flags, double* logp, double* d_alpha,
                            double* d_beta, smc_matrix* d_x) {
  Context& cx = ctx();
  const int64_t N = x->rows, K = x->cols;
  const int64_t ld = (N + 15) & ~(int64_t)15;
  const size_t bytes = sizeof(double) * (size_t)ld * C;
  void *lin_p = nullptr, *t_p = nullptr;
  if (int rc = cache_alloc(&lin_p, bytes)) return rc;
  if (int rc = cache_alloc(&t_p, bytes)) {
    cache_free(lin_p, bytes);
    return rc;
  }
  smc_matrix lin, T;
  lin.data = lin_p;
  lin.rows = N;
  lin.cols = C;
  lin.ld = ld;
  lin.dtype = SMC_F64;
  lin.device = cx.device;
  T = lin;
  T.data = t_p;
  int rc = smc_linear_predictor_matrix(x, beta, C, alpha, &lin);
  if (!rc) rc = smc_categorical_logit_lpmf(y, y_scalar, &lin, SMC_VAR_ALPHA, logp, &T);
  if (rc == SMC_ERR_DOMAIN) {
    // a non-finite linear predictor: the caller's lazy checks name the operand
    *logp = std::numeric_limits<double>::quiet_NaN();
    rc = SMC_OK;
  } else if (!rc) {
    const bool need_alpha = (flags & SMC_VAR_ALPHA) && d_alpha;
    const bool need_beta = (flags & SMC_VAR_BETA) && d_beta;
    if (need_alpha || need_beta)
      rc = smc_linear_predictor_matrix_adjoint(x, &T, need_beta ? d_beta : nullptr,
                                               need_alpha ? d_alpha : nullptr);
    if (!rc && (flags & SMC_VAR_X) && d_x && K > 0) {
      const size_t off_bt = ((size_t)K * C + 15) & ~(size_t)15;
      rc = ensure_params(sizeof(double) * (off_bt + (size_t)K * (kCatBlock + 4) + 16));
      if (!rc && cudaMemcpyAsync(cx.params_dev, beta, sizeof(double) * (size_t)K * C,
                                 cudaMemcpyHostToDevice, cx.stream) != cudaSuccess)
        rc = fail(SMC_ERR_CUDA, "%s: uploading beta failed", fn);
      d_x->version++;
      for (int64_t c0 = 0; !rc && c0 < C; c0 += kCatBlock) {
        CatArgs a;
        memset(&a, 0, sizeof(a));
        a.N = N;
        a.K = (int)K;
        a.C = (int)(C - c0 < kCatBlock ? C - c0 : kCatBlock);
        a.C8 = (a.C + 7) & ~7;
        a.beta_t = cx.params_dev + off_bt;
        a.T = static_cast<double*>(T.data) + (size_t)c0 * ld;
        a.ldT = ld;
        a.beta = cx.params_dev + (size_t)c0 * K;
        a.d_x = static_cast<double*>(d_x->data);
        a.ld_dx = d_x->ld;
        a.dx_acc = c0 > 0;
        rc = run_cat_dx(a, false);
      }
      if (!rc && cudaStreamSynchronize(cx.stream) != cudaSuccess)
        rc = fail(SMC_ERR_CUDA, "%s: stream synchronise failed", fn);
    }
  }
  cache_free(lin_p, bytes);
  cache_free(t_p, bytes);
  return rc;
}

extern "C" int smc_categorical_logit_glm_device(const smc_matrix* y, int y_scalar,
                                                const smc_matrix* x,
                                                const double* params_dev,
                                                int64_t n_classes, unsigned flags,
                                                double* out_dev, smc_matrix* d_x) {
  static const char* fn = "categorical_logit_glm_device";
  if (int rc = ensure_ctx()) return rc;
  if (int rc = refuse_sharded(fn, {x, y, d_x})) return rc;
  if (!x || x->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: x must be an f64 device matrix", fn);
  const int64_t N = x->rows, K = x->cols, C = n_classes;
  if (y && (y->dtype != SMC_I32 || y->rows * y->cols != N || !vec_contiguous(y)))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: size of y does not match rows of x", fn);
  if (C < 1 || !params_dev || !out_dev)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL or empty params_dev / out_dev", fn);
  if ((flags & SMC_VAR_X)
      && (!d_x || d_x->dtype != SMC_F64 || d_x->rows != N || d_x->cols != K))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: SMC_VAR_X needs d_x shaped like x", fn);
  if (((C + 7) & ~(int64_t)7) > kCatBlock)
    return fail(SMC_ERR_UNSUPPORTED,
                "%s: more than 64 classes take the synchronous smc_categorical_logit_glm",
                fn);
  if (N == 0 || C == 1) {  // L73-75: this rank contributes nothing
    SMC_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double) * (size_t)(2 + C + K * C),
                             ctx().stream));
    return SMC_OK;
  }
  return launch_cat_impl(kCatGlm, y, y_scalar, x, nullptr, nullptr, C, flags, nullptr,
                         nullptr, nullptr, d_x, nullptr, params_dev, out_dev);
}

extern "C" int smc_categorical_logit_glm(const smc_matrix* y, int y_scalar,
                                         const smc_matrix* x, const double* alpha,
                                         const double* beta, int64_t n_classes,
                                         unsigned flags, double* logp,
                                         double* d_alpha, double* d_beta,
                                         smc_matrix* d_x) {
  static const char* fn = "categorical_logit_glm_lpmf";
  if (int rc = ensure_ctx()) return rc;
  if (!x || x->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: x must be an f64 device matrix", fn);
  const int64_t N = x->rows, K = x->cols, C = n_classes;
  if (y && (y->dtype != SMC_I32 || y->rows * y->cols != N || !vec_contiguous(y)))
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "%s: size of y does not match rows of x", fn);  // L68-72
  if (C < 1 || !alpha || (K > 0 && !beta) || !logp)
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: NULL or empty alpha / beta / logp", fn);
  if ((flags & SMC_VAR_X)
      && (!d_x || d_x->dtype != SMC_F64 || d_x->rows != N || d_x->cols != K))
    return fail(SMC_ERR_INVALID_ARGUMENT, "%s: SMC_VAR_X needs d_x shaped like x", fn);
  *logp = 0.0;
  if (N == 0 || C == 1) return SMC_OK;  // L73-75
  {
    int mn, mx;  // check_bounded(y, 1, C), L77
    if (y) {
      if (int rc = y_range(y, &mn, &mx)) return rc;
    } else {
      mn = mx = y_scalar;
    }
    if (mn < 1 || mx > C)
      return fail(SMC_ERR_DOMAIN, "%s: categorical outcome out of support", fn);
  }
  if ((flags & SMC_PROPTO) && !(flags & (SMC_VAR_X | SMC_VAR_ALPHA | SMC_VAR_BETA)))
    return SMC_OK;  // L80-82
  double lp = 0.0;
  if (is_sharded(x) || is_sharded(y) || is_sharded(d_x)) {
    // row-sharded x: the device form on every GPU, packed results reduced (sharded.cu)
    if (!is_sharded(x)) return fail(SMC_ERR_INVALID_ARGUMENT, "%s: y / d_x is sharded but x is not", fn);
    if (((C + 7) & ~(int64_t)7) > kCatBlock)
      return fail(SMC_ERR_UNSUPPORTED, "%s: more than %d classes over sharded x", fn,
                  kCatBlock);
    std::vector<double> params((size_t)K * C + C);
    memcpy(params.data(), beta, sizeof(double) * (size_t)K * C);
    memcpy(params.data() + (size_t)K * C, alpha, sizeof(double) * (size_t)C);
    const double* o;
    if (int rc = run_sharded_categorical(y, y_scalar, x, params.data(), C, flags, d_x, &o))
      return rc;
    lp = o[0];
    if (d_alpha && (flags & SMC_VAR_ALPHA)) memcpy(d_alpha, o + 2, sizeof(double) * (size_t)C);
    if (d_beta && (flags & SMC_VAR_BETA))
      memcpy(d_beta, o + 2 + C, sizeof(double) * (size_t)K * C);
  } else if (((C + 7) & ~(int64_t)7) > kCatBlock) {
    if (int rc = categorical_wide(fn, y, y_scalar, x, alpha, beta, C, flags, &lp, d_alpha,
                                  d_beta, d_x))
      return rc;
  } else if (int rc = launch_categorical(y, y_scalar, x, alpha, beta, C, flags, &lp,
                                         d_alpha, d_beta, d_x))
    return rc;
  if (!std::isfinite(lp)) {  // lazy checks, L122-126
    for (int64_t i = 0; i < K * C; ++i)
      if (!std::isfinite(beta[i]))
        return fail(SMC_ERR_DOMAIN, "%s: Weight vector is not finite", fn);
    for (int64_t i = 0; i < C; ++i)
      if (!std::isfinite(alpha[i]))
        return fail(SMC_ERR_DOMAIN, "%s: Intercept is not finite", fn);
    int ok = 1;
    if (int rc = smc_matrix_all_finite(x, &ok)) return rc;
    if (!ok)
      return fail(SMC_ERR_DOMAIN,
                  "%s: Matrix of independent variables is not finite", fn);
  }
  *logp = lp;
  return SMC_OK;
}
