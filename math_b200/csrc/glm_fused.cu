// Single-pass fused GLM kernel for sm_100a.
//
// One evaluation = ONE sweep over the column-major design matrix x (N x K):
//   theta_i = sum_k x[i,k] beta[k]      (row reduction)
//   link:   per-row log-density term and theta-derivative d_i
//   d_beta[k] = sum_i x[i,k] d_i        (column reduction)   [+ d_x = beta (x) d]
// fused so x is read from HBM exactly once (algorithmic bytes = N*K*8).
//
// Data path: a producer warp streams row tiles of x ({R rows} x {CW cols},
// R = 32*G, CW = 32*S) into a 3-stage shared-memory ring with TMA
// (cp.async.bulk.tensor.2d + mbarrier complete_tx).  Consumer warp (g, s) owns
// row group g (32 rows, lane = row) and column slab s (32 columns): it pulls its
// 32 x values from shared memory into registers ONCE, releases the stage,
// forms its partial dot product, exchanges partials through shared memory,
// evaluates the link, and accumulates d_beta for its 32 columns in registers.
// Column-major tiles make every shared-memory access lane-contiguous
// (conflict-free) and every global access coalesced.
//
// Reductions are deterministic: fixed static tile->CTA schedule, per-thread
// sequential accumulation, xor-butterfly across lanes, fixed-order sum over row
// groups, per-CTA partials in global memory, and a fixed-order final sum by the
// last CTA to finish (ticket counter; order of summation does not depend on
// arrival order).  No floating-point atomics anywhere.
//
// Reference semantics (value + partials): stan/math/prim/prob/
//   normal_id_glm_lpdf.hpp L122-213, bernoulli_logit_glm_lpmf.hpp L105-164,
//   poisson_log_glm_lpmf.hpp L107-161, neg_binomial_2_log_glm_lpmf.hpp L143-244,
//   ordered_logistic_glm_lpmf.hpp L108-207, binomial_logit_glm_lpmf.hpp L104-154.
#include <atomic>
#include <cmath>
#include <cstring>
#include <vector>

#include "glm_link.cuh"
#include "tma_utils.cuh"

namespace smc {

constexpr int kStagesX = 3, kStagesDx = 2;
constexpr uint32_t kSlabBytes = 32 * 32 * 8;  // one warp's 32 rows x 32 columns
constexpr int kCutsPerThread = 4;
constexpr int kFastCuts = 16;  // up to this many cut points: lane-private d_cuts slots

__host__ __device__ inline int link_tab_doubles(int fam, int ncuts, int tab_n) {
  if (fam == kOrdered) return 6 * (ncuts + 1);
  if (fam == kNegBinomial) return 2 * tab_n;
  return 0;
}

// Families whose derivative part of the link is expensive enough (two exp and
// two divisions per row for the ordered link) that evaluating it in every one of
// the S warps of a row group costs more than a second exchange: only the tile's
// lead warp runs the link and publishes d through shared memory + an mbarrier;
// the other warps of the group pick it up (while the previous tile's lead is
// still busy with its deferred log-density part).  The cheap links stay
// redundant: no second exchange on their critical path.
template <int FAM>
constexpr bool kLeadOnlyLink = (FAM == kOrdered);

// ------------------------------------------------------------------- the kernel
// Raw per-row inputs, loaded one tile ahead so their DRAM latency overlaps the
// current tile's arithmetic.
struct RowRaw {
  double y, alpha, aux;
  int yi;
};

template <int FAM>
__device__ __forceinline__ RowRaw load_row(const FusedArgs& a, int64_t row) {
  RowRaw r;
  r.y = a.y_scalar;
  r.yi = 0;
  r.alpha = a.alpha;
  r.aux = a.aux;
  if (row < a.N) {
    if (a.y) {
      if constexpr (FAM == kNormal)
        r.y = static_cast<const double*>(a.y)[row];
      else
        r.yi = static_cast<const int*>(a.y)[row];
    }
    if (a.alpha_vec) r.alpha = a.alpha_vec[row];
    if constexpr (FAM == kBinomial) {
      if (a.aux_ivec) r.aux = (double)a.aux_ivec[row];
    } else {
      if (a.aux_vec) r.aux = a.aux_vec[row];
    }
  }
  return r;
}

// DX: d_x = beta (x) d is written (x is an autodiff variable); a compile-time
// switch because the two cases want the per-row loads in different places.
template <int FAM, int G, bool DX>
__global__ void __launch_bounds__(256, 1)
    glm_fused_kernel(const __grid_constant__ CUtensorMap tmap,
                     const __grid_constant__ CUtensorMap tmap_dx,
                     const __grid_constant__ FusedArgs a) {
  // d_x leaves through per-warp staging slabs and TMA bulk stores; the ring gives
  // up one stage to make room for them
  constexpr int kStages = DX ? kStagesDx : kStagesX;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int S = a.S;
  constexpr int R = 32 * G;
  const int CW = 32 * S;
  const int n_cons_warps = S * G;
  const uint32_t stage_bytes = (uint32_t)R * CW * 8u;

  // shared-memory carve-up
  double* tiles = reinterpret_cast<double*>(smem_raw);
  unsigned char* p = smem_raw + (size_t)kStages * stage_bytes;
  double* dx_stage = reinterpret_cast<double*>(p);  // [warps][32 cols][32 rows]
  if (DX) p += (size_t)n_cons_warps * kSlabBytes;
  double* beta_s = reinterpret_cast<double*>(p);  // CW doubles
  p += (size_t)CW * 8;
  double* cuts_s = reinterpret_cast<double*>(p);  // ncuts (padded) doubles
  p += (size_t)((a.ncuts + 1) & ~1) * 8;
  // link tables: ordered -> 4 doubles per class; neg-binomial -> lgamma / digamma
  // of (y + phi) for y < tab_n
  double* tab_s = reinterpret_cast<double*>(p);
  p += (size_t)link_tab_doubles(FAM, a.ncuts, a.tab_n) * 8;
  // ordered, few cut points: d_cuts accumulators [G][ncuts][32 lanes], one
  // private slot per (cut, lane) of each row group's lead warp
  const bool fast_cuts = FAM == kOrdered && a.ncuts <= kFastCuts;
  double* cacc_s = reinterpret_cast<double*>(p);
  if (fast_cuts) p += (size_t)G * a.ncuts * 32 * 8;
  double* partial_s = reinterpret_cast<double*>(p);  // [2][S][R]
  p += (size_t)2 * S * R * 8;
  double* d1_s = reinterpret_cast<double*>(p);  // [2][R] (ordered)
  double* d2_s = d1_s + 2 * R;
  int* yc_s = reinterpret_cast<int*>(d2_s + 2 * R);  // [2][R]
  p += (size_t)(4 * R) * 8 + (size_t)2 * R * 4;
  p = reinterpret_cast<unsigned char*>(((uintptr_t)p + 7) & ~(uintptr_t)7);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(p);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* part_bar = empty_bar + kStages;  // [G][2]: partial dot products published
  uint64_t* d_bar = part_bar + 2 * G;        // [G][2]: d published by the lead warp
  uint64_t* fin_bar = d_bar + 2 * G;         // the CTAs' partials have landed (last CTA)
  double* hdr_s = reinterpret_cast<double*>(fin_bar + 1);  // [warps][8] row sums
  double* dv_s = hdr_s + (size_t)n_cons_warps * kHdr;  // [2][R] (lead-only links)
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
#ifdef SMC_FUSED_TRACE
  // profiling build: %globaltimer (ns) of thread 0 at the phases of the kernel
  auto stamp = [&](int k) {
    if (tid == 0 && a.trace) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.trace[blockIdx.x * 8 + k] = t;
    }
  };
  stamp(0);
#else
  auto stamp = [&](int) {};
#endif

  // Thread 0 doubles as the TMA producer (a ninth warp would put three warps
  // on one SM sub-partition and cap every thread at 168 registers).  It sets the
  // barriers up and fills the ring before anything else: the first tile is in flight
  // while the CTA stages its parameters (0.3 us of a 10 us evaluation at N = 1e4).
  uint64_t pol = 0;
  if (tid == 0) {
    for (int st = 0; st < kStages; ++st) {
      mbar_init(&full_bar[st], 1);
      mbar_init(&empty_bar[st], n_cons_warps);
    }
    for (int j = 0; j < 2 * G; ++j) mbar_init(&part_bar[j], S);
    for (int j = 0; j < 2 * G; ++j) mbar_init(&d_bar[j], 1);
    mbar_init(fin_bar, 1);
    fence_barrier_init();
    pol = policy_evict_first();
    for (int j = 0; j < kStages; ++j) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)j * gridDim.x;
      if (tile < a.ntiles) {
        mbar_expect_tx(&full_bar[j], stage_bytes);
        tma_load_2d(tiles + (size_t)j * (stage_bytes / 8), &tmap, (int)tile * R, 0,
                    &full_bar[j], pol);
      }
    }
  }

  // parameters -> shared memory
  const double* params = a.params_dev ? a.params_dev : a.inline_params;
  for (int j = tid; j < CW; j += blockDim.x) beta_s[j] = j < a.K ? params[j] : 0.0;
  for (int j = tid; j < a.ncuts; j += blockDim.x) cuts_s[j] = params[a.K + j];
  if (fast_cuts)
    for (int j = tid; j < G * a.ncuts * 32; j += blockDim.x) cacc_s[j] = 0.0;
  LinkTab tab;
  tab.cuts = cuts_s;
  if constexpr (FAM == kOrdered) {
    tab.cls = tab_s;
    for (int c = 1 + tid; c <= a.ncuts + 1; c += blockDim.x) {
      ordered_class_entry(params + a.K, a.ncuts, c, tab_s + 4 * (c - 1));
      ordered_class_l1m(tab_s + 4 * (c - 1), tab_s + 4 * (a.ncuts + 1) + 2 * (c - 1));
    }
  }
  if constexpr (FAM == kNegBinomial) {
    if (a.tab_n > 0) {
      tab.lg = tab_s;
      tab.dg = tab_s + a.tab_n;
      tab.tab_n = a.tab_n;
      const bool need_dg = a.flags & SMC_VAR_AUX;
      for (int j = tid; j < a.tab_n; j += blockDim.x) {
        tab_s[j] = lgamma((double)j + a.aux);
        tab_s[a.tab_n + j] = need_dg ? digamma((double)j + a.aux) : 0.0;
      }
    }
  }
  __syncthreads();
  stamp(1);

  // consumer identity: slab-major, so the S warps of a row group share one SM
  // sub-partition when G = 4 and the row groups spread over all four
  const int s = warp / G, g = warp - s * G;
  double acc[kColsPerThread];
#pragma unroll
  for (int kk = 0; kk < kColsPerThread; ++kk) acc[kk] = 0.0;
  double cacc[kCutsPerThread] = {0, 0, 0, 0};
  RowAcc racc;
  const bool need_beta = a.flags & SMC_VAR_BETA;
  constexpr bool need_dx = DX;
  const bool need_cuts = FAM == kOrdered && (a.flags & SMC_VAR_AUX);

  const uint64_t pol_dx = DX && lane == 0 ? policy_evict_first() : 0;
  {
    // Software pipeline over this CTA's tiles t = blockIdx.x + it * gridDim.x:
    //   stage A(it)  wait for the tile, pull the warp's 32 x 32 slab into
    //                registers, release the ring slot, partial dot product ->
    //                shared memory, ARRIVE on the row group's mbarrier
    //   stage C(it)  WAIT for the S partials, theta, derivative part of the link,
    //                d_beta / d_x / d_cuts accumulation
    //   stage L(it)  log-density part of the link, by the tile's lead warp only
    // executed in the order A(0) | C(0) A(1) L(0) | C(1) A(2) L(1) | ...  The lead
    // role rotates over the S warps of a row group (lead(it) = it mod S) and L(it)
    // comes after the warp has already published its partial for tile it+1, so the
    // transcendental-heavy L never holds up the other warps of the group.
    const int rloc = 32 * g + lane;
    const double* bs = beta_s + 32 * s;
    double xv[kColsPerThread];
    RowRaw nxt;
    const int my_tiles
        = a.ntiles > (int)blockIdx.x ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    auto stage_a = [&](int it) {
      const int st = it % kStages;
      const int par = it & 1;
      const int tile = blockIdx.x + it * gridDim.x;
      // per-row inputs of this tile: in flight while the partials are formed
      if (!need_dx) nxt = load_row<FAM>(a, (int64_t)tile * R + rloc);
      if (tid == 0 && it >= 1) {
        // refill the slot tile it-1 occupied, once every warp has released it
        const int64_t nt = (int64_t)tile + (int64_t)(kStages - 1) * gridDim.x;
        if (nt < a.ntiles) {
          const int rs = (it - 1) % kStages;
          mbar_wait(&empty_bar[rs], (uint32_t)((it - 1) / kStages) & 1u);
          mbar_expect_tx(&full_bar[rs], stage_bytes);
          tma_load_2d(tiles + (size_t)rs * (stage_bytes / 8), &tmap, (int)nt * R, 0,
                      &full_bar[rs], pol);
        }
      }
      mbar_wait(&full_bar[st], (uint32_t)(it / kStages) & 1u);
      const double* xs
          = tiles + (size_t)st * (stage_bytes / 8) + (size_t)(32 * s) * R + rloc;
#pragma unroll
      for (int kk = 0; kk < kColsPerThread; ++kk) xv[kk] = xs[kk * R];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[st]);  // slot is free again
      // partial dot product: four independent FMA chains (fixed association)
      double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
#pragma unroll
      for (int kk = 0; kk < kColsPerThread; kk += 4) {
        p0 = fma(xv[kk + 0], bs[kk + 0], p0);
        p1 = fma(xv[kk + 1], bs[kk + 1], p1);
        p2 = fma(xv[kk + 2], bs[kk + 2], p2);
        p3 = fma(xv[kk + 3], bs[kk + 3], p3);
      }
      const double part = (p0 + p1) + (p2 + p3);
      partial_s[(par * S + s) * R + rloc] = part;
      if (S > 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&part_bar[g * 2 + par]);
      }
    };

    if (my_tiles > 0) {
      if (need_dx) nxt = load_row<FAM>(a, (int64_t)blockIdx.x * R + rloc);
      stage_a(0);
    }
    stamp(2);
    for (int it = 0; it < my_tiles; ++it) {
      const int par = it & 1;
      const int tile = blockIdx.x + it * gridDim.x;
      const int64_t row = (int64_t)tile * R + rloc;
      const bool valid = row < a.N;
      const bool lead = (it % S) == s;
      const RowRaw cur = nxt;
      // d_x is written: the per-row inputs of the NEXT tile are requested before
      // this tile's 32 streaming stores per thread, not behind them (ncu: 17 % of
      // the samples waited for y at the top of this loop)
      if (need_dx && it + 1 < my_tiles)
        nxt = load_row<FAM>(a, row + (int64_t)gridDim.x * R);
      RowIn<FAM> in;
      if constexpr (FAM == kNormal)
        in.y = cur.y;
      else
        in.y = a.y ? (double)cur.yi : cur.y;
      in.alpha = cur.alpha;
      in.aux = cur.aux;

      // ---- stage C: every warp of the row group adds the S partials in the same
      // fixed tree, so all of them hold the identical theta
      const bool shared_d = kLeadOnlyLink<FAM> && S > 1;
      double d1 = 0, d2 = 0, d = 0;
      LinkStash<FAM> stash;
      if (!shared_d || lead) {
        double xb;
        if (S > 1) {
          mbar_wait(&part_bar[g * 2 + par], (uint32_t)(it >> 1) & 1u);
          double q[8];
#pragma unroll
          for (int ss = 0; ss < 8; ++ss)
            q[ss] = ss < S ? partial_s[(par * S + ss) * R + rloc] : 0.0;
          xb = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
        } else {
          xb = partial_s[par * R + rloc];
        }
        d = link_d<FAM>(a, xb, in, valid, lead, row, racc, tab, d1, d2, stash);
      }
      if (shared_d) {
        if (lead) {
          dv_s[par * R + rloc] = d;
          __syncwarp();
          if (lane == 0) mbar_arrive(&d_bar[g * 2 + par]);
        } else {
          mbar_wait(&d_bar[g * 2 + par], (uint32_t)(it >> 1) & 1u);
          d = dv_s[par * R + rloc];
        }
      }

      if (need_beta) {
#pragma unroll
        for (int kk = 0; kk < kColsPerThread; ++kk) acc[kk] = fma(xv[kk], d, acc[kk]);
      }
      if constexpr (DX) {
        // d_x = beta (x) d: the warp fills its staging slab (lane = row, conflict
        // free) and hands it to ONE TMA bulk store -- no per-element addresses, no
        // LSU store traffic; rows >= N and columns >= K are clipped by the map.
        // The previous tile's store has long since read the slab.
        double* xo = dx_stage + (size_t)warp * (kSlabBytes / 8);
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
#pragma unroll
        for (int kk = 0; kk < kColsPerThread; ++kk) xo[kk * 32 + lane] = bs[kk] * d;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d_plain(&tmap_dx, tile * R + 32 * g, 32 * s, xo);
          bulk_commit();
        }
      }
      if constexpr (FAM == kOrdered) {
        if (need_cuts && fast_cuts) {
          // ordered_logistic_glm_lpmf.hpp L197-207: cuts'[y-1] += d2, cuts'[y-2] -= d1.
          // Each lane of the tile's lead warp owns a private slot per cut point,
          // so the scatter is conflict-free and its order of additions is fixed
          // (tile order; consecutive leads are ordered by the partial mbarrier).
          if (lead && valid) {
            const int yy = (int)in.y;
            double* cs = cacc_s + (size_t)g * a.ncuts * 32 + lane;
            if (yy - 1 < a.ncuts) cs[(yy - 1) * 32] += d2;
            if (yy >= 2) cs[(yy - 2) * 32] -= d1;
          }
        } else if (need_cuts) {
          // many cut points: owner-computes loop over the rows of the tile
          if (lead) {
            d1_s[par * R + rloc] = d1;
            d2_s[par * R + rloc] = d2;
            yc_s[par * R + rloc] = valid ? (int)in.y : 0;
          }
          if (S > 1)
            group_bar(1 + g, 32 * S);
          else
            __syncwarp();
          const int tg = s * 32 + lane;
#pragma unroll
          for (int j = 0; j < kCutsPerThread; ++j) {
            const int c = tg + j * 32 * S;
            if (c < a.ncuts) {
              double v = cacc[j];
              for (int r = 0; r < 32; ++r) {
                const int yy = yc_s[par * R + 32 * g + r];
                if (yy - 1 == c) v += d2_s[par * R + 32 * g + r];
                if (yy - 2 == c) v -= d1_s[par * R + 32 * g + r];
              }
              cacc[j] = v;
            }
          }
        }
      }
      // ---- stage A of the next tile, then the deferred stage L of this one
      if (it + 1 < my_tiles) stage_a(it + 1);
      if (lead) link_lp<FAM>(a, stash, racc, tab);
    }
  }

  // ------------------------------------------------ CTA-level reduction
  if (DX && lane == 0) bulk_wait_read0();  // the last d_x slabs have left
  __syncthreads();  // every tile consumed; the ring is reusable as scratch
  stamp(3);
  double* red = tiles;  // [G][pstride]
  const int ps = a.pstride;
  // (only the d_cuts columns can stay unwritten below)
  for (int j = tid; j < G * ps; j += blockDim.x)
    if (j % ps >= kHdr + CW) red[j] = 0.0;
  if (need_cuts) __syncthreads();
  {
    // column sums of the warp's 32 x 32 slab: a transposing butterfly -- each step
    // halves the number of columns a lane carries (the lane keeps one half and sends
    // the other to its partner), 31 shuffles instead of 32 five-step warp sums; lane l
    // ends up with the total of column l (fixed association)
    static_assert(kColsPerThread == 32, "one column per lane");
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool upper = lane & off;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const double send = upper ? acc[i] : acc[i + off];
        const double keep = upper ? acc[i + off] : acc[i];
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    red[g * ps + kHdr + 32 * s + lane] = acc[0];
    {
      // every warp led some of its row group's tiles: row sums per warp
      const double v0 = warp_sum(racc.lp), v1 = warp_sum(racc.sd),
                   v2 = warp_sum(racc.s2), v3 = warp_sum(racc.s3),
                   vb = warp_sum((double)racc.bad);
      if (lane == 0) {
        double* h = hdr_s + warp * kHdr;
        h[SMC_OUT_LOGP] = v0;
        h[SMC_OUT_SUM_D] = v1;
        h[SMC_OUT_AUX] = v2;
        h[SMC_OUT_NONFINITE] = vb;
        h[SMC_OUT_AUX2] = v3;
        h[5] = h[6] = h[7] = 0.0;
      }
    }
    if (need_cuts && fast_cuts) {
      if (s == 0)
        for (int c = 0; c < a.ncuts; ++c) {
          const double v = warp_sum(cacc_s[((size_t)g * a.ncuts + c) * 32 + lane]);
          if (lane == 0) red[g * ps + kHdr + CW + c] = v;
        }
    } else if (need_cuts) {
      const int tg = s * 32 + lane;
#pragma unroll
      for (int j = 0; j < kCutsPerThread; ++j) {
        const int c = tg + j * 32 * S;
        if (c < a.ncuts) red[g * ps + kHdr + CW + c] = cacc[j];
      }
    }
  }
  __syncthreads();
  double* my_partial = a.partials + (size_t)blockIdx.x * ps;
  for (int j = tid; j < ps; j += blockDim.x) {
    double v = 0.0;
    if (j < kHdr) {
      for (int w = 0; w < n_cons_warps; ++w) v += hdr_s[w * kHdr + j];  // fixed order
    } else {
      for (int gg = 0; gg < G; ++gg) v += red[gg * ps + j];
    }
    my_partial[j] = v;
  }

  // ------------------------------------------------ grid-level reduction
  stamp(4);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(a.counter, 1u);
    s_last = ticket == gridDim.x - 1;
  }
  __syncthreads();
  stamp(5);
  if (s_last) {
    __threadfence();
    const int nb = gridDim.x;
    // The CTAs' partials come into the (now free) ring with bulk copies when they fit:
    // one thread walking 148 L2 lines per column, sixteen loads at a time, was 4 us of
    // the 11 us this kernel takes at N = 1e4 (profiles/r02/r02_fused_trace_cfg1.txt);
    // the copy engine fetches the same 160 KB in a third of that.
    const uint32_t fin_bytes = (uint32_t)nb * ps * 8u;
    const bool staged = fin_bytes <= (uint32_t)kStages * stage_bytes;
    const double* part = a.partials;
    if (staged) {
      if (tid == 0) {
        // the ring was written by generic-proxy stores (red), the partials by other
        // CTAs' generic-proxy stores: order both before the async-proxy copy
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_expect_tx(fin_bar, fin_bytes);
        for (uint32_t off = 0; off < fin_bytes; off += 32768u) {
          const uint32_t n = fin_bytes - off < 32768u ? fin_bytes - off : 32768u;
          bulk_load_1d(reinterpret_cast<unsigned char*>(tiles) + off,
                       reinterpret_cast<const unsigned char*>(a.partials) + off, n, fin_bar);
        }
      }
      mbar_wait(fin_bar, 0);
      part = tiles;
    }
    for (int j = tid; j < ps; j += blockDim.x) {
      // fixed order over CTAs in eight interleaved chains (still one fixed
      // association), so that sixteen independent loads are in flight per thread
      double v8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      int b = 0;
      if (staged) {
#pragma unroll 2
        for (; b + 7 < nb; b += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) v8[u] += part[(size_t)(b + u) * ps + j];
        }
        for (; b < nb; ++b) v8[0] += part[(size_t)b * ps + j];
      } else {
#pragma unroll 2
        for (; b + 7 < nb; b += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) v8[u] += __ldcg(part + (size_t)(b + u) * ps + j);
        }
        for (; b < nb; ++b) v8[0] += __ldcg(part + (size_t)b * ps + j);
      }
      double v = ((v8[0] + v8[1]) + (v8[2] + v8[3])) + ((v8[4] + v8[5]) + (v8[6] + v8[7]));
      if (j == SMC_OUT_LOGP) v += a.c0;
      // packed output: header, d_beta[K], d_cuts[ncuts]
      if (j < kHdr) {
        if (!a.out_skip_header) a.out[j] = v;
      } else if (j < kHdr + CW) {
        if (j - kHdr < a.K) a.out[a.out_beta_off + j] = v;
      } else if (j - kHdr - CW < a.ncuts)
        a.out[kHdr + a.out_K_total + (j - kHdr - CW)] = v;
    }
    __syncthreads();  // (uniform: s_last is shared) every store to out is issued
    stamp(6);
    if (tid == 0) {
      *a.counter = 0;  // ready for the next launch on this stream
      if (a.done_flag) {
        __threadfence_system();  // the packed result is visible to the host first
        *reinterpret_cast<volatile unsigned long long*>(a.done_flag) = a.done_val;
      }
    }
  }
}

// ---------------------------------------------------------------- host side
static void tile_shape(int64_t K, int* S, int* G) {
  int s = (int)((K + kColsPerThread - 1) / kColsPerThread);
  if (s < 1) s = 1;
  int g = 1;
  while (g * 2 * s <= 8) g *= 2;
  *S = s;
  *G = g;
}

bool fused_layout_ok(const smc_matrix* x) {
  if (!x || x->dtype != SMC_F64) return false;
  if (x->cols < 1) return false;
  if (x->rows < 1 || x->rows > 0x7fffff00ll) return false;
  if ((reinterpret_cast<uintptr_t>(x->data) & 15) != 0) return false;
  if (x->cols > 1 && (x->ld & 1)) return false;  // TMA: 16-byte global strides
  return get_encode() != nullptr;
}

bool fused_supported(const smc_matrix* x) {
  return fused_layout_ok(x) && x->cols <= kMaxFusedK;
}

static int get_tmap(const smc_matrix* xc, int R, int CW, CUtensorMap* out) {
  smc_matrix* x = const_cast<smc_matrix*>(xc);
  std::lock_guard<std::mutex> lock(cache_mutex());  // x is shared between chains
  if (x->tmap_rows == R && x->tmap_cols == CW) {
    *out = x->tmap;
    return SMC_OK;
  }
  if (!get_encode()) return fail(SMC_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  CUresult r = encode_tmap_f64(&x->tmap, x->data, x->rows, x->cols, x->ld, R, CW);
  if (r != CUDA_SUCCESS)
    return fail(SMC_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d) for %lld x %lld ld %lld box "
                "%d x %d",
                (int)r, (long long)x->rows, (long long)x->cols, (long long)x->ld,
                R, CW);
  x->tmap_rows = R;
  x->tmap_cols = CW;
  *out = x->tmap;
  return SMC_OK;
}

template <int FAM, int G, bool DX>
static int launch_tgx(const CUtensorMap& tmap, const CUtensorMap& tmap_dx,
                      const FusedArgs& a, int grid, int threads, size_t smem) {
  static std::atomic<size_t> attr_smem[16];
  Context& c = ctx();
  if (attr_smem[c.device & 15] < smem) {
    SMC_CUDA(cudaFuncSetAttribute(glm_fused_kernel<FAM, G, DX>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    attr_smem[c.device & 15] = smem;
  }
  glm_fused_kernel<FAM, G, DX><<<grid, threads, smem, c.stream>>>(tmap, tmap_dx, a);
  SMC_CUDA(cudaGetLastError());
  c.launches += 1;
  return SMC_OK;
}

template <int FAM, int G>
static int launch_tg(const CUtensorMap& tmap, const CUtensorMap& tmap_dx,
                     const FusedArgs& a, int grid, int threads, size_t smem) {
  return ((a.flags & SMC_VAR_X) && a.d_x)
             ? launch_tgx<FAM, G, true>(tmap, tmap_dx, a, grid, threads, smem)
             : launch_tgx<FAM, G, false>(tmap, tmap_dx, a, grid, threads, smem);
}

template <int FAM>
static int launch_t(const CUtensorMap& tmap, const CUtensorMap& tmap_dx,
                    const FusedArgs& a, int grid, int threads, size_t smem) {
  switch (a.G) {
    case 1:
      return launch_tg<FAM, 1>(tmap, tmap_dx, a, grid, threads, smem);
    case 2:
      return launch_tg<FAM, 2>(tmap, tmap_dx, a, grid, threads, smem);
    case 4:
      return launch_tg<FAM, 4>(tmap, tmap_dx, a, grid, threads, smem);
    default:
      return launch_tg<FAM, 8>(tmap, tmap_dx, a, grid, threads, smem);
  }
}

// Dynamic shared memory of one CTA for a tile shape (mirrors the kernel's carve-up).
static size_t fused_smem_bytes(const GlmCall& c, const FusedArgs& a, bool dx) {
  const int R = 32 * a.G, CW = 32 * a.S;
  const int kStages = dx ? kStagesDx : kStagesX;
  const size_t stage_bytes = (size_t)R * CW * 8;
  return kStages * stage_bytes + (dx ? (size_t)a.S * a.G * kSlabBytes : 0)
         + (size_t)CW * 8
         + (size_t)((a.ncuts + 1) & ~1) * 8
         + (size_t)link_tab_doubles(c.family, a.ncuts, a.tab_n) * 8
         + (c.family == kOrdered && a.ncuts <= kFastCuts ? (size_t)a.G * a.ncuts * 32 * 8
                                                         : 0)
         + (size_t)2 * a.S * R * 8
         + (size_t)4 * R * 8 + (size_t)2 * R * 4 + 8 + 2 * kStages * 8
         + (size_t)4 * a.G * 8 + 8 + (size_t)a.S * a.G * kHdr * 8
         + (c.family == kOrdered && a.S > 1 ? (size_t)2 * R * 8 : 0);  // dv_s
}

constexpr size_t kMaxDynamicSmem = 227 * 1024;  // opt-in limit per CTA on sm_100

#ifdef SMC_FUSED_TRACE
static void dump_fused_trace(const FusedArgs& a, int grid) {
  if (const char* f = getenv("SMC_FUSED_TRACE_FILE")) {
    std::vector<unsigned long long> h(8 * 1024);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), a.trace, 8 * 8 * 1024, cudaMemcpyDeviceToHost);
    if (FILE* fp = fopen(f, "wb")) {
      long long g = grid;
      fwrite(&g, 8, 1, fp);
      fwrite(h.data(), 8, 8 * 1024, fp);
      fclose(fp);
    }
  }
}
#endif

int launch_glm_fused(const GlmCall& c) {
  Context& cx = ctx();
  const smc_matrix* x = c.x;
  FusedArgs a;
  if (int rc = prepare_args(c, &a)) return rc;
  const bool dx = (a.flags & SMC_VAR_X) && c.d_x;
  tile_shape(a.K, &a.S, &a.G);
  // fewer row groups if the per-group tables push the CTA over the shared-memory
  // limit (many cut points, or the widest tiles with every optional buffer)
  while (a.G > 1 && fused_smem_bytes(c, a, dx) > kMaxDynamicSmem) a.G /= 2;
  const int R = 32 * a.G, CW = 32 * a.S;
  a.ntiles = (int)((a.N + R - 1) / R);
  if (a.ncuts > kCutsPerThread * 32 * a.S || a.ncuts > kMaxCuts)
    return fail(SMC_ERR_UNSUPPORTED, "too many cut points for the fused kernel");

  // parameters: by value when they fit, else staged through device memory
  const int nparam = a.K + a.ncuts;
  if (c.params_dev) {
    a.params_dev = c.params_dev;
  } else if (nparam <= kMaxParamDoubles) {
    memcpy(a.inline_params, c.beta_host, sizeof(double) * a.K);
    if (a.ncuts) memcpy(a.inline_params + a.K, c.cuts_host, sizeof(double) * a.ncuts);
  } else {
    if (int rc = ensure_params(sizeof(double) * nparam)) return rc;
    // pageable source: the copy is staged by the driver before returning
    SMC_CUDA(cudaMemcpyAsync(cx.params_dev, c.beta_host, sizeof(double) * a.K,
                             cudaMemcpyHostToDevice, cx.stream));
    if (a.ncuts)
      SMC_CUDA(cudaMemcpyAsync(cx.params_dev + a.K, c.cuts_host,
                               sizeof(double) * a.ncuts, cudaMemcpyHostToDevice,
                               cx.stream));
    a.params_dev = cx.params_dev;
  }

  a.pstride = kHdr + CW + ((a.ncuts + 3) & ~3);
  int grid = cx.sm_count;
  if (grid > a.ntiles) grid = a.ntiles;
  // the fewest CTAs that finish in the same number of tile rounds: 157 tiles are two
  // rounds on 148 CTAs and on 79, and the last CTA then sums half as many partials
  // (the tail of a small evaluation, nothing for a large one)
  if (grid > 0) {
    const int rounds = (a.ntiles + grid - 1) / grid;
    grid = (a.ntiles + rounds - 1) / rounds;
  }
  if (int rc = ensure_partials(sizeof(double) * (size_t)grid * a.pstride)) return rc;
  a.partials = cx.partials;
  a.counter = cx.counter;
  a.out = c.out;
  a.done_flag = c.done_flag;
  a.done_val = c.done_val;
  a.out_beta_off = c.out_beta_off;
  a.out_K_total = c.out_K_total > 0 ? c.out_K_total : a.K;
  a.out_skip_header = c.out_skip_header ? 1 : 0;
  cx.flag_armed = c.done_flag != nullptr;
#ifdef SMC_FUSED_TRACE
  static unsigned long long* trace_dev = nullptr;
  if (!trace_dev) cudaMalloc(&trace_dev, 8 * 8 * 1024);
  a.trace = trace_dev;
#endif

  CUtensorMap tmap, tmap_dx;
  if (int rc = get_tmap(x, R, CW, &tmap)) return rc;
  tmap_dx = tmap;  // unused unless d_x is written
  if (dx) {
    if (int rc = get_tmap(c.d_x, 32, 32, &tmap_dx)) return rc;
  }

  const int kStages = dx ? kStagesDx : kStagesX;
  const size_t stage_bytes = (size_t)R * CW * 8;
  const size_t smem = fused_smem_bytes(c, a, dx);
  if (smem > kMaxDynamicSmem)
    return fail(SMC_ERR_UNSUPPORTED, "fused kernel: %zu bytes of shared memory needed",
                smem);
  const size_t red_bytes = (size_t)a.G * a.pstride * 8;
  if (red_bytes > kStages * stage_bytes)
    return fail(SMC_ERR_UNSUPPORTED, "reduction scratch exceeds the tile ring");
  const int threads = 32 * a.S * a.G;

#ifdef SMC_FUSED_TRACE
  struct Dump {
    const FusedArgs& a;
    int grid;
    ~Dump() { dump_fused_trace(a, grid); }
  } dump{a, grid};
#endif
  switch (c.family) {
    case kNormal:
      return launch_t<kNormal>(tmap, tmap_dx, a, grid, threads, smem);
    case kBernoulli:
      return launch_t<kBernoulli>(tmap, tmap_dx, a, grid, threads, smem);
    case kPoisson:
      return launch_t<kPoisson>(tmap, tmap_dx, a, grid, threads, smem);
    case kNegBinomial:
      return launch_t<kNegBinomial>(tmap, tmap_dx, a, grid, threads, smem);
    case kOrdered:
      return launch_t<kOrdered>(tmap, tmap_dx, a, grid, threads, smem);
    case kBinomial:
      return launch_t<kBinomial>(tmap, tmap_dx, a, grid, threads, smem);
    case kLinear:
      return launch_t<kLinear>(tmap, tmap_dx, a, grid, threads, smem);
  }
  return fail(SMC_ERR_INVALID_ARGUMENT, "unknown family %d", c.family);
}

}  // namespace smc
