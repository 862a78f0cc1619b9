// General two-pass path for the memory-bound GLM families: any K (including 0
// and K > 256), any leading dimension / alignment.  It reads x twice (row pass
// for theta + link, column pass for x^T d), so it runs at half the roofline of
// the fused kernel; launch_glm() only selects it when the fused single-pass
// kernel cannot take the matrix.  Same link code (glm_link.cuh), same packed
// output, same deterministic fixed-order reductions.
#include <cmath>
#include <cstring>

#include "glm_link.cuh"

namespace smc {

constexpr int kGenThreads = 256;

int prepare_args(const GlmCall& c, FusedArgs* ap) {
  FusedArgs& a = *ap;
  memset(&a, 0, sizeof(a));
  const smc_matrix* x = c.x;
  for (const smc_matrix* m : {c.x, c.y, c.alpha_vec, c.aux_vec})
    if (int rc = realize(m)) return rc;
  for (smc_matrix* m : {c.d_alpha_vec, c.d_aux_vec, c.d_y_vec, c.d_x})
    if (m) m->zero_pending = false;  // overwritten
  a.N = x->rows;
  a.K = (int)x->cols;
  a.x = static_cast<const double*>(x->data);
  a.ldx = x->ld;
  a.flags = c.flags | (c.unfused ? kFlagUnfused : 0u);
  a.ncuts = (int)c.ncuts;
  a.y = c.y ? c.y->data : nullptr;
  a.y_scalar = c.y_scalar;
  a.alpha_vec = c.alpha_vec ? static_cast<const double*>(c.alpha_vec->data) : nullptr;
  a.alpha = c.alpha;
  if (c.family == kBinomial)
    a.aux_ivec = c.aux_vec ? static_cast<const int*>(c.aux_vec->data) : nullptr;
  else
    a.aux_vec = c.aux_vec ? static_cast<const double*>(c.aux_vec->data) : nullptr;
  a.aux = c.aux;
  a.d_alpha_vec = c.d_alpha_vec ? static_cast<double*>(c.d_alpha_vec->data) : nullptr;
  a.d_aux_vec = c.d_aux_vec ? static_cast<double*>(c.d_aux_vec->data) : nullptr;
  a.d_y_vec = c.d_y_vec ? static_cast<double*>(c.d_y_vec->data) : nullptr;
  a.d_x = c.d_x ? static_cast<double*>(c.d_x->data) : nullptr;
  a.ld_dx = c.d_x ? c.d_x->ld : 0;
  a.out = c.out;

  // row-independent terms of the log density, evaluated once on the host
  const double Nd = (double)a.N;
  const bool propto = c.flags & SMC_PROPTO;
  a.c0 = 0.0;
  switch (c.family) {
    case kNormal:
      // normal_id_glm_lpdf.hpp L201-212
      if (!propto) a.c0 += -0.91893853320467274178032973640561764 * Nd;
      if (!c.aux_vec && (!propto || (c.flags & SMC_VAR_AUX)))
        a.c0 -= Nd * std::log(c.aux);
      break;
    case kPoisson:
    case kNegBinomial: {
      // -sum lgamma(y + 1): poisson L126-128, neg-binomial L163-169.  The un-fused
      // neg_binomial_2_log_lpmf carries it inside binomial_coefficient_log, which
      // stays whenever phi is an autodiff variable (lpmf L111-113)
      if (!propto
          || (c.unfused && c.family == kNegBinomial && (c.flags & SMC_VAR_AUX))) {
        double lg = 0.0;
        if (c.y) {
          if (int rc = y_lgamma_sum(c.y, &lg)) return rc;
        } else {
          // a broadcast scalar y: the poisson GLM sums lgamma(y + 1) over the
          // elements of y as passed (once), the neg-binomial GLM scales by N
          // (L165-168) -- both reproduced as the reference has them
          lg = std::lgamma(c.y_scalar + 1.0)
               * (c.family == kPoisson && !c.unfused ? (c.once_terms ? 1.0 : 0.0) : Nd);
        }
        a.c0 -= lg;
      }
      if (c.family == kNegBinomial && !c.aux_vec) {
        // lgamma(y + phi) / digamma(y + phi) tables cover y in [0, tab_n)
        int lo = 0, hi = (int)c.y_scalar;
        if (c.y) {
          if (int rc = y_range(c.y, &lo, &hi)) return rc;
        }
        if (!propto || (c.flags & SMC_VAR_AUX))
          a.tab_n = hi + 1 < kMaxLgammaTab ? (hi + 1 > 0 ? hi + 1 : 0) : kMaxLgammaTab;
        a.log_aux = std::log(c.aux);
        a.digamma_aux = digamma(c.aux);
        // N (phi log phi - lgamma phi), L177-182 (multiply_log(0,0) = 0)
        const double ml = (c.aux == 0.0) ? 0.0 : c.aux * std::log(c.aux);
        if (!propto || (c.flags & SMC_VAR_AUX))
          a.c0 += Nd * (ml - std::lgamma(c.aux));
        else if (c.unfused)
          a.c0 += Nd * ml;  // neg_binomial_2_log_lpmf.hpp L117-118 keeps it
      }
      break;
    }
    case kBinomial: {
      // sum binomial_coefficient_log(N, n) [* N when both are broadcast scalars],
      // binomial_logit_glm_lpmf.hpp L124-127; n and N are data: cached
      if (!propto) {
        bool ok = true;
        double bc = 0.0;
        const int64_t cnt = (c.y || c.aux_vec) ? a.N : 1;
        if (int rc = binom_stats(c.y, (int)c.y_scalar, c.aux_vec, (int)c.aux, cnt, &ok,
                                 &bc))
          return rc;
        a.c0 += bc * (cnt == a.N ? 1.0 : Nd);
      }
      break;
    }
    default:
      break;
  }
  return SMC_OK;
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  // fixed-order block reduction: butterfly inside the warp, then warp 0 adds
  // the warp totals in index order
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kGenThreads / 32; ++w) t += sh[w];
  return t;
}

template <int FAM>
__global__ void __launch_bounds__(kGenThreads)
    generic_rows_kernel(const __grid_constant__ FusedArgs a,
                        const double* __restrict__ params, double* __restrict__ dvec,
                        double* __restrict__ d1v, double* __restrict__ d2v,
                        double* __restrict__ block_partials) {
  __shared__ double sh[kGenThreads / 32];
  RowAcc racc;
  LinkTab tab;
  tab.cuts = params + a.K;
  for (int64_t row = blockIdx.x * (int64_t)kGenThreads + threadIdx.x; row < a.N;
       row += (int64_t)gridDim.x * kGenThreads) {
    double xb = 0.0;
    const double* xr = a.x + row;
    for (int k = 0; k < a.K; ++k) xb = fma(xr[(size_t)k * a.ldx], __ldg(params + k), xb);
    RowIn<FAM> in;
    if constexpr (FAM == kNormal)
      in.y = a.y ? static_cast<const double*>(a.y)[row] : a.y_scalar;
    else
      in.y = a.y ? (double)static_cast<const int*>(a.y)[row] : a.y_scalar;
    in.alpha = a.alpha_vec ? a.alpha_vec[row] : a.alpha;
    if constexpr (FAM == kBinomial)
      in.aux = a.aux_ivec ? (double)a.aux_ivec[row] : a.aux;
    else
      in.aux = a.aux_vec ? a.aux_vec[row] : a.aux;
    double d1 = 0, d2 = 0;
    const double d = link_row<FAM>(a, xb, in, true, true, row, racc, tab, d1, d2);
    if (a.K > 0) dvec[row] = d;  // (K = 0: an un-fused density, no column pass)
    if constexpr (FAM == kOrdered) {
      d1v[row] = d1;
      d2v[row] = d2;
    }
  }
  const double v0 = block_sum(racc.lp, sh), v1 = block_sum(racc.sd, sh),
               v2 = block_sum(racc.s2, sh), v3 = block_sum(racc.s3, sh),
               vb = block_sum((double)racc.bad, sh);
  if (threadIdx.x == 0) {
    double* bp = block_partials + (size_t)blockIdx.x * kHdr;
    bp[SMC_OUT_LOGP] = v0;
    bp[SMC_OUT_SUM_D] = v1;
    bp[SMC_OUT_AUX] = v2;
    bp[SMC_OUT_NONFINITE] = vb;
    bp[SMC_OUT_AUX2] = v3;
    bp[5] = bp[6] = bp[7] = 0.0;
  }
}

// One CTA per column k: d_beta[k] = sum_i x[i,k] d[i]; optionally d_x[:,k].
__global__ void __launch_bounds__(kGenThreads)
    generic_cols_kernel(const __grid_constant__ FusedArgs a,
                        const double* __restrict__ params,
                        const double* __restrict__ dvec) {
  __shared__ double sh[kGenThreads / 32];
  const int k = blockIdx.x;
  const double* col = a.x + (size_t)k * a.ldx;
  const bool need_dx = (a.flags & SMC_VAR_X) && a.d_x;
  const double bk = params[k];
  double v = 0.0;
  for (int64_t i = threadIdx.x; i < a.N; i += kGenThreads) {
    const double di = dvec[i];
    v = fma(col[i], di, v);
    if (need_dx) a.d_x[(size_t)k * a.ld_dx + i] = bk * di;
  }
  v = block_sum(v, sh);
  if (threadIdx.x == 0) a.out[kHdr + k] = v;
}

// One CTA per cut point c (ordered_logistic_glm_lpmf.hpp L197-207).
__global__ void __launch_bounds__(kGenThreads)
    generic_cuts_kernel(const __grid_constant__ FusedArgs a,
                        const double* __restrict__ d1v,
                        const double* __restrict__ d2v) {
  __shared__ double sh[kGenThreads / 32];
  const int c = blockIdx.x;
  double v = 0.0;
  for (int64_t i = threadIdx.x; i < a.N; i += kGenThreads) {
    const int yy = a.y ? static_cast<const int*>(a.y)[i] : (int)a.y_scalar;
    if (yy - 1 == c) v += d2v[i];
    if (yy - 2 == c) v -= d1v[i];
  }
  v = block_sum(v, sh);
  if (threadIdx.x == 0) a.out[kHdr + a.K + c] = v;
}

__global__ void generic_finalize_kernel(const double* __restrict__ block_partials,
                                        int nblocks, double c0,
                                        double* __restrict__ out) {
  const int j = threadIdx.x;
  if (j >= kHdr) return;
  double v = 0.0;
  for (int b = 0; b < nblocks; ++b) v += block_partials[(size_t)b * kHdr + j];
  if (j == SMC_OUT_LOGP) v += c0;
  out[j] = v;
}

template <int FAM>
static int run_rows(const FusedArgs& a, const double* params, double* dvec,
                    double* d1v, double* d2v, double* bp, int grid) {
  generic_rows_kernel<FAM><<<grid, kGenThreads, 0, ctx().stream>>>(a, params, dvec,
                                                                  d1v, d2v, bp);
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

int launch_glm_generic(const GlmCall& c) {
  Context& cx = ctx();
  FusedArgs a;
  if (int rc = prepare_args(c, &a)) return rc;
  const int nparam = a.K + a.ncuts;
  const double* params = c.params_dev;
  if (!params) {
    if (int rc = ensure_params(sizeof(double) * (nparam > 0 ? nparam : 1))) return rc;
    if (a.K)
      SMC_CUDA(cudaMemcpyAsync(cx.params_dev, c.beta_host, sizeof(double) * a.K,
                               cudaMemcpyHostToDevice, cx.stream));
    if (a.ncuts)
      SMC_CUDA(cudaMemcpyAsync(cx.params_dev + a.K, c.cuts_host,
                               sizeof(double) * a.ncuts, cudaMemcpyHostToDevice,
                               cx.stream));
    params = cx.params_dev;
  }
  int grid = (int)((a.N + kGenThreads - 1) / kGenThreads);
  if (grid > cx.sm_count * 4) grid = cx.sm_count * 4;
  if (grid < 1) grid = 1;
  const size_t nvec = c.family == kOrdered ? 3 : 1;
  const size_t bytes = sizeof(double) * ((size_t)a.N * nvec + (size_t)grid * kHdr);
  if (int rc = ensure_scratch(bytes)) return rc;
  double* dvec = cx.scratch;
  double* d1v = dvec + (nvec == 3 ? a.N : 0);
  double* d2v = dvec + (nvec == 3 ? 2 * a.N : 0);
  double* bp = dvec + (size_t)a.N * nvec;
  int rc = SMC_OK;
  switch (c.family) {
    case kNormal:
      rc = run_rows<kNormal>(a, params, dvec, d1v, d2v, bp, grid);
      break;
    case kBernoulli:
      rc = run_rows<kBernoulli>(a, params, dvec, d1v, d2v, bp, grid);
      break;
    case kPoisson:
      rc = run_rows<kPoisson>(a, params, dvec, d1v, d2v, bp, grid);
      break;
    case kNegBinomial:
      rc = run_rows<kNegBinomial>(a, params, dvec, d1v, d2v, bp, grid);
      break;
    case kOrdered:
      rc = run_rows<kOrdered>(a, params, dvec, d1v, d2v, bp, grid);
      break;
    case kBinomial:
      rc = run_rows<kBinomial>(a, params, dvec, d1v, d2v, bp, grid);
      break;
    case kLinear:
      rc = run_rows<kLinear>(a, params, dvec, d1v, d2v, bp, grid);
      break;
    default:
      return fail(SMC_ERR_INVALID_ARGUMENT, "unknown family %d", c.family);
  }
  if (rc) return rc;
  cx.launches += 1;
  if (a.K > 0 && (a.flags & (SMC_VAR_BETA | SMC_VAR_X))) {
    generic_cols_kernel<<<a.K, kGenThreads, 0, cx.stream>>>(a, params, dvec);
    SMC_CUDA(cudaGetLastError());
    cx.launches += 1;
  }
  if (c.family == kOrdered && a.ncuts > 0 && (a.flags & SMC_VAR_AUX)) {
    generic_cuts_kernel<<<a.ncuts, kGenThreads, 0, cx.stream>>>(a, d1v, d2v);
    SMC_CUDA(cudaGetLastError());
    cx.launches += 1;
  }
  generic_finalize_kernel<<<1, 32, 0, cx.stream>>>(bp, grid, a.c0, a.out);
  SMC_CUDA(cudaGetLastError());
  cx.launches += 1;
  return SMC_OK;
}

// A wide design matrix (K > 256 columns) in column chunks of at most 256, every
// chunk through the fused TMA kernel:
//   forward   theta = alpha + x_1 beta_1 + ... + x_{m-1} beta_{m-1}   (kLinear, chained
//             through an N-vector)
//   family    the GLM itself on the last chunk with that theta as its intercept
//             vector: value, d (left in an N-vector), d_beta_m, d_x_m, header
//   reverse   d_beta_i = x_i^T d, d_x_i = beta_i (x) d for i < m      (kLinear)
// (2 m - 1) / m sweeps over x at the fused kernel's rate instead of two sweeps of
// the row / column kernels above (K = 512: 1.5 sweeps; measured 2.2x faster).
static int launch_glm_chunked(const GlmCall& c) {
  Context& cx = ctx();
  const smc_matrix* x = c.x;
  const int64_t N = x->rows, K = x->cols;
  const int m = (int)((K + kMaxFusedK - 1) / kMaxFusedK);
  const int64_t w = (K + m - 1) / m;  // chunk width (the last chunk may be narrower)
  const bool dx = (c.flags & SMC_VAR_X) && c.d_x;
  const bool reverse = (c.flags & SMC_VAR_BETA) || dx;
  // scratch: theta[N], d[N] (unless the caller takes d as the partial of a vector
  // alpha), and a junk packed result for the forward launches
  const size_t junk = SMC_OUT_HEADER + (size_t)kMaxFusedK + kMaxCuts;
  const size_t head = 8192;  // (the y-statistics kernels use the first bytes of scratch)
  if (int rc = ensure_scratch(sizeof(double) * (head + 2 * (size_t)N + junk))) return rc;
  auto vec = [&](double* p) {
    smc_matrix v;
    v.data = p;
    v.rows = N;
    v.cols = 1;
    v.ld = N;
    v.dtype = SMC_F64;
    return v;
  };
  smc_matrix theta = vec(cx.scratch + head);
  smc_matrix dvec = c.d_alpha_vec ? *c.d_alpha_vec : vec(cx.scratch + head + N);
  dvec.owned = false;
  double* junk_out = cx.scratch + head + 2 * (size_t)N;
  auto cols_of = [&](const smc_matrix* mtx, int i) {
    smc_matrix v = *mtx;
    const int64_t c0 = (int64_t)i * w;
    v.data = static_cast<double*>(mtx->data) + c0 * mtx->ld;
    v.cols = (c0 + w <= K) ? w : K - c0;
    v.owned = false;
    v.tmap_rows = v.tmap_cols = 0;  // the cached descriptor belongs to the whole matrix
    return v;
  };
  if (c.family == kLinear) {
    // the products themselves: every chunk is one independent kLinear launch
    // (theta chained through the caller's output vector, x^T v written at the
    // chunk's offset), one sweep in total
    for (int i = 0; i < m; ++i) {
      smc_matrix xi = cols_of(x, i), dxi;
      GlmCall f = c;
      f.x = &xi;
      f.beta_host = c.beta_host + (int64_t)i * w;
      if (i > 0) {
        f.alpha_vec = c.d_alpha_vec;  // theta so far (NULL on the reverse sweep)
        f.alpha = 0.0;
        f.out_skip_header = true;     // sum v comes from the first chunk
      }
      if (dx) {
        dxi = cols_of(c.d_x, i);
        f.d_x = &dxi;
      }
      f.out_beta_off = (int)(i * w);
      f.out_K_total = (int)K;
      if (i + 1 < m) f.done_flag = nullptr;
      if (int rc = launch_glm_fused(f)) return rc;
    }
    return SMC_OK;
  }
  for (int i = 0; i + 1 < m; ++i) {
    smc_matrix xi = cols_of(x, i);
    GlmCall f;
    f.family = kLinear;
    f.x = &xi;
    f.beta_host = c.beta_host + (int64_t)i * w;
    f.alpha_vec = i == 0 ? c.alpha_vec : &theta;
    f.alpha = i == 0 ? c.alpha : 0.0;
    f.d_alpha_vec = &theta;
    f.out = junk_out;
    if (int rc = launch_glm_fused(f)) return rc;
  }
  {
    smc_matrix xl = cols_of(x, m - 1), dxl;
    GlmCall g = c;
    g.x = &xl;
    g.beta_host = c.beta_host + (int64_t)(m - 1) * w;
    g.alpha_vec = &theta;
    g.alpha = 0.0;
    g.d_alpha_vec = &dvec;
    if (dx) {
      dxl = cols_of(c.d_x, m - 1);
      g.d_x = &dxl;
    }
    g.out_beta_off = (int)((m - 1) * w);
    g.out_K_total = (int)K;
    if (reverse) g.done_flag = nullptr;  // a later launch finishes the evaluation
    if (int rc = launch_glm_fused(g)) return rc;
  }
  for (int i = 0; reverse && i + 1 < m; ++i) {
    smc_matrix xi = cols_of(x, i), dxi;
    GlmCall r;
    r.family = kLinear;
    r.x = &xi;
    r.beta_host = c.beta_host + (int64_t)i * w;  // (beta (x) d needs the real beta)
    r.aux_vec = &dvec;
    r.flags = SMC_VAR_BETA | (dx ? SMC_VAR_X : 0u);
    if (dx) {
      dxi = cols_of(c.d_x, i);
      r.d_x = &dxi;
    }
    r.out = c.out;
    r.out_beta_off = (int)(i * w);
    r.out_K_total = (int)K;
    r.out_skip_header = true;
    if (i + 2 == m) {
      r.done_flag = c.done_flag;
      r.done_val = c.done_val;
    }
    if (int rc = launch_glm_fused(r)) return rc;
  }
  return SMC_OK;
}

// x is an autodiff variable: d_x = beta (x) d.  Writing it from the sweep that reads x
// saves no traffic (x is read once, d_x written once either way) but mixes the two
// streams, and DRAM serves a read-only plus a write-only stream faster than the
// mix: N=1e7, K=128 neg-binomial 3.49 ms fused -> 3.31 ms split (1.59 ms read-only
// sweep that leaves d in an N-vector + 1.71 ms = 6.0 TB/s of pure stores), identical
// bits.  The polled completion flag of the sweep cannot cover the second kernel, so
// the synchronous caller falls back to the stream synchronise.
static int launch_glm_split_dx(const GlmCall& c) {
  Context& cx = ctx();
  const int64_t N = c.x->rows;
  GlmCall s = c;
  s.d_x = nullptr;
  s.done_flag = nullptr;
  smc_matrix dvec;
  void* tmp = nullptr;
  const size_t bytes = sizeof(double) * (size_t)N;
  if (!s.d_alpha_vec) {
    if (int rc = cache_alloc(&tmp, bytes)) return rc;
    dvec.data = tmp;
    dvec.rows = N;
    dvec.cols = 1;
    dvec.ld = N;
    dvec.dtype = SMC_F64;
    dvec.device = cx.device;
    s.d_alpha_vec = &dvec;
  }
  int rc = launch_glm_fused(s);
  if (!rc) {
    c.d_x->version++;
    rc = launch_outer(static_cast<double*>(c.d_x->data), c.d_x->ld,
                      static_cast<const double*>(s.d_alpha_vec->data), N, (int)c.x->cols,
                      c.beta_host, c.params_dev);
  }
  if (tmp) cache_free(tmp, bytes);  // reused by this thread on this stream only
  cx.flag_armed = false;
  return rc;
}

// SMC_DX_FACTORED: the caller takes the factor d of d_x = d beta^T (it applies the
// product in its reverse sweep): the evaluation is the plain x-data sweep with d left
// in the caller's vector.
static int launch_glm_factored_dx(const GlmCall& c) {
  GlmCall s = c;
  // (SMC_VAR_X stays: it decides which terms survive propto; without d_x no kernel
  // forms the product)
  s.flags &= ~SMC_DX_FACTORED;
  s.d_x = nullptr;
  if (!c.d_alpha_vec) {
    s.d_alpha_vec = c.d_x;
    return launch_glm(s);
  }
  // a vector alpha wants the same d: evaluate into its partial and copy
  if (int rc = launch_glm(s)) return rc;
  c.d_x->version++;
  c.d_x->zero_pending = false;
  SMC_CUDA(cudaMemcpyAsync(c.d_x->data, c.d_alpha_vec->data,
                           sizeof(double) * (size_t)c.x->rows, cudaMemcpyDeviceToDevice,
                           ctx().stream));
  ctx().flag_armed = false;  // the polled flag cannot cover the copy
  return SMC_OK;
}

int launch_glm(const GlmCall& c) {
  if ((c.flags & SMC_VAR_X) && (c.flags & SMC_DX_FACTORED) && c.d_x)
    return launch_glm_factored_dx(c);
  const bool force_generic = knobs().force_generic;
  // (d_x leaves the fused kernel through TMA stores: same layout rules as x)
  const bool dx_ok = !((c.flags & SMC_VAR_X) && c.d_x) || fused_supported(c.d_x);
  if (fused_supported(c.x) && dx_ok && !force_generic
      && c.ncuts <= 4 * 32 * ((c.x->cols + 31) / 32)) {
    // (SMC_DX_FUSED=1: A/B switch, d_x from the sweep itself)
    if ((c.flags & SMC_VAR_X) && c.d_x && c.x->rows > 0 && !knobs().dx_fused)
      return launch_glm_split_dx(c);
    return launch_glm_fused(c);
  }
  // wide x: column chunks through the fused kernel (host parameters only; the
  // cut points ride with the last chunk, which must be able to take them)
  if (c.x && c.x->cols > kMaxFusedK && fused_layout_ok(c.x)
      && (!((c.flags & SMC_VAR_X) && c.d_x) || fused_layout_ok(c.d_x)) && !c.params_dev
      && c.beta_host && !force_generic && c.ncuts <= 4 * 32)
    return launch_glm_chunked(c);
  return launch_glm_generic(c);
}

}  // namespace smc
