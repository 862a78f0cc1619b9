// Link functions of the five memory-bound GLM families, shared by the fused
// single-pass kernel (glm_fused.cu) and the general two-pass path
// (glm_generic.cu).  Each follows the reference's branches line by line:
//   normal_id_glm_lpdf.hpp L122-213, bernoulli_logit_glm_lpmf.hpp L105-164,
//   poisson_log_glm_lpmf.hpp L107-161, neg_binomial_2_log_glm_lpmf.hpp L143-244,
//   ordered_logistic_glm_lpmf.hpp L108-207, binomial_logit_glm_lpmf.hpp L104-154.
#pragma once
#include "device_math.cuh"
#include "smc_internal.h"

namespace smc {

constexpr int kHdr = SMC_OUT_HEADER;

struct FusedArgs {
#ifdef SMC_FUSED_TRACE
  unsigned long long* trace;  // profiling build: per-CTA phase stamps (glm_fused.cu)
#endif
  int64_t N;
  int K, S, G, ntiles;
  const double* x;  // general path only (the fused kernel goes through TMA)
  int64_t ldx;
  unsigned flags;
  int ncuts;
  const void* y;
  double y_scalar;
  const double* alpha_vec;
  double alpha;
  const double* aux_vec;
  const int* aux_ivec;  // binomial: per-row number of trials (else `aux`)
  double aux, log_aux, digamma_aux;
  const double* params_dev;  // beta[K] then cuts[ncuts]; NULL -> inline_params
  double* d_alpha_vec;
  double* d_aux_vec;
  double* d_y_vec;
  double* d_x;
  int64_t ld_dx;
  double* partials;
  int pstride;
  unsigned int* counter;
  double* out;
  // host-visible completion flag (pinned, mapped): the last CTA stores done_val
  // after the packed result, so a synchronous caller can poll instead of paying
  // the wake-up latency of a stream synchronise; NULL = not used
  unsigned long long* done_flag;
  unsigned long long done_val;
  // column-chunked evaluation of a wide x (glm_generic.cu, launch_glm_chunked): this
  // launch covers columns [out_beta_off, out_beta_off + K) of out_K_total and may
  // have to leave the header of the packed result alone
  int out_beta_off, out_K_total, out_skip_header;
  double c0;
  int tab_n;  // neg-binomial, scalar phi: rows with y < tab_n read lgamma / digamma
              // of (y + phi) from a per-CTA table instead of evaluating them
  double inline_params[kMaxParamDoubles];
};

// Per-CTA lookup tables that take row-independent transcendentals out of the
// per-row link (values are produced by the very same expressions, so results
// are bit-identical to evaluating them per row -- except l1m, see below):
//   ordered   cls[4 c' .. 4 c' + 3], c' = class - 1:  c1, c2 (the two cut points
//             of the class, +-inf at the ends), ed / (ed - 1), 1 / (1 - ed) with
//             ed = exp(c2 - c1)          (ordered_logistic_glm_lpmf.hpp L108-121, L165-174)
//             cls[4 C + 2 c' .. + 1]:  l1m = log1m_exp(c2 - c1) and the bound `lim`
//             on |cut1| under which a row may use it.  prim evaluates
//             log1m_exp(cut1 - cut2) per row (L160) with cut1 - cut2 =
//             (loc - c1) - (loc - c2): mathematically c2 - c1 for every row of the
//             class, numerically off by up to 2 ulp(loc).  The table value is the
//             exact-argument one; a row takes it only while that rounding cannot
//             move the term by more than 5e-13 (|loc| <= 1e3 (c1 - c2), enforced
//             through |cut1| <= lim = 1e3 (c1 - c2) - |c1|), else it evaluates the
//             term per row as prim does.  Saves two of the six transcendentals of
//             an interior-class row.
//   neg-binomial (scalar phi)  lg[y] = lgamma(y + phi), dg[y] = digamma(y + phi)
//             for integer y < tab_n     (neg_binomial_2_log_glm_lpmf.hpp L189-195, L235-244)
// Null pointers mean "evaluate per row" (the general two-pass path).
struct LinkTab {
  const double* cuts = nullptr;
  const double* cls = nullptr;
  const double* lg = nullptr;
  const double* dg = nullptr;
  int tab_n = 0;
};

constexpr int kMaxLgammaTab = 512;

__device__ __forceinline__ void ordered_class_entry(const double* cuts, int ncuts,
                                                    int c, double* e) {
  const int C = ncuts + 1;
  const double c1 = c != C ? cuts[c - 1] : __longlong_as_double(0x7ff0000000000000ll);
  const double c2 = c != 1 ? cuts[c - 2] : __longlong_as_double(0xfff0000000000000ll);
  const double ed = exp(c2 - c1);
  e[0] = c1;
  e[1] = c2;
  e[2] = ed / (ed - 1.0);
  e[3] = 1.0 / (1.0 - ed);
}

// l1m entry of one class (see LinkTab): e = the class's c1, c2 from
// ordered_class_entry; out[0] = log1m_exp(c2 - c1), out[1] = lim.
__device__ __forceinline__ void ordered_class_l1m(const double* e, double* out) {
  const double c1 = e[0], c2 = e[1];
  const bool interior = isfinite(c1) && isfinite(c2);
  out[0] = interior ? log1m_exp(c2 - c1) : 0.0;
  out[1] = interior ? 1e3 * (c1 - c2) - fabs(c1) : -1.0;  // end classes: term unused
}

// The first and last class of the ordered link carry +-inf cut points
// (ordered_logistic_glm_lpmf.hpp L111-120), so with end classes in the data
// nearly every warp holds a lane whose exp(-|cut|) is exactly 0 -- and a zero
// numerator sends the WHOLE warp through the out-of-line slow path of the FP64
// division, directly (0 / (1 + 0)) and inside log1p (0 / (2 + 0)); ncu showed
// 15 % of the kernel's samples in that subroutine, called by 98 % of the warps.
// These helpers divide / take log1p of a harmless 1.0 on such lanes and select
// the exact result (0) afterwards: same bits, inline path only.  The empty asm
// keeps the compiler from folding the select back into the division.
__device__ __forceinline__ double opaque(double v) {
  asm volatile("" : "+d"(v));
  return v;
}
__device__ __forceinline__ double div_or_zero(double num, double den) {
  const bool z = num == 0.0;
  const double q = opaque(z ? 1.0 : num) / den;
  return z ? 0.0 : q;
}
__device__ __forceinline__ double log1p_or_zero(double e) {
  return log1p_nonneg(e);  // no division inside, exact 0 at e == 0
}

// Fills everything in `a` that does not depend on the kernel variant: shapes,
// pointers, flags, host-side constants of the log density (c0, log phi, ...).
int prepare_args(const GlmCall& c, FusedArgs* a);

// --------------------------------------------------------------- link functions
// Per-row result of a link: theta-derivative d, log-density term lp, two aux
// sums and the non-finite flag.  `lead` marks the one warp per row group that
// owns the scalar sums and the N-vector outputs.
struct RowAcc {
  double lp = 0, sd = 0, s2 = 0, s3 = 0;
  int bad = 0;
};

template <int FAM>
struct RowIn {
  double y;      // response (int families: exact small integer in a double)
  double alpha;  // intercept for this row
  double aux;    // sigma / phi for this row
};

// What the deferred log-density part of a row needs from the derivative part.
template <int FAM>
struct LinkStash {
  double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
  int64_t row = 0;
  int c = 0;
  bool valid = false;
};

// Derivative part of the link of one row: returns d = d logp_i / d theta_i, writes
// the per-row (N-vector) partials, accumulates the sums that depend on d only, and
// leaves in `st` what link_lp() needs.  Every warp of a row group runs this (all
// of them need d for their columns of d_beta); only the row's `lead` warp writes
// and accumulates.
template <int FAM>
__device__ __forceinline__ double link_d(const FusedArgs& a, double xb,
                                         const RowIn<FAM>& in, bool valid, bool lead,
                                         int64_t row, RowAcc& acc, const LinkTab& tab,
                                         double& d1o, double& d2o,
                                         LinkStash<FAM>& st) {
  double d = 0, lp = 0, s3 = 0;
  int bad = 0;
  st.valid = valid;
  st.row = row;
  if constexpr (FAM == kBernoulli) {
    const double sgn = 2.0 * in.y - 1.0;
    const double t = sgn * (xb + in.alpha);
    const double e = exp(-t);
    // bernoulli_logit_glm_lpmf.hpp L137-142 (the t > 20 derivative branch is -e
    // whatever the sign: reproduced for parity)
    d = t > 20.0 ? -e : (t < -20.0 ? sgn : sgn * e / (e + 1.0));
    bad = !isfinite(t);
    st.v0 = t;
    st.v1 = e;
    if (lead && valid && a.d_alpha_vec) a.d_alpha_vec[row] = d;
  } else if constexpr (FAM == kBinomial) {
    // binomial_logit_glm_lpmf.hpp L104-115, L131-132.  exp(-|theta|) serves both
    // log_inv_logit (log_inv_logit.hpp L52-58) and log1m_inv_logit
    // (log1m_inv_logit.hpp L44-50): their log1p_exp arguments are -|theta|
    const double th = xb + in.alpha;
    const double l = log1p_nonneg(exp(-fabs(th)));
    const double lil = th < 0.0 ? th - l : -l;
    d = in.y - in.aux * exp(lil);
    bad = !isfinite(th);
    st.v0 = th;
    st.v1 = l;
    st.v2 = in.y;
    st.v3 = in.aux;
    if (lead && valid && a.d_alpha_vec) a.d_alpha_vec[row] = d;
  } else if constexpr (FAM == kPoisson) {
    const double th = xb + in.alpha;
    const double e = exp(th);
    d = in.y - e;        // poisson_log_glm_lpmf.hpp L117-118
    lp = in.y * th - e;  // L130-131
    bad = !isfinite(th);
    if (lead && valid && a.d_alpha_vec) a.d_alpha_vec[row] = d;
  } else if constexpr (FAM == kNormal) {
    const double inv = 1.0 / in.aux;
    const double r = (in.y - xb - in.alpha) * inv;  // normal_id_glm_lpdf.hpp L130-133
    d = inv * r;                                    // mu_derivative L140
    const double r2 = r * r;
    lp = -0.5 * r2;
    if (a.aux_vec && (!(a.flags & SMC_PROPTO) || (a.flags & SMC_VAR_AUX)))
      lp -= log(in.aux);  // L204-206
    s3 = r2;
    // check_positive_finite(sigma) L93 for a vector sigma is folded into the sweep
    if (a.aux_vec) bad = !(in.aux > 0.0) || !isfinite(in.aux);
    if (lead && valid) {
      if (a.d_alpha_vec) a.d_alpha_vec[row] = d;
      if (a.d_y_vec) a.d_y_vec[row] = -d;
      if (a.d_aux_vec) a.d_aux_vec[row] = (r2 - 1.0) * inv;  // L177-178
    }
  } else if constexpr (FAM == kNegBinomial) {
    const double th = xb + in.alpha;
    const double ph = in.aux;
    // check_finite(theta) L152; check_positive_finite(phi) L130 for vector phi
    bad = !isfinite(th) || (a.aux_vec && (!(ph > 0.0) || !isfinite(ph)));
    const double ypp = in.y + ph;
    const double te = exp(th);
    d = in.y - te * ypp / (te + ph);  // L203-204
    st.v0 = th;
    st.v1 = te;
    st.v2 = in.y;
    st.v3 = ph;
    if (lead && valid && a.d_alpha_vec) a.d_alpha_vec[row] = d;
  } else if constexpr (FAM == kOrdered) {
    const int c = valid ? (int)in.y : 1;
    // ordered_logistic_glm_lpmf.hpp L108-121 and the class-only factors of L165-174
    double ce[4];
    if (tab.cls) {
      const double2 w0 = *reinterpret_cast<const double2*>(tab.cls + 4 * (c - 1));
      const double2 w1 = *reinterpret_cast<const double2*>(tab.cls + 4 * (c - 1) + 2);
      ce[0] = w0.x;
      ce[1] = w0.y;
      ce[2] = w1.x;
      ce[3] = w1.y;
    } else {
      ordered_class_entry(tab.cuts, a.ncuts, c, ce);
    }
    // (alpha is 0 for the GLM; the un-fused ordered_logistic_lpmf passes its
    // location vector through it)
    const double loc = xb + in.alpha;
    const double cut2 = loc - ce[1], cut1 = loc - ce[0];  // L129-132
    // exp(-|cut|) serves both the stable log1p_exp forms (L135-138) and the
    // stable inv_logit selects (L168-174): for cut > 0 it IS exp(-cut), for
    // cut <= 0 it IS exp(cut) -- same argument, same bits
    const double e1 = exp_nonpos(-fabs(cut1)), e2 = exp_nonpos(-fabs(cut2));
    // (one division per select: the numerator is chosen, the quotient is the same)
    const double d1 = div_or_zero(cut2 > 0.0 ? e2 : 1.0, 1.0 + e2) - ce[2];
    const double d2 = ce[3] - div_or_zero(cut1 > 0.0 ? e1 : 1.0, 1.0 + e1);
    d = d1 - d2;
    d1o = d1;
    d2o = d2;
    st.v0 = cut1;
    st.v1 = cut2;
    st.v2 = e1;
    st.v3 = e2;
    st.c = c;
    s3 = loc;  // sum(location) for the lazy finiteness check, L124
    if (lead && valid && a.d_alpha_vec) a.d_alpha_vec[row] = d;
  } else if constexpr (FAM == kLinear) {
    // theta = x beta + alpha goes out per row; d = v_i, so that the column
    // reduction of the same sweep yields x^T v (and sum_i v_i)
    d = in.aux;
    if (lead && valid && a.d_alpha_vec) a.d_alpha_vec[row] = xb + in.alpha;
  }
  if (!valid) return 0.0;
  if (lead) {
    acc.lp += lp;
    acc.sd += d;
    acc.s3 += s3;
    acc.bad += bad;
  }
  return d;
}

// Log-density part of the link of one row (plus d_phi of the neg-binomial): the
// transcendental-heavy terms that only feed sums owned by the row's lead warp.
// It runs after the derivative part, possibly later (the fused kernel defers it
// past the next tile's partial dot product so that no other warp waits for it).
template <int FAM>
__device__ __forceinline__ void link_lp(const FusedArgs& a, const LinkStash<FAM>& st,
                                        RowAcc& acc, const LinkTab& tab) {
  if (!st.valid) return;
  if constexpr (FAM == kBernoulli) {
    const double t = st.v0, e = st.v1;
    acc.lp += t > 20.0 ? -e : (t < -20.0 ? t : -log1p_nonneg(e));  // L120-126
  } else if constexpr (FAM == kBinomial) {
    const double th = st.v0, l = st.v1, n = st.v2, nt = st.v3;
    const double lil = th < 0.0 ? th - l : -l;
    const double l1m = th > 0.0 ? -th - l : -l;
    acc.lp += n * lil + (nt - n) * l1m;  // L114-115
  } else if constexpr (FAM == kNegBinomial) {
    const double th = st.v0, te = st.v1, y = st.v2, ph = st.v3;
    const double ypp = y + ph;
    const bool propto = a.flags & SMC_PROPTO;
    const bool inc_phi = !propto || (a.flags & SMC_VAR_AUX);
    const bool inc_lin
        = !propto || (a.flags & (SMC_VAR_X | SMC_VAR_ALPHA | SMC_VAR_BETA));
    const double log_phi = a.aux_vec ? log(ph) : a.log_aux;
    // neg_binomial_2_log_glm_lpmf.hpp L154-157
    const double lse = th > log_phi ? th + log1p_exp(log_phi - th)
                                    : log_phi + log1p_exp(th - log_phi);
    double lp = -ypp * lse;     // L184
    if (inc_lin) lp += y * th;  // L186-188
    const int yi = (int)y;
    const bool in_tab = tab.lg != nullptr && yi < tab.tab_n;
    if (inc_phi) {
      lp += in_tab ? tab.lg[yi] : lgamma(ypp);  // L189-195
      if (a.aux_vec) lp += multiply_log(ph, ph) - lgamma(ph);  // L171-176
    } else if (a.aux_vec && (a.flags & kFlagUnfused)) {
      // neg_binomial_2_log_lpmf.hpp L117-118 keeps phi log(phi) under propto
      lp += multiply_log(ph, ph);
    }
    acc.lp += lp;
    if (a.flags & SMC_VAR_AUX) {
      const double dg_phi = a.aux_vec ? digamma(ph) : a.digamma_aux;
      const double dg_ypp = in_tab ? tab.dg[yi] : digamma(ypp);
      const double dp
          = 1.0 - ypp / (te + ph) + log_phi - lse + dg_ypp - dg_phi;  // L235-244
      if (a.d_aux_vec)
        a.d_aux_vec[st.row] = dp;
      else
        acc.s2 += dp;
    }
  } else if constexpr (FAM == kOrdered) {
    const double cut1 = st.v0, cut2 = st.v1, e1 = st.v2, e2 = st.v3;
    const int C = a.ncuts + 1;
    const double A = (cut1 > 0.0 ? -cut1 : 0.0) - log1p_or_zero(e1);
    const double B = (cut2 <= 0.0 ? cut2 : 0.0) - log1p_or_zero(e2);
    if (st.c == 1)
      acc.lp += A;
    else if (st.c == C)
      acc.lp += B;
    else {
      double l1m;  // L160: a class constant up to the rounding of cut1 - cut2
      bool per_row = true;
      if (tab.cls) {
        const double2 w = *reinterpret_cast<const double2*>(tab.cls + 4 * C + 2 * (st.c - 1));
        per_row = !(fabs(cut1) <= w.y);
        l1m = w.x;
      }
      if (per_row) l1m = log1m_exp(cut1 - cut2);
      acc.lp += B + l1m + A;  // L141-161
    }
  }
}

// Both parts back to back (the general two-pass path).
template <int FAM>
__device__ __forceinline__ double link_row(const FusedArgs& a, double xb,
                                           const RowIn<FAM>& in, bool valid,
                                           bool lead, int64_t row, RowAcc& acc,
                                           const LinkTab& tab, double& d1o,
                                           double& d2o) {
  LinkStash<FAM> st;
  const double d = link_d<FAM>(a, xb, in, valid, lead, row, acc, tab, d1o, d2o, st);
  if (lead) link_lp<FAM>(a, st, acc, tab);
  return d;
}

}  // namespace smc
