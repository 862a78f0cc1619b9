/* oracle/glm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's six GLM log-density + gradient
 * functions (Stan Math, stan/math/prim/prob/ *_glm_*.hpp).  It exists only so
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can CHECK the
 * CUDA path; the product (math_b200/, include/) never links, imports or calls
 * it.
 *
 * PARITY: PINNED.  tests/test_oracle.py checks every function here against
 *   (a) the known answers of SURVEY.md 8(c) (reference test inputs),
 *   (b) tests/golden/ *.json produced by the UNMODIFIED reference
 *       (oracle/ref_driver.cpp -> oracle/_ref/libstan_ref.so,
 *        script tests/golden/make_golden.py), and
 *   (c) oracle/_ref/libstan_ref.so live, when it is present.
 *
 * Layout: x column-major N x K with leading dimension ldx (>= N); beta for
 * the categorical GLM is column-major K x C.  `ny`/`nalpha`/`naux` are 1
 * (scalar broadcast) or N.  Output pointers may be NULL.
 * flags: bit0 propto; bit1 x var; bit2 alpha var; bit3 beta var;
 *        bit4 sigma/phi/cuts var; bit5 y var (normal only).
 * They mirror the reference's compile-time include_summand<propto, ...> /
 * is_constant_all<...> switches.
 * Return: 0 ok, 1 size error (std::invalid_argument), 2 value error
 * (std::domain_error).  On error outputs are untouched.
 *
 * Summation order is NOT part of the contract (Eigen's differs too): long
 * sums use Neumaier compensation so the oracle is the more accurate side.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define F_PROPTO 1u
#define F_VAR_X 2u
#define F_VAR_ALPHA 4u
#define F_VAR_BETA 8u
#define F_VAR_AUX 16u
#define F_VAR_Y 32u

extern double lgamma_r(double, int*);

/* ----- compensated accumulator ------------------------------------------------ */
typedef struct {
  double s, c;
} acc_t;
static inline void acc_add(acc_t* a, double v) {
  double t = a->s + v;
  if (isfinite(t)) {
    if (fabs(a->s) >= fabs(v))
      a->c += (a->s - t) + v;
    else
      a->c += (v - t) + a->s;
  }
  a->s = t;
}
static inline double acc_get(const acc_t* a) {
  return isfinite(a->s) ? a->s + a->c : a->s;
}

/* ----- scalar kernels ----------------------------------------------------------- */

/* reference: prim/fun/lgamma.hpp L63-67 (glibc lgamma_r) */
double oracle_lgamma(double x) {
  int sign = 1;
  return lgamma_r(x, &sign);
}

/* reference: prim/fun/log1p_exp.hpp L45-52 */
double oracle_log1p_exp(double a) {
  if (a > 0.0) return a + log1p(exp(-a));
  return log1p(exp(a));
}

/* reference: prim/fun/log1m_exp.hpp L47-57 (log1m(x) = log1p(-x), log1m.hpp) */
double oracle_log1m_exp(double a) {
  if (a > 0) return NAN;
  if (a > -0.693147) return log(-expm1(a));
  return log1p(-exp(a));
}

/* reference: prim/fun/multiply_log.hpp L49-56 */
static double multiply_log(double a, double b) {
  if (b == 0.0 && a == 0.0) return 0.0;
  return a * log(b);
}

/* reference: prim/fun/digamma.hpp L47-49 -> boost::math::digamma, Boost 1.84.0
 * (vendored: lib/boost_1.84.0/boost/math/special_functions/digamma.hpp),
 * double precision (53-bit) branch with promote_double<false>:
 *   x <= -1: reflection  pi / tan(pi * rem)           (L404-421)
 *   x >= 10: asymptotic  log(x-1) + 1/(2(x-1)) - z P(z), z = 1/(x-1)^2  (L117-136)
 *   else   : recurrence into [1,2] (L439-452), then
 *            (x - root)(Y + P(x-1)/Q(x-1))             (L311-353)
 * Coefficients are the published Boost minimax / Bernoulli constants. */
double oracle_digamma(double x) {
  static const double PL[8] = {0.083333333333333333333333333333333333333,
                               -0.0083333333333333333333333333333333333333,
                               0.003968253968253968253968253968253968254,
                               -0.0041666666666666666666666666666666666667,
                               0.0075757575757575757575757575757575757576,
                               -0.021092796092796092796092796092796092796,
                               0.083333333333333333333333333333333333333,
                               -0.44325980392156862745098039215686274510};
  static const double P12[6]
      = {0.25479851061131551,   -0.32555031186804491,  -0.65031853770896507,
         -0.28919126444774784,  -0.045251321448739056, -0.0020713321167745952};
  static const double Q12[7]
      = {1.0,
         2.0767117023730469,
         1.4606242909763515,
         0.43593529692665969,
         0.054151797245674225,
         0.0021284987017821144,
         -0.55789841321675513e-6};
  const double Y = 0.99558162689208984f;
  const double root1 = 1569415565.0 / 1073741824.0;
  const double root2 = (381566830.0 / 1073741824.0) / 1073741824.0;
  const double root3 = 0.9016312093258695918615325266959189453125e-19;
  double result = 0.0;
  if (x <= -1) {
    x = 1 - x;
    double rem = x - floor(x);
    if (rem > 0.5) rem -= 1;
    if (rem == 0) return NAN; /* pole */
    result = M_PI / tan(M_PI * rem);
  }
  if (x == 0) return NAN; /* pole */
  if (x >= 10) {
    double xm = x - 1;
    double r = log(xm);
    r += 1 / (2 * xm);
    double z = 1 / (xm * xm);
    double p = PL[7];
    for (int i = 6; i >= 0; --i) p = p * z + PL[i];
    r -= z * p;
    return result + r;
  }
  while (x > 2) {
    x -= 1;
    result += 1 / x;
  }
  while (x < 1) {
    result -= 1 / x;
    x += 1;
  }
  double g = x - root1;
  g -= root2;
  g -= root3;
  double t = x - 1;
  double p = P12[5];
  for (int i = 4; i >= 0; --i) p = p * t + P12[i];
  double q = Q12[6];
  for (int i = 5; i >= 0; --i) q = q * t + Q12[i];
  double r = p / q;
  return result + (g * Y + g * r);
}

/* ----- shared helpers ----------------------------------------------------------- */

/* theta[i] = sum_k x[i,k] beta[k]  (column sweep, like a column-major GEMV) */
static void xbeta(long N, long K, const double* x, long ldx, const double* beta,
                  double* theta) {
  acc_t* a = (acc_t*)calloc((size_t)(N > 0 ? N : 1), sizeof(acc_t));
  for (long k = 0; k < K; ++k) {
    const double* col = x + (size_t)k * (size_t)ldx;
    const double b = beta[k];
    for (long i = 0; i < N; ++i) acc_add(&a[i], col[i] * b);
  }
  for (long i = 0; i < N; ++i) theta[i] = acc_get(&a[i]);
  free(a);
}

/* d_beta[k] = sum_i x[i,k] d[i];  d_x[i,k] = beta[k] d[i] */
static void xt_d(long N, long K, const double* x, long ldx, const double* d,
                 double* d_beta) {
  if (!d_beta) return;
  for (long k = 0; k < K; ++k) {
    const double* col = x + (size_t)k * (size_t)ldx;
    acc_t a = {0, 0};
    for (long i = 0; i < N; ++i) acc_add(&a, col[i] * d[i]);
    d_beta[k] = acc_get(&a);
  }
}
static void outer_bd(long N, long K, const double* beta, const double* d,
                     double* d_x) {
  if (!d_x) return;
  for (long k = 0; k < K; ++k)
    for (long i = 0; i < N; ++i) d_x[(size_t)k * (size_t)N + i] = beta[k] * d[i];
}
static int all_finite(const double* p, long n) {
  for (long i = 0; i < n; ++i)
    if (!isfinite(p[i])) return 0;
  return 1;
}
static double sum_vec(const double* p, long n) {
  acc_t a = {0, 0};
  for (long i = 0; i < n; ++i) acc_add(&a, p[i]);
  return acc_get(&a);
}
static int bad_len(long n, long N) { return !(n == 1 || n == N); }
#define BR(p, n, i) ((n) == 1 ? (p)[0] : (p)[i])

/* ----- bernoulli_logit_glm_lpmf --------------------------------------------------
 * reference: prim/prob/bernoulli_logit_glm_lpmf.hpp L49-167 */
int oracle_bernoulli_logit_glm(long N, long K, const int* y, long ny,
                               const double* x, long ldx, const double* alpha,
                               long nalpha, const double* beta, unsigned flags,
                               double* logp, double* d_alpha, double* d_beta,
                               double* d_x) {
  if (bad_len(ny, N) || bad_len(nalpha, N)) return 1; /* L76-79 */
  if (N == 0) {                                       /* size_zero(y) L80-82 */
    if (logp) *logp = 0;
    return 0;
  }
  for (long i = 0; i < (ny == 1 ? 1 : N); ++i) /* check_bounded L85 */
    if (y[i] < 0 || y[i] > 1) return 2;
  if ((flags & F_PROPTO)
      && !(flags & (F_VAR_X | F_VAR_ALPHA | F_VAR_BETA))) { /* L87-89 */
    if (logp) *logp = 0;
    return 0;
  }
  double* yt = (double*)malloc(sizeof(double) * (size_t)N);
  double* d = (double*)malloc(sizeof(double) * (size_t)N);
  xbeta(N, K, x, ldx, beta, yt);
  const double cutoff = 20.0; /* L120 */
  acc_t lp = {0, 0};
  for (long i = 0; i < N; ++i) {
    double s = 2.0 * BR(y, ny, i) - 1.0;           /* L105-106 */
    double t = s * (yt[i] + BR(alpha, nalpha, i)); /* L114-115 */
    double e = exp(-t);                            /* L121 */
    yt[i] = t;
    acc_add(&lp, t > cutoff ? -e : (t < -cutoff ? t : -log1p(e))); /* L122-126 */
    /* L137-142 -- NB the t > cutoff branch is -e whatever the sign */
    d[i] = t > cutoff ? -e : (t < -cutoff ? s : s * e / (e + 1));
  }
  double lpv = acc_get(&lp);
  if (!isfinite(lpv)) { /* L128-132 */
    if (!all_finite(beta, K) || !all_finite(alpha, nalpha)
        || !all_finite(yt, N)) {
      free(yt);
      free(d);
      return 2;
    }
  }
  if (logp) *logp = lpv;
  xt_d(N, K, x, ldx, d, d_beta); /* L149 */
  outer_bd(N, K, beta, d, d_x);  /* L158-159 */
  if (d_alpha) {                 /* L162-164: scalar edge sums the vector */
    if (nalpha == 1)
      d_alpha[0] = sum_vec(d, N);
    else
      memcpy(d_alpha, d, sizeof(double) * (size_t)N);
  }
  free(yt);
  free(d);
  return 0;
}

/* ----- poisson_log_glm_lpmf ------------------------------------------------------
 * reference: prim/prob/poisson_log_glm_lpmf.hpp L51-163 */
int oracle_poisson_log_glm(long N, long K, const int* y, long ny,
                           const double* x, long ldx, const double* alpha,
                           long nalpha, const double* beta, unsigned flags,
                           double* logp, double* d_alpha, double* d_beta,
                           double* d_x) {
  if (bad_len(ny, N) || bad_len(nalpha, N)) return 1; /* L79-82 */
  for (long i = 0; i < (ny == 1 ? 1 : N); ++i)        /* check_nonnegative L84 */
    if (y[i] < 0) return 2;
  if (N == 0
      || ((flags & F_PROPTO)
          && !(flags & (F_VAR_X | F_VAR_ALPHA | F_VAR_BETA)))) { /* L86-91 */
    if (logp) *logp = 0;
    return 0;
  }
  double* th = (double*)malloc(sizeof(double) * (size_t)N);
  double* d = (double*)malloc(sizeof(double) * (size_t)N);
  xbeta(N, K, x, ldx, beta, th);
  acc_t lp = {0, 0}, sd = {0, 0}, lg = {0, 0};
  for (long i = 0; i < N; ++i) {
    double yi = BR(y, ny, i);
    th[i] += BR(alpha, nalpha, i); /* L107-115 */
    double e = exp(th[i]);
    d[i] = yi - e; /* L117-118 */
    acc_add(&sd, d[i]);
    acc_add(&lp, yi * th[i] - e); /* L130-131 */
    /* L126-128: sum(lgamma(y + 1)) over the elements of y AS PASSED -- a
     * broadcast scalar y contributes the term once, not N times (reference
     * behaviour, reproduced for parity; the neg-binomial GLM does scale by N) */
    if (!(flags & F_PROPTO) && (ny != 1 || i == 0))
      acc_add(&lg, oracle_lgamma(yi + 1));
  }
  double sdv = acc_get(&sd);
  if (!isfinite(sdv)) { /* L120-124 */
    if (!all_finite(beta, K) || !all_finite(alpha, nalpha)
        || !all_finite(th, N)) {
      free(th);
      free(d);
      return 2;
    }
  }
  if (logp) *logp = acc_get(&lp) - acc_get(&lg);
  xt_d(N, K, x, ldx, d, d_beta); /* L142 */
  outer_bd(N, K, beta, d, d_x);  /* L151-152 */
  if (d_alpha) {                 /* L155-161 */
    if (nalpha == 1)
      d_alpha[0] = sdv;
    else
      memcpy(d_alpha, d, sizeof(double) * (size_t)N);
  }
  free(th);
  free(d);
  return 0;
}

/* ----- normal_id_glm_lpdf --------------------------------------------------------
 * reference: prim/prob/normal_id_glm_lpdf.hpp L54-216 */
int oracle_normal_id_glm(long N, long K, const double* y, long ny,
                         const double* x, long ldx, const double* alpha,
                         long nalpha, const double* beta, const double* sigma,
                         long nsigma, unsigned flags, double* logp,
                         double* d_alpha, double* d_beta, double* d_sigma,
                         double* d_x, double* d_y) {
  const double NEG_LOG_SQRT_TWO_PI = -0.91893853320467274178032973640561764;
  if (bad_len(ny, N) || bad_len(nalpha, N) || bad_len(nsigma, N))
    return 1;                                          /* L84-89 */
  for (long i = 0; i < (nsigma == 1 ? 1 : N); ++i)     /* L93 positive finite */
    if (!(sigma[i] > 0) || !isfinite(sigma[i])) return 2;
  if (N == 0
      || ((flags & F_PROPTO)
          && !(flags
               & (F_VAR_X | F_VAR_ALPHA | F_VAR_BETA | F_VAR_AUX | F_VAR_Y)))) {
    if (logp) *logp = 0; /* L95-100 */
    return 0;
  }
  double* r = (double*)malloc(sizeof(double) * (size_t)N);
  double* mu = (double*)malloc(sizeof(double) * (size_t)N);
  xbeta(N, K, x, ldx, beta, r);
  acc_t ss = {0, 0}, sl = {0, 0};
  for (long i = 0; i < N; ++i) {
    double inv = 1.0 / BR(sigma, nsigma, i); /* L116 */
    r[i] = (BR(y, ny, i) - r[i] - BR(alpha, nalpha, i)) * inv; /* L130-133 */
    mu[i] = inv * r[i];                                       /* L140 */
    acc_add(&ss, r[i] * r[i]);
    if (nsigma != 1) acc_add(&sl, log(sigma[i]));
  }
  double ssv = acc_get(&ss);
  if (!isfinite(ssv)) { /* L192-198: the last check always fails */
    free(r);
    free(mu);
    return 2;
  }
  double lp = 0; /* L201-213 */
  if (!(flags & F_PROPTO)) lp += NEG_LOG_SQRT_TWO_PI * (double)N;
  if (!(flags & F_PROPTO) || (flags & F_VAR_AUX)) {
    if (nsigma != 1)
      lp -= acc_get(&sl);
    else
      lp -= (double)N * log(sigma[0]);
  }
  lp -= 0.5 * ssv;
  if (logp) *logp = lp;
  xt_d(N, K, x, ldx, mu, d_beta); /* L158-166 */
  outer_bd(N, K, beta, mu, d_x);  /* L148-157 */
  double smu = sum_vec(mu, N);
  if (d_y) { /* L141-147 */
    if (ny == 1)
      d_y[0] = -smu;
    else
      for (long i = 0; i < N; ++i) d_y[i] = -mu[i];
  }
  if (d_alpha) { /* L167-173 */
    if (nalpha == 1)
      d_alpha[0] = smu;
    else
      memcpy(d_alpha, mu, sizeof(double) * (size_t)N);
  }
  if (d_sigma) { /* L174-184 */
    if (nsigma == 1)
      d_sigma[0] = (ssv - (double)N) * (1.0 / sigma[0]);
    else
      for (long i = 0; i < N; ++i)
        d_sigma[i] = (r[i] * r[i] - 1) * (1.0 / sigma[i]);
  }
  free(r);
  free(mu);
  return 0;
}

/* ----- neg_binomial_2_log_glm_lpmf -----------------------------------------------
 * reference: prim/prob/neg_binomial_2_log_glm_lpmf.hpp L64-248 */
int oracle_neg_binomial_2_log_glm(long N, long K, const int* y, long ny,
                                  const double* x, long ldx,
                                  const double* alpha, long nalpha,
                                  const double* beta, const double* phi,
                                  long nphi, unsigned flags, double* logp,
                                  double* d_alpha, double* d_beta,
                                  double* d_phi, double* d_x) {
  if (bad_len(ny, N) || bad_len(nalpha, N) || bad_len(nphi, N))
    return 1;                                                   /* L102-107 */
  if (!all_finite(beta, K) || !all_finite(alpha, nalpha)) return 2; /* L114-115 */
  if (N == 0) { /* L117-119 size_zero(y, phi) */
    if (logp) *logp = 0;
    return 0;
  }
  for (long i = 0; i < (ny == 1 ? 1 : N); ++i) /* L129 */
    if (y[i] < 0) return 2;
  for (long i = 0; i < (nphi == 1 ? 1 : N); ++i) /* L130 */
    if (!(phi[i] > 0) || !isfinite(phi[i])) return 2;
  const int any_var
      = (flags & (F_VAR_X | F_VAR_ALPHA | F_VAR_BETA | F_VAR_AUX)) != 0;
  if ((flags & F_PROPTO) && !any_var) { /* L132-134 */
    if (logp) *logp = 0;
    return 0;
  }
  double* th = (double*)malloc(sizeof(double) * (size_t)N);
  double* d = (double*)malloc(sizeof(double) * (size_t)N);
  xbeta(N, K, x, ldx, beta, th);
  for (long i = 0; i < N; ++i) th[i] += BR(alpha, nalpha, i); /* L143-151 */
  if (!all_finite(th, N)) {                                   /* L152 */
    free(th);
    free(d);
    return 2;
  }
  const int inc_const = !(flags & F_PROPTO);
  const int inc_phi = !(flags & F_PROPTO) || (flags & F_VAR_AUX);
  const int inc_lin
      = !(flags & F_PROPTO) || (flags & (F_VAR_X | F_VAR_ALPHA | F_VAR_BETA));
  acc_t lp = {0, 0}, dphi = {0, 0};
  for (long i = 0; i < N; ++i) {
    double yi = BR(y, ny, i), ph = BR(phi, nphi, i);
    double log_phi = log(ph); /* L153 */
    double lse = th[i] > log_phi ? th[i] + oracle_log1p_exp(log_phi - th[i])
                                 : log_phi + oracle_log1p_exp(th[i] - log_phi);
    double ypp = yi + ph; /* L159 */
    if (inc_const) acc_add(&lp, -oracle_lgamma(yi + 1.0)); /* L163-169 */
    if (inc_phi)
      acc_add(&lp, multiply_log(ph, ph) - oracle_lgamma(ph)); /* L170-183 */
    acc_add(&lp, -ypp * lse);                                 /* L184 */
    if (inc_lin) acc_add(&lp, yi * th[i]);                    /* L186-188 */
    if (inc_phi) acc_add(&lp, oracle_lgamma(ypp));            /* L189-195 */
    double te = exp(th[i]);                                   /* L201 */
    d[i] = yi - te * ypp / (te + ph);                         /* L203-204 */
    double dp = 1 - ypp / (te + ph) + log_phi - lse + oracle_digamma(ypp)
                - oracle_digamma(ph); /* L235-244 */
    if (nphi == 1)
      acc_add(&dphi, dp);
    else if (d_phi)
      d_phi[i] = dp;
  }
  if (logp) *logp = acc_get(&lp);
  xt_d(N, K, x, ldx, d, d_beta); /* L211-212 */
  outer_bd(N, K, beta, d, d_x);  /* L221-222 */
  if (d_alpha) {                 /* L225-231 */
    if (nalpha == 1)
      d_alpha[0] = sum_vec(d, N);
    else
      memcpy(d_alpha, d, sizeof(double) * (size_t)N);
  }
  if (d_phi && nphi == 1) d_phi[0] = acc_get(&dphi);
  free(th);
  free(d);
  return 0;
}

/* ----- ordered_logistic_glm_lpmf -------------------------------------------------
 * reference: prim/prob/ordered_logistic_glm_lpmf.hpp L46-210;  ncuts = C-1 */
int oracle_ordered_logistic_glm(long N, long K, const int* y, long ny,
                                const double* x, long ldx, const double* beta,
                                const double* cuts, long ncuts, unsigned flags,
                                double* logp, double* d_beta, double* d_cuts,
                                double* d_x) {
  const long C = ncuts + 1;
  if (bad_len(ny, N)) return 1;                /* L75-77 */
  for (long i = 0; i < (ny == 1 ? 1 : N); ++i) /* L82 */
    if (y[i] < 1 || y[i] > C) return 2;
  for (long c = 1; c < ncuts; ++c) /* check_ordered L83 */
    if (!(cuts[c] > cuts[c - 1])) return 2;
  if (ncuts == 1 && isnan(cuts[0])) return 2;
  if (C > 1) { /* L84-89 */
    if (C > 2 && !isfinite(cuts[C - 2])) return 2;
    if (!isfinite(cuts[0])) return 2;
  }
  if (N == 0 || ncuts == 0
      || ((flags & F_PROPTO)
          && !(flags & (F_VAR_X | F_VAR_BETA | F_VAR_AUX)))) { /* L91-96 */
    if (logp) *logp = 0;
    return 0;
  }
  double* loc = (double*)malloc(sizeof(double) * (size_t)N);
  double* d = (double*)malloc(sizeof(double) * (size_t)N);
  xbeta(N, K, x, ldx, beta, loc); /* L123 */
  if (!isfinite(sum_vec(loc, N))) { /* L124-127 */
    int ok = all_finite(beta, K);
    for (long k = 0; ok && k < K; ++k)
      ok = all_finite(x + (size_t)k * (size_t)ldx, N);
    if (!ok) {
      free(loc);
      free(d);
      return 2;
    }
  }
  acc_t lp = {0, 0};
  acc_t* dc = (acc_t*)calloc((size_t)ncuts, sizeof(acc_t));
  for (long i = 0; i < N; ++i) {
    int c = BR(y, ny, i);
    double c1 = c != C ? cuts[c - 1] : INFINITY;  /* L108-121 */
    double c2 = c != 1 ? cuts[c - 2] : -INFINITY;
    double cut2 = loc[i] - c2, cut1 = loc[i] - c1; /* L129-132 */
    double A = (cut1 > 0.0 ? -cut1 : 0) - log1p(exp(-fabs(cut1)));  /* L135-136 */
    double B = (cut2 <= 0.0 ? cut2 : 0) - log1p(exp(-fabs(cut2)));  /* L137-138 */
    double term; /* L141-161 */
    if (c == 1)
      term = A;
    else if (c == C)
      term = B;
    else
      term = B + oracle_log1m_exp(cut1 - cut2) + A;
    acc_add(&lp, term);
    double em1 = exp(-cut1), em2 = exp(-cut2), ed = exp(c2 - c1); /* L165-167 */
    double d1 = (cut2 > 0 ? em2 / (1 + em2) : 1 / (1 + exp(cut2)))
                - ed / (ed - 1); /* L168-170 */
    double d2 = 1 / (1 - ed)
                - (cut1 > 0 ? em1 / (1 + em1) : 1 / (1 + exp(cut1))); /* L171-174 */
    d[i] = d1 - d2; /* L176 */
    if (c != C) acc_add(&dc[c - 1], d2);  /* L197-207 */
    if (c != 1) acc_add(&dc[c - 2], -d1);
  }
  if (logp) *logp = acc_get(&lp);
  xt_d(N, K, x, ldx, d, d_beta); /* L192-193 */
  outer_bd(N, K, beta, d, d_x);  /* L182-183 */
  if (d_cuts)
    for (long c = 0; c < ncuts; ++c) d_cuts[c] = acc_get(&dc[c]);
  free(dc);
  free(loc);
  free(d);
  return 0;
}

/* ----- categorical_logit_glm_lpmf ------------------------------------------------
 * reference: prim/prob/categorical_logit_glm_lpmf.hpp L43-195
 * beta column-major K x C; alpha C; d_beta K x C; d_x N x K (ld N). */
int oracle_categorical_logit_glm(long N, long K, long C, const int* y, long ny,
                                 const double* x, long ldx, const double* alpha,
                                 const double* beta, unsigned flags,
                                 double* logp, double* d_alpha, double* d_beta,
                                 double* d_x) {
  if (bad_len(ny, N)) return 1; /* L68-72 */
  if (N == 0 || C == 1) {       /* L73-75 */
    if (logp) *logp = 0;
    return 0;
  }
  for (long i = 0; i < (ny == 1 ? 1 : N); ++i) /* L77 */
    if (y[i] < 1 || y[i] > C) return 2;
  if ((flags & F_PROPTO)
      && !(flags & (F_VAR_X | F_VAR_ALPHA | F_VAR_BETA))) { /* L80-82 */
    if (logp) *logp = 0;
    return 0;
  }
  /* lin = x beta + alpha^T  (N x C, column-major)  L95-96 */
  double* lin = (double*)malloc(sizeof(double) * (size_t)N * (size_t)C);
  double* inv = (double*)malloc(sizeof(double) * (size_t)N);
  for (long c = 0; c < C; ++c) {
    xbeta(N, K, x, ldx, beta + (size_t)c * (size_t)K, lin + (size_t)c * (size_t)N);
    for (long i = 0; i < N; ++i) lin[(size_t)c * N + i] += alpha[c];
  }
  acc_t lp = {0, 0};
  for (long i = 0; i < N; ++i) {
    double m = -INFINITY; /* L97-98 */
    for (long c = 0; c < C; ++c)
      if (lin[(size_t)c * N + i] > m || isnan(lin[(size_t)c * N + i]))
        m = lin[(size_t)c * N + i];
    double yl = lin[(size_t)(BR(y, ny, i) - 1) * N + i];
    double s = 0;
    for (long c = 0; c < C; ++c) { /* L101-102: exp_lin overwrites lin */
      double e = exp(lin[(size_t)c * N + i] - m);
      lin[(size_t)c * N + i] = e;
      s += e;
    }
    inv[i] = 1 / s;                         /* L103-104 */
    acc_add(&lp, log(inv[i]) - m);          /* L106 */
    acc_add(&lp, yl);                       /* L110-116 */
  }
  double lpv = acc_get(&lp);
  if (!isfinite(lpv)) { /* L122-126 */
    int ok = all_finite(beta, K * C) && all_finite(alpha, C);
    for (long k = 0; ok && k < K; ++k)
      ok = all_finite(x + (size_t)k * (size_t)ldx, N);
    if (!ok) {
      free(lin);
      free(inv);
      return 2;
    }
  }
  if (logp) *logp = lpv;
  if (d_x) { /* L142-150: beta[:, y_i-1] - (exp_lin beta^T) inv */
    for (long k = 0; k < K; ++k)
      for (long i = 0; i < N; ++i) {
        acc_t a = {0, 0};
        for (long c = 0; c < C; ++c)
          acc_add(&a, lin[(size_t)c * N + i] * beta[(size_t)c * K + k]);
        d_x[(size_t)k * N + i]
            = beta[(size_t)(BR(y, ny, i) - 1) * K + k] - acc_get(&a) * inv[i];
      }
  }
  /* neg_softmax_lin = exp_lin * -inv  L158-159 (in place) */
  for (long c = 0; c < C; ++c)
    for (long i = 0; i < N; ++i) lin[(size_t)c * N + i] *= -inv[i];
  if (d_alpha) { /* L160-169 */
    for (long c = 0; c < C; ++c) d_alpha[c] = sum_vec(lin + (size_t)c * N, N);
    for (long i = 0; i < N; ++i) d_alpha[BR(y, ny, i) - 1] += 1;
  }
  if (d_beta) { /* L171-190 */
    for (long c = 0; c < C; ++c)
      xt_d(N, K, x, ldx, lin + (size_t)c * N, d_beta + (size_t)c * K);
    for (long k = 0; k < K; ++k) {
      /* scatter: column (y_i - 1) += x.row(i), compensated per class */
      acc_t* a = (acc_t*)calloc((size_t)C, sizeof(acc_t));
      for (long i = 0; i < N; ++i)
        acc_add(&a[BR(y, ny, i) - 1], x[(size_t)k * (size_t)ldx + i]);
      for (long c = 0; c < C; ++c) d_beta[(size_t)c * K + k] += acc_get(&a[c]);
      free(a);
    }
  }
  free(lin);
  free(inv);
  return 0;
}

/* ----- categorical_logit_lpmf, one row of log odds per outcome ----------------------
 * reference: prim/prob/categorical_logit_lpmf.hpp L16-32 applied to every row of
 * the column-major N x C matrix lin (what a model writes as a loop):
 *   check_bounded(n, 1, C) L19; check_finite(beta) L22; propto exit L24-26;
 *   value beta[n-1] - log_sum_exp(beta) L30-31, log_sum_exp = max + log(sum(exp(v -
 *   max))) (prim/fun/log_sum_exp.hpp L81-93); its reverse sweep adds
 *   exp(v - log_sum_exp(v)) = softmax(v) (rev/fun/log_sum_exp.hpp), so
 *   d lin[i, c] = [c == y_i - 1] - softmax(lin[i, :])[c].
 * flags: bit0 propto, bit2 lin var. */
int oracle_categorical_logit_lpmf(long N, long C, const int* y, long ny,
                                  const double* lin, long ld, unsigned flags,
                                  double* logp, double* d_lin) {
  if (bad_len(ny, N)) return 1;
  if (logp) *logp = 0;
  if (N == 0) return 0;
  for (long i = 0; i < (ny == 1 ? 1 : N); ++i)
    if (y[i] < 1 || y[i] > C) return 2;
  for (long c = 0; c < C; ++c)
    for (long i = 0; i < N; ++i)
      if (!isfinite(lin[c * ld + i])) return 2;
  if ((flags & F_PROPTO) && !(flags & F_VAR_ALPHA)) return 0;
  acc_t lp = {0, 0};
  for (long i = 0; i < N; ++i) {
    const long yi = (ny == 1 ? y[0] : y[i]) - 1;
    double m = -INFINITY, s = 0;
    for (long c = 0; c < C; ++c) m = fmax(m, lin[c * ld + i]);
    for (long c = 0; c < C; ++c) s += exp(lin[c * ld + i] - m);
    const double lse = m + log(s);
    acc_add(&lp, lin[yi * ld + i] - lse);
    if (d_lin && (flags & F_VAR_ALPHA))
      for (long c = 0; c < C; ++c)
        d_lin[c * N + i] = (c == yi ? 1.0 : 0.0) - exp(lin[c * ld + i] - lse);
  }
  if (logp) *logp = acc_get(&lp);
  return 0;
}

/* ----- binomial_logit_glm_lpmf ---------------------------------------------------
 * (the seventh GLM)
 * reference: prim/prob/binomial_logit_glm_lpmf.hpp L54-160 and the scalar
 * kernels it uses: prim/fun/log_inv_logit.hpp L52-58, log1m_inv_logit.hpp L44-50,
 * binomial_coefficient_log.hpp L79-141, lbeta.hpp L64-118,
 * lgamma_stirling_diff.hpp L44-76, lgamma_stirling.hpp L27-29. */

/* log_inv_logit.hpp L52-58 */
double oracle_log_inv_logit(double u) {
  if (u < 0.0) return u - oracle_log1p_exp(u);
  return -oracle_log1p_exp(-u);
}
/* log1m_inv_logit.hpp L44-50 */
double oracle_log1m_inv_logit(double u) {
  if (u > 0.0) return -u - oracle_log1p_exp(-u);
  return -oracle_log1p_exp(u);
}

#define HALF_LOG_TWO_PI 0.91893853320467274178032973640561764
#define STIRLING_DIFF_USEFUL 10.0

/* lgamma_stirling_diff.hpp L44-76 (x >= 0) */
static double lgamma_stirling_diff(double x) {
  static const double series[6]
      = {0.0833333333333333333333333,   -0.00277777777777777777777778,
         0.000793650793650793650793651, -0.000595238095238095238095238,
         0.000841750841750841750841751, -0.00191752691752691752691753};
  if (isnan(x)) return NAN;
  if (x == 0) return INFINITY;
  if (x < STIRLING_DIFF_USEFUL) /* lgamma(x) - lgamma_stirling(x) */
    return oracle_lgamma(x) - (HALF_LOG_TWO_PI + (x - 0.5) * log(x) - x);
  double result = 0.0;
  double multiplier = 1.0 / x;
  const double inv_x_squared = multiplier * multiplier;
  for (int n = 0; n < 6; ++n) {
    if (n > 0) multiplier *= inv_x_squared;
    result += series[n] * multiplier;
  }
  return result;
}

/* lbeta.hpp L64-118 (a, b >= 0) */
static double lbeta(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  double x, y; /* x is the smaller of the two */
  if (a < b) {
    x = a;
    y = b;
  } else {
    x = b;
    y = a;
  }
  if (x == 0) return INFINITY;
  if (isinf(y)) return -INFINITY;
  if (y < STIRLING_DIFF_USEFUL) /* both small */
    return oracle_lgamma(x) + oracle_lgamma(y) - oracle_lgamma(x + y);
  const double x_over_xy = x / (x + y);
  if (x < STIRLING_DIFF_USEFUL) { /* y large, x small */
    const double sd = lgamma_stirling_diff(y) - lgamma_stirling_diff(x + y);
    const double st = (y - 0.5) * log1p(-x_over_xy) + x * (1 - log(x + y));
    return st + oracle_lgamma(x) + sd;
  }
  /* both large */
  const double sd = lgamma_stirling_diff(x) + lgamma_stirling_diff(y)
                    - lgamma_stirling_diff(x + y);
  const double st = (x - 0.5) * log(x_over_xy) + y * log1p(-x_over_xy)
                    + HALF_LOG_TWO_PI - 0.5 * log(y);
  return st + sd;
}

/* binomial_coefficient_log.hpp L79-115, value only; the GLM has already
 * checked 0 <= k <= n, so the function's own domain checks cannot fire */
double oracle_binomial_coefficient_log(double n, double k) {
  if (isnan(n) || isnan(k)) return NAN;
  if (n > -1 && k > n / 2.0 + 1e-8) /* the more stable symmetric branch */
    return oracle_binomial_coefficient_log(n, n - k);
  const double n_plus_1 = n + 1;
  const double n_plus_1_mk = n_plus_1 - k;
  if (k == 0) return 0;
  if (n_plus_1 < STIRLING_DIFF_USEFUL)
    return oracle_lgamma(n_plus_1) - oracle_lgamma(k + 1)
           - oracle_lgamma(n_plus_1_mk);
  return -lbeta(n_plus_1_mk, k + 1) - log1p(n);
}

/* n: successes (nn = 1 or N), Nt: trials (nNt = 1 or N) */
int oracle_binomial_logit_glm(long N, long K, const int* n, long nn,
                              const int* Nt, long nNt, const double* x, long ldx,
                              const double* alpha, long nalpha,
                              const double* beta, unsigned flags, double* logp,
                              double* d_alpha, double* d_beta, double* d_x) {
  /* size_zero(n, N, alpha, beta, x) L76-78: tested BEFORE the size checks */
  if (N == 0 || K == 0 || nn == 0 || nNt == 0 || nalpha == 0) {
    if (logp) *logp = 0;
    return 0;
  }
  if ((flags & F_PROPTO)
      && !(flags & (F_VAR_X | F_VAR_ALPHA | F_VAR_BETA))) { /* L80-82 */
    if (logp) *logp = 0;
    return 0;
  }
  if ((nn != 1 && nNt != 1 && nn != nNt) || bad_len(nn, N) || bad_len(nNt, N)
      || bad_len(nalpha, N))
    return 1; /* L88-93 */
  const long nb = nn > nNt ? nn : nNt;
  for (long i = 0; i < nb; ++i) /* check_bounded(n, 0, N) L98 */
    if (BR(n, nn, i) < 0 || BR(n, nn, i) > BR(Nt, nNt, i)) return 2;
  for (long i = 0; i < nNt; ++i) /* check_nonnegative(N) L99 */
    if (Nt[i] < 0) return 2;
  double* th = (double*)malloc(sizeof(double) * (size_t)N);
  double* d = (double*)malloc(sizeof(double) * (size_t)N);
  xbeta(N, K, x, ldx, beta, th);
  acc_t lp = {0, 0};
  for (long i = 0; i < N; ++i) {
    const double ni = BR(n, nn, i), Ni = BR(Nt, nNt, i);
    th[i] += BR(alpha, nalpha, i); /* L104-109 */
    const double lil = oracle_log_inv_logit(th[i]); /* L112 */
    acc_add(&lp, ni * lil + (Ni - ni) * oracle_log1m_inv_logit(th[i])); /* L114-115 */
    d[i] = ni - Ni * exp(lil); /* L131-132 */
  }
  double lpv = acc_get(&lp);
  if (!isfinite(lpv)) { /* L118-122: beta, alpha, x itself */
    int ok = all_finite(beta, K) && all_finite(alpha, nalpha);
    for (long k = 0; ok && k < K; ++k)
      ok = all_finite(x + (size_t)k * (size_t)ldx, N);
    if (!ok) {
      free(th);
      free(d);
      return 2;
    }
  }
  if (!(flags & F_PROPTO)) { /* include_summand<propto, T_n, T_N> L124-127 */
    acc_t bc = {0, 0};
    for (long i = 0; i < nb; ++i)
      acc_add(&bc, oracle_binomial_coefficient_log(BR(Nt, nNt, i), BR(n, nn, i)));
    const double broadcast_n = nb == N ? 1.0 : (double)N;
    lpv += acc_get(&bc) * broadcast_n;
  }
  if (logp) *logp = lpv;
  xt_d(N, K, x, ldx, d, d_beta); /* L139 */
  outer_bd(N, K, beta, d, d_x);  /* L148-150 */
  if (d_alpha) {                 /* L152-154 */
    if (nalpha == 1)
      d_alpha[0] = sum_vec(d, N);
    else
      memcpy(d_alpha, d, sizeof(double) * (size_t)N);
  }
  free(th);
  free(d);
  return 0;
}

/* ------------------------------------------------------------------ synthetic inputs
 * Host replica of the counter-based generator behind smc_matrix_fill_synthetic
 * (include/stanmath_cuda.h): splitmix-style hash of (seed, global row, column), exact
 * integer arithmetic and one correctly rounded multiply, so host and device hold the
 * same bits.  Column-major with leading dimension ld.  This is how the CPU arms of
 * bench.py get the GPU arm's inputs at full size without a 20 GB transfer; the
 * threads only spread the row range over the host cores. */
#include <pthread.h>
#include <stdint.h>
#include <unistd.h>

static inline uint64_t synth_hash(uint64_t seed, uint64_t row, uint64_t col) {
  uint64_t z = seed + row * 0x9E3779B97F4A7C15ull + col * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

typedef struct {
  void* out;
  long ld, r_lo, r_hi, cols, row0;
  uint64_t seed;
  int kind, lo, hi;
  double c;
} synth_job;

static void* synth_worker(void* arg) {
  const synth_job* j = (const synth_job*)arg;
  const uint64_t span = (uint64_t)((int64_t)j->hi - (int64_t)j->lo + 1);
  for (long k = 0; k < j->cols; ++k)
    for (long r = j->r_lo; r < j->r_hi; ++r) {
      const uint64_t z = synth_hash(j->seed, (uint64_t)(j->row0 + r), (uint64_t)k);
      if (j->kind == 0) {
        const uint32_t u = (uint32_t)(z & 0xffff) + (uint32_t)((z >> 16) & 0xffff)
                           + (uint32_t)((z >> 32) & 0xffff) + (uint32_t)(z >> 48);
        ((double*)j->out)[k * j->ld + r] = ((double)u - 131070.0) * j->c;
      } else {
        ((int*)j->out)[k * j->ld + r] = (int)((int64_t)j->lo + (int64_t)(z % span));
      }
    }
  return NULL;
}

static void synth_run(synth_job base, long rows) {
  long nt = sysconf(_SC_NPROCESSORS_ONLN);
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if (rows * base.cols < 1000000) nt = 1;
  pthread_t th[64];
  synth_job jobs[64];
  for (long t = 0; t < nt; ++t) {
    jobs[t] = base;
    jobs[t].r_lo = rows * t / nt;
    jobs[t].r_hi = rows * (t + 1) / nt;
    if (nt == 1)
      synth_worker(&jobs[t]);
    else
      pthread_create(&th[t], NULL, synth_worker, &jobs[t]);
  }
  if (nt > 1)
    for (long t = 0; t < nt; ++t) pthread_join(th[t], NULL);
}

void oracle_synthetic_f64(double* out, long ld, long rows, long cols, uint64_t seed,
                          long row0, double scale) {
  synth_job j = {out, ld, 0, 0, cols, row0, seed, 0, 0, 0,
                 scale / sqrt(4294967295.0 / 3.0)};
  synth_run(j, rows);
}

void oracle_synthetic_i32(int* out, long ld, long rows, long cols, uint64_t seed, long row0,
                          int lo, int hi) {
  synth_job j = {out, ld, 0, 0, cols, row0, seed, 1, lo, hi, 0.0};
  synth_run(j, rows);
}
