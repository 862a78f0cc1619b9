// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin extern "C" shim over the UNMODIFIED reference (Stan Math at
// /root/reference, header-only).  It is compiled by oracle/Makefile against
// the reference headers where they lie and produces oracle/_ref/libstan_ref.so
// (and libstan_ref_mt.so with -DSTAN_THREADS, which adds the reduce_sum
// entry points).  Nothing from the reference is copied: every function below
// only *calls* stan::math::<family>_glm_l[pd|pm]f with reverse-mode `var`
// parameters, runs grad() and hands value + adjoints back through plain
// pointers.
//
// Used for: (1) pinning the C restatement in oracle/glm_oracle.c, (2)
// generating tests/golden/*.json (tests/golden/make_golden.py), (3) the
// "reference" CPU baseline of bench.py (--impl reference / cpu_baseline).
//
// Conventions: x is column-major N x K (Eigen default).  `ny`, `nalpha`,
// `nsigma`, `nphi` are 1 (scalar, broadcast) or N (vector).  All real
// parameters (alpha, beta, sigma/phi/cuts) are `var`; x (and, for normal,
// y) are `var` when `data_var` != 0.  Output pointers may be NULL.
// Return: 0 ok, 1 std::invalid_argument, 2 std::domain_error, 3 other.
#include <stan/math.hpp>

#include <chrono>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace {

using stan::math::var;
using Eigen::Dynamic;
using Eigen::Map;
using Eigen::Matrix;
using Eigen::MatrixXd;
using Eigen::VectorXd;
using VecV = Matrix<var, Dynamic, 1>;
using MatV = Matrix<var, Dynamic, Dynamic>;
using VecI = std::vector<int>;

template <typename F>
int guarded(F&& f) {
  int rc = 0;
  try {
    f();
  } catch (const std::invalid_argument&) {
    rc = 1;
  } catch (const std::domain_error&) {
    rc = 2;
  } catch (...) {
    rc = 3;
  }
  stan::math::recover_memory();
  return rc;
}

inline void put(double* dst, const var& v) {
  if (dst) *dst = v.adj();
}
template <typename M>
inline void put_mat(double* dst, const M& m) {
  if (!dst) return;
  for (Eigen::Index j = 0; j < m.cols(); ++j)
    for (Eigen::Index i = 0; i < m.rows(); ++i)
      dst[j * m.rows() + i] = m(i, j).adj();
}

// A scalar-or-vector real argument, as data or var.
struct RealArg {
  const double* p;
  long n;
};

// Invoke `body(y, x, a, s)` with the right static types for scalar/vector and
// data/var operands.  Kept as nested generic lambdas so every family shares it.
template <typename T>
auto make_vec(const double* p, long n) {
  Matrix<T, Dynamic, 1> v(n);
  for (long i = 0; i < n; ++i) v(i) = p[i];
  return v;
}
template <typename T>
auto make_mat(const double* p, long r, long c) {
  Matrix<T, Dynamic, Dynamic> m(r, c);
  for (long j = 0; j < c; ++j)
    for (long i = 0; i < r; ++i) m(i, j) = p[j * r + i];
  return m;
}

// Dispatch helper: calls f(std::integral_constant<bool, B>{}).
template <typename F>
void with_bool(bool b, F&& f) {
  if (b)
    f(std::true_type{});
  else
    f(std::false_type{});
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- bernoulli
// reference: stan/math/prim/prob/bernoulli_logit_glm_lpmf.hpp L49-167
int ref_bernoulli_logit_glm(long N, long K, const int* y, long ny,
                            const double* x, const double* alpha, long nalpha,
                            const double* beta, int propto, int data_var,
                            double* logp, double* d_alpha, double* d_beta,
                            double* d_x) {
  return guarded([&] {
    with_bool(propto, [&](auto P) {
      with_bool(data_var, [&](auto XV) {
        with_bool(nalpha != 1, [&](auto AV) {
          with_bool(ny != 1, [&](auto YV) {
            using TX = std::conditional_t<decltype(XV)::value, var, double>;
            auto xm = make_mat<TX>(x, N, K);
            VecV b = make_vec<var>(beta, K);
            auto run = [&](const auto& yy, const auto& aa) {
              var lp = stan::math::bernoulli_logit_glm_lpmf<decltype(P)::value>(
                  yy, xm, aa, b);
              lp.grad();
              if (logp) *logp = lp.val();
              put_mat(d_beta, b);
              if constexpr (decltype(XV)::value) put_mat(d_x, xm);
              if constexpr (decltype(AV)::value)
                put_mat(d_alpha, aa);
              else
                put(d_alpha, aa);
            };
            auto with_alpha = [&](const auto& yy) {
              if constexpr (decltype(AV)::value) {
                VecV a = make_vec<var>(alpha, nalpha);
                run(yy, a);
              } else {
                var a = alpha[0];
                run(yy, a);
              }
            };
            if constexpr (decltype(YV)::value) {
              VecI yy(y, y + ny);
              with_alpha(yy);
            } else {
              int yy = y[0];
              with_alpha(yy);
            }
          });
        });
      });
    });
  });
}

// ------------------------------------------------------------------ poisson
// reference: stan/math/prim/prob/poisson_log_glm_lpmf.hpp L51-163
int ref_poisson_log_glm(long N, long K, const int* y, long ny, const double* x,
                        const double* alpha, long nalpha, const double* beta,
                        int propto, int data_var, double* logp, double* d_alpha,
                        double* d_beta, double* d_x) {
  return guarded([&] {
    with_bool(propto, [&](auto P) {
      with_bool(data_var, [&](auto XV) {
        with_bool(nalpha != 1, [&](auto AV) {
          with_bool(ny != 1, [&](auto YV) {
            using TX = std::conditional_t<decltype(XV)::value, var, double>;
            auto xm = make_mat<TX>(x, N, K);
            VecV b = make_vec<var>(beta, K);
            auto run = [&](const auto& yy, const auto& aa) {
              var lp = stan::math::poisson_log_glm_lpmf<decltype(P)::value>(
                  yy, xm, aa, b);
              lp.grad();
              if (logp) *logp = lp.val();
              put_mat(d_beta, b);
              if constexpr (decltype(XV)::value) put_mat(d_x, xm);
              if constexpr (decltype(AV)::value)
                put_mat(d_alpha, aa);
              else
                put(d_alpha, aa);
            };
            auto with_alpha = [&](const auto& yy) {
              if constexpr (decltype(AV)::value) {
                VecV a = make_vec<var>(alpha, nalpha);
                run(yy, a);
              } else {
                var a = alpha[0];
                run(yy, a);
              }
            };
            if constexpr (decltype(YV)::value) {
              VecI yy(y, y + ny);
              with_alpha(yy);
            } else {
              int yy = y[0];
              with_alpha(yy);
            }
          });
        });
      });
    });
  });
}

// ------------------------------------------------------------------- normal
// reference: stan/math/prim/prob/normal_id_glm_lpdf.hpp L54-216
// data_var makes BOTH x and y var (d_y, d_x returned).
int ref_normal_id_glm(long N, long K, const double* y, long ny, const double* x,
                      const double* alpha, long nalpha, const double* beta,
                      const double* sigma, long nsigma, int propto,
                      int data_var, double* logp, double* d_alpha,
                      double* d_beta, double* d_sigma, double* d_x,
                      double* d_y) {
  return guarded([&] {
    with_bool(propto, [&](auto P) {
      with_bool(data_var, [&](auto XV) {
        with_bool(nalpha != 1, [&](auto AV) {
          with_bool(nsigma != 1, [&](auto SV) {
            with_bool(ny != 1, [&](auto YV) {
              using TX = std::conditional_t<decltype(XV)::value, var, double>;
              auto xm = make_mat<TX>(x, N, K);
              VecV b = make_vec<var>(beta, K);
              auto run = [&](auto& yy, const auto& aa, const auto& ss) {
                var lp = stan::math::normal_id_glm_lpdf<decltype(P)::value>(
                    yy, xm, aa, b, ss);
                lp.grad();
                if (logp) *logp = lp.val();
                put_mat(d_beta, b);
                if constexpr (decltype(XV)::value) {
                  put_mat(d_x, xm);
                  if constexpr (decltype(YV)::value)
                    put_mat(d_y, yy);
                  else
                    put(d_y, yy);
                }
                if constexpr (decltype(AV)::value)
                  put_mat(d_alpha, aa);
                else
                  put(d_alpha, aa);
                if constexpr (decltype(SV)::value)
                  put_mat(d_sigma, ss);
                else
                  put(d_sigma, ss);
              };
              auto with_sigma = [&](auto& yy, const auto& aa) {
                if constexpr (decltype(SV)::value) {
                  VecV s = make_vec<var>(sigma, nsigma);
                  run(yy, aa, s);
                } else {
                  var s = sigma[0];
                  run(yy, aa, s);
                }
              };
              auto with_alpha = [&](auto& yy) {
                if constexpr (decltype(AV)::value) {
                  VecV a = make_vec<var>(alpha, nalpha);
                  with_sigma(yy, a);
                } else {
                  var a = alpha[0];
                  with_sigma(yy, a);
                }
              };
              if constexpr (decltype(YV)::value) {
                auto yy = make_vec<TX>(y, ny);
                with_alpha(yy);
              } else {
                TX yy = y[0];
                with_alpha(yy);
              }
            });
          });
        });
      });
    });
  });
}

// ----------------------------------------------------------------- neg-binomial
// reference: stan/math/prim/prob/neg_binomial_2_log_glm_lpmf.hpp L64-248
int ref_neg_binomial_2_log_glm(long N, long K, const int* y, long ny,
                               const double* x, const double* alpha,
                               long nalpha, const double* beta,
                               const double* phi, long nphi, int propto,
                               int data_var, double* logp, double* d_alpha,
                               double* d_beta, double* d_phi, double* d_x) {
  return guarded([&] {
    with_bool(propto, [&](auto P) {
      with_bool(data_var, [&](auto XV) {
        with_bool(nalpha != 1, [&](auto AV) {
          with_bool(nphi != 1, [&](auto SV) {
            with_bool(ny != 1, [&](auto YV) {
              using TX = std::conditional_t<decltype(XV)::value, var, double>;
              auto xm = make_mat<TX>(x, N, K);
              VecV b = make_vec<var>(beta, K);
              auto run = [&](const auto& yy, const auto& aa, const auto& ss) {
                var lp
                    = stan::math::neg_binomial_2_log_glm_lpmf<decltype(P)::value>(
                        yy, xm, aa, b, ss);
                lp.grad();
                if (logp) *logp = lp.val();
                put_mat(d_beta, b);
                if constexpr (decltype(XV)::value) put_mat(d_x, xm);
                if constexpr (decltype(AV)::value)
                  put_mat(d_alpha, aa);
                else
                  put(d_alpha, aa);
                if constexpr (decltype(SV)::value)
                  put_mat(d_phi, ss);
                else
                  put(d_phi, ss);
              };
              auto with_phi = [&](const auto& yy, const auto& aa) {
                if constexpr (decltype(SV)::value) {
                  VecV s = make_vec<var>(phi, nphi);
                  run(yy, aa, s);
                } else {
                  var s = phi[0];
                  run(yy, aa, s);
                }
              };
              auto with_alpha = [&](const auto& yy) {
                if constexpr (decltype(AV)::value) {
                  VecV a = make_vec<var>(alpha, nalpha);
                  with_phi(yy, a);
                } else {
                  var a = alpha[0];
                  with_phi(yy, a);
                }
              };
              if constexpr (decltype(YV)::value) {
                VecI yy(y, y + ny);
                with_alpha(yy);
              } else {
                int yy = y[0];
                with_alpha(yy);
              }
            });
          });
        });
      });
    });
  });
}

// ------------------------------------------------------------------ ordered
// reference: stan/math/prim/prob/ordered_logistic_glm_lpmf.hpp L46-210
// ncuts = C-1.
int ref_ordered_logistic_glm(long N, long K, const int* y, long ny,
                             const double* x, const double* beta,
                             const double* cuts, long ncuts, int propto,
                             int data_var, double* logp, double* d_beta,
                             double* d_cuts, double* d_x) {
  return guarded([&] {
    with_bool(propto, [&](auto P) {
      with_bool(data_var, [&](auto XV) {
        with_bool(ny != 1, [&](auto YV) {
          using TX = std::conditional_t<decltype(XV)::value, var, double>;
          auto xm = make_mat<TX>(x, N, K);
          VecV b = make_vec<var>(beta, K);
          VecV c = make_vec<var>(cuts, ncuts);
          auto run = [&](const auto& yy) {
            var lp = stan::math::ordered_logistic_glm_lpmf<decltype(P)::value>(
                yy, xm, b, c);
            lp.grad();
            if (logp) *logp = lp.val();
            put_mat(d_beta, b);
            put_mat(d_cuts, c);
            if constexpr (decltype(XV)::value) put_mat(d_x, xm);
          };
          if constexpr (decltype(YV)::value) {
            VecI yy(y, y + ny);
            run(yy);
          } else {
            int yy = y[0];
            run(yy);
          }
        });
      });
    });
  });
}

// -------------------------------------------------------------- categorical
// reference: stan/math/prim/prob/categorical_logit_glm_lpmf.hpp L43-195
// beta is column-major K x C, alpha has C entries.
int ref_categorical_logit_glm(long N, long K, long C, const int* y, long ny,
                              const double* x, const double* alpha,
                              const double* beta, int propto, int data_var,
                              double* logp, double* d_alpha, double* d_beta,
                              double* d_x) {
  return guarded([&] {
    with_bool(propto, [&](auto P) {
      with_bool(data_var, [&](auto XV) {
        with_bool(ny != 1, [&](auto YV) {
          using TX = std::conditional_t<decltype(XV)::value, var, double>;
          auto xm = make_mat<TX>(x, N, K);
          MatV b = make_mat<var>(beta, K, C);
          VecV a = make_vec<var>(alpha, C);
          auto run = [&](const auto& yy) {
            var lp = stan::math::categorical_logit_glm_lpmf<decltype(P)::value>(
                yy, xm, a, b);
            lp.grad();
            if (logp) *logp = lp.val();
            put_mat(d_beta, b);
            put_mat(d_alpha, a);
            if constexpr (decltype(XV)::value) put_mat(d_x, xm);
          };
          if constexpr (decltype(YV)::value) {
            VecI yy(y, y + ny);
            run(yy);
          } else {
            int yy = y[0];
            run(yy);
          }
        });
      });
    });
  });
}

// ------------------------------------------ categorical_logit_lpmf, row-wise
// reference: stan/math/prim/prob/categorical_logit_lpmf.hpp L16-32, called once
// per row of the column-major N x C matrix of log odds (the loop a model writes).
int ref_categorical_logit_lpmf(long N, long C, const int* y, long ny,
                               const double* lin, int propto, double* logp,
                               double* d_lin) {
  return guarded([&] {
    with_bool(propto, [&](auto P) {
      MatV l = make_mat<var>(lin, N, C);
      var lp = 0.0;
      for (long i = 0; i < N; ++i) {
        VecV row = l.row(i).transpose();
        lp += stan::math::categorical_logit_lpmf<decltype(P)::value>(
            ny == 1 ? y[0] : y[i], row);
      }
      lp.grad();
      if (logp) *logp = lp.val();
      put_mat(d_lin, l);
    });
  });
}

// ----------------------------------------------------------------- binomial
// reference: stan/math/prim/prob/binomial_logit_glm_lpmf.hpp L54-160
// n: successes (nn = 1 or N), Nt: trials (nNt = 1 or N).
int ref_binomial_logit_glm(long N, long K, const int* n, long nn, const int* Nt,
                           long nNt, const double* x, const double* alpha,
                           long nalpha, const double* beta, int propto,
                           int data_var, double* logp, double* d_alpha,
                           double* d_beta, double* d_x) {
  return guarded([&] {
    with_bool(propto, [&](auto P) {
      with_bool(data_var, [&](auto XV) {
        with_bool(nalpha != 1, [&](auto AV) {
          with_bool(nn != 1, [&](auto YV) {
            with_bool(nNt != 1, [&](auto TV) {
              using TX = std::conditional_t<decltype(XV)::value, var, double>;
              auto xm = make_mat<TX>(x, N, K);
              VecV b = make_vec<var>(beta, K);
              auto run = [&](const auto& ns, const auto& Ns, const auto& aa) {
                var lp = stan::math::binomial_logit_glm_lpmf<decltype(P)::value>(
                    ns, Ns, xm, aa, b);
                lp.grad();
                if (logp) *logp = lp.val();
                put_mat(d_beta, b);
                if constexpr (decltype(XV)::value) put_mat(d_x, xm);
                if constexpr (decltype(AV)::value)
                  put_mat(d_alpha, aa);
                else
                  put(d_alpha, aa);
              };
              auto with_alpha = [&](const auto& ns, const auto& Ns) {
                if constexpr (decltype(AV)::value) {
                  VecV a = make_vec<var>(alpha, nalpha);
                  run(ns, Ns, a);
                } else {
                  var a = alpha[0];
                  run(ns, Ns, a);
                }
              };
              auto with_trials = [&](const auto& ns) {
                if constexpr (decltype(TV)::value) {
                  VecI Ns(Nt, Nt + nNt);
                  with_alpha(ns, Ns);
                } else {
                  int Ns = Nt[0];
                  with_alpha(ns, Ns);
                }
              };
              if constexpr (decltype(YV)::value) {
                VecI ns(n, n + nn);
                with_trials(ns);
              } else {
                int ns = n[0];
                with_trials(ns);
              }
            });
          });
        });
      });
    });
  });
}

// ------------------------------------------------------------ scalar helpers
// reference: prim/fun/digamma.hpp L47-49, lgamma.hpp L63-67, log1p_exp.hpp
// L45-52, log1m_exp.hpp L47-57 -- exposed so the C restatement of each scalar
// kernel can be pinned pointwise.
double ref_digamma(double x) { return stan::math::digamma(x); }
double ref_lgamma(double x) { return stan::math::lgamma(x); }
double ref_log1p_exp(double x) { return stan::math::log1p_exp(x); }
double ref_log1m_exp(double x) { return stan::math::log1m_exp(x); }
// log_inv_logit.hpp L52-58, log1m_inv_logit.hpp L44-50,
// binomial_coefficient_log.hpp L79-141 (binomial_logit_glm_lpmf's kernels)
double ref_log_inv_logit(double x) { return stan::math::log_inv_logit(x); }
double ref_log1m_inv_logit(double x) { return stan::math::log1m_inv_logit(x); }
double ref_binomial_coefficient_log(double n, double k) {
  return stan::math::binomial_coefficient_log(n, k);
}

// ------------------------------------------------------------------- timing
// CPU baseline #1: the reference prim path, single call on
// one core: lpdf + grad() + recover_memory() per iteration, beta/alpha var,
// x data.  family: 0 normal(sigma scalar var) 1 bernoulli 2 poisson 3 negbin
// (phi scalar var).  Returns seconds per evaluation (best of `reps`).
double ref_time_glm(int family, long N, long K, const void* y, const double* x,
                    double alpha, const double* beta, double aux, int reps,
                    double* logp_out, double* d_beta_out) {
  Map<const MatrixXd> xm(x, N, K);
  Map<const VectorXd> bm(beta, K);
  std::vector<int> yi;
  VectorXd yd;
  if (family == 0) {
    yd = Map<const VectorXd>(static_cast<const double*>(y), N);
  } else {
    const int* p = static_cast<const int*>(y);
    yi.assign(p, p + N);
  }
  double best = 1e300;
  for (int r = 0; r < reps; ++r) {
    auto t0 = std::chrono::steady_clock::now();
    {
      VecV b = bm.cast<var>();
      var a = alpha;
      var s = aux;
      var lp;
      switch (family) {
        case 0:
          lp = stan::math::normal_id_glm_lpdf(yd, xm, a, b, s);
          break;
        case 1:
          lp = stan::math::bernoulli_logit_glm_lpmf(yi, xm, a, b);
          break;
        case 2:
          lp = stan::math::poisson_log_glm_lpmf(yi, xm, a, b);
          break;
        default:
          lp = stan::math::neg_binomial_2_log_glm_lpmf(yi, xm, a, b, s);
      }
      lp.grad();
      if (logp_out) *logp_out = lp.val();
      if (d_beta_out)
        for (long k = 0; k < K; ++k) d_beta_out[k] = b(k).adj();
    }
    stan::math::recover_memory();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now()
                                              - t0)
                    .count();
    if (dt < best) best = dt;
  }
  return best;
}

#ifdef STAN_THREADS
}  // extern "C"

namespace {
// Slice functor for reduce_sum: the GLM on the row
// block x.middleRows(start, len).
template <int Family>
struct glm_slice {
  template <typename TB, typename TA, typename TS>
  auto operator()(const std::vector<int>& y_slice, std::size_t start,
                  std::size_t end, std::ostream*, const Map<const MatrixXd>& x,
                  const Map<const VectorXd>& yd, const TA& alpha, const TB& beta,
                  const TS& aux) const {
    const std::size_t len = end - start + 1;
    if constexpr (Family == 0)
      return stan::math::normal_id_glm_lpdf(yd.segment(start, len),
                                            x.middleRows(start, len), alpha,
                                            beta, aux);
    else if constexpr (Family == 1)
      return stan::math::bernoulli_logit_glm_lpmf(
          y_slice, x.middleRows(start, len), alpha, beta);
    else if constexpr (Family == 2)
      return stan::math::poisson_log_glm_lpmf(
          y_slice, x.middleRows(start, len), alpha, beta);
    else
      return stan::math::neg_binomial_2_log_glm_lpmf(
          y_slice, x.middleRows(start, len), alpha, beta, aux);
  }
};
}  // namespace

extern "C" {
// CPU baseline #2: reduce_sum over TBB with `threads` threads.
// Returns seconds per evaluation (best of reps).
double ref_time_glm_reduce_sum(int family, long N, long K, const void* y,
                               const double* x, double alpha,
                               const double* beta, double aux, int threads,
                               long grainsize, int reps, double* logp_out,
                               double* d_beta_out) {
  stan::math::init_threadpool_tbb(threads);
  Map<const MatrixXd> xm(x, N, K);
  Map<const VectorXd> bm(beta, K);
  std::vector<int> yi(N, 0);
  VectorXd yd_store;
  const double* ydp = nullptr;
  if (family == 0) {
    ydp = static_cast<const double*>(y);
  } else {
    const int* p = static_cast<const int*>(y);
    yi.assign(p, p + N);
  }
  Map<const VectorXd> yd(ydp ? ydp : x, ydp ? N : 0);
  double best = 1e300;
  for (int r = 0; r < reps; ++r) {
    auto t0 = std::chrono::steady_clock::now();
    {
      VecV b = bm.cast<var>();
      var a = alpha;
      var s = aux;
      var lp;
      switch (family) {
        case 0:
          lp = stan::math::reduce_sum<glm_slice<0>>(yi, grainsize, nullptr, xm,
                                                    yd, a, b, s);
          break;
        case 1:
          lp = stan::math::reduce_sum<glm_slice<1>>(yi, grainsize, nullptr, xm,
                                                    yd, a, b, s);
          break;
        case 2:
          lp = stan::math::reduce_sum<glm_slice<2>>(yi, grainsize, nullptr, xm,
                                                    yd, a, b, s);
          break;
        default:
          lp = stan::math::reduce_sum<glm_slice<3>>(yi, grainsize, nullptr, xm,
                                                    yd, a, b, s);
      }
      lp.grad();
      if (logp_out) *logp_out = lp.val();
      if (d_beta_out)
        for (long k = 0; k < K; ++k) d_beta_out[k] = b(k).adj();
    }
    stan::math::recover_memory();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now()
                                              - t0)
                    .count();
    if (dt < best) best = dt;
  }
  return best;
}
#endif

int ref_has_threads(void) {
#ifdef STAN_THREADS
  return 1;
#else
  return 0;
#endif
}

}  // extern "C"
