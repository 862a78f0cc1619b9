// CUDA-vs-CPU parity for the step either side of the fused GLMs (SURVEY.md
// 8(f)3): the device matrix-vector product + add, and the un-fused densities on
// a device linear predictor.  The oracle is the reference's own prim
// implementation (bernoulli_logit_lpmf, poisson_log_lpmf,
// neg_binomial_2_log_lpmf, ordered_logistic_lpmf, categorical_logit_lpmf) compiled
// into this binary;
// the protocol is the one of test/unit/math/opencl/util.hpp L129-191.
#include "cuda_test_util.hpp"

#include <boost/random/mersenne_twister.hpp>

#include <array>

using Eigen::Dynamic;
using Eigen::Matrix;
using Eigen::MatrixXd;
using Eigen::VectorXd;
using stan::math::matrix_cuda;
using stan::math::var;
using std::vector;
using namespace cuda_test;  // NOLINT

namespace {
auto bern = [](const auto& n, const auto& theta) {
  return stan::math::bernoulli_logit_lpmf(n, theta);
};
auto bern_propto = [](const auto& n, const auto& theta) {
  return stan::math::bernoulli_logit_lpmf<true>(n, theta);
};
auto pois = [](const auto& n, const auto& alpha) {
  return stan::math::poisson_log_lpmf(n, alpha);
};
auto pois_propto = [](const auto& n, const auto& alpha) {
  return stan::math::poisson_log_lpmf<true>(n, alpha);
};
auto negb = [](const auto& n, const auto& eta, const auto& phi) {
  return stan::math::neg_binomial_2_log_lpmf(n, eta, phi);
};
auto negb_propto = [](const auto& n, const auto& eta, const auto& phi) {
  return stan::math::neg_binomial_2_log_lpmf<true>(n, eta, phi);
};
auto ordl = [](const auto& y, const auto& lambda, const auto& c) {
  return stan::math::ordered_logistic_lpmf(y, lambda, c);
};
auto ordl_propto = [](const auto& y, const auto& lambda, const auto& c) {
  return stan::math::ordered_logistic_lpmf<true>(y, lambda, c);
};

// categorical_logit_lpmf with one row of log odds per outcome: on the device one
// call, on the host the loop over the rows a Stan model writes with the
// reference's own (n, column vector) signature.
template <bool propto>
struct cat_rows {
  template <typename T_n, typename T_lin>
  auto operator()(const T_n& ns, const T_lin& lin) const {
    if constexpr (stan::is_cuda_operand<T_lin>::value) {
      return stan::math::categorical_logit_lpmf<propto>(ns, lin);
    } else {
      stan::return_type_t<T_lin> lp = 0.0;
      for (Eigen::Index i = 0; i < lin.rows(); ++i) {
        int n_i;
        if constexpr (std::is_same<T_n, int>::value) {
          n_i = ns;
        } else {
          n_i = ns[i];
        }
        Matrix<stan::value_type_t<T_lin>, Dynamic, 1> row = lin.row(i).transpose();
        lp += stan::math::categorical_logit_lpmf<propto>(n_i, row);
      }
      return lp;
    }
  }
};

VectorXd random_theta(int N, double scale, unsigned seed) {
  srand(seed);
  return VectorXd::Random(N) * scale;
}
}  // namespace

TEST(CudaUnfused, bernoulli_logit_lpmf) {
  vector<int> n{1, 0, 1, 1, 0};
  VectorXd theta(5);
  theta << 0.3, -2.5, 25.0, -31.0, 0.0;  // both tails of the |theta| > 20 cut-off
  compare_cpu_cuda_prim_rev(bern, std::make_tuple(DEV, DEV), n, theta);
  compare_cpu_cuda_prim_rev(bern_propto, std::make_tuple(DEV, DEV), n, theta);
  compare_cpu_cuda_prim_rev(bern, std::make_tuple(HOST, DEV), 1, theta);
  int N = 4099;
  vector<int> nb(N);
  for (int i = 0; i < N; ++i) nb[i] = (i * 7) % 2;
  compare_cpu_cuda_prim_rev(bern, std::make_tuple(DEV, DEV), nb, random_theta(N, 3.0, 1));
  // +-inf is a legal logit, NaN is not; sizes must agree; n in {0, 1}
  VectorXd t_inf(3), t_nan(3);
  t_inf << INFINITY, -INFINITY, 0.5;
  t_nan << 0.1, NAN, 0.5;
  vector<int> n3{1, 0, 1}, n_bad{0, 2, 1};
  matrix_cuda<int> n3_d(n3), n_bad_d(n_bad), n5_d(n);
  matrix_cuda<double> t_inf_d(t_inf), t_nan_d(t_nan);
  EXPECT_NEAR(stan::math::bernoulli_logit_lpmf(n3_d, t_inf_d),
              stan::math::bernoulli_logit_lpmf(n3, t_inf), 1e-14);
  EXPECT_THROW(stan::math::bernoulli_logit_lpmf(n3_d, t_nan_d), std::domain_error);
  EXPECT_THROW(stan::math::bernoulli_logit_lpmf(n_bad_d, t_inf_d), std::domain_error);
  EXPECT_THROW(stan::math::bernoulli_logit_lpmf(n5_d, t_inf_d), std::invalid_argument);
  vector<int> e{};
  compare_cpu_cuda_prim_rev(bern, std::make_tuple(DEV, DEV), e, VectorXd(0));
}

TEST(CudaUnfused, poisson_log_lpmf) {
  vector<int> n{14, 0, 5, 2};
  VectorXd alpha(4);
  alpha << 0.3, -2.5, 1.7, 0.0;
  compare_cpu_cuda_prim_rev(pois, std::make_tuple(DEV, DEV), n, alpha);
  compare_cpu_cuda_prim_rev(pois_propto, std::make_tuple(DEV, DEV), n, alpha);
  compare_cpu_cuda_prim_rev(pois, std::make_tuple(HOST, DEV), 3, alpha);
  int N = 4099;
  vector<int> nb(N);
  for (int i = 0; i < N; ++i) nb[i] = (i * 7) % 9;
  compare_cpu_cuda_prim_rev(pois, std::make_tuple(DEV, DEV), nb, random_theta(N, 2.0, 2));
  // log(0) exits and value checks (poisson_log_lpmf.hpp L46-66)
  VectorXd a_pinf(3), a_ninf(3), a_nan(3);
  a_pinf << 0.1, INFINITY, 0.5;
  a_ninf << 0.1, -INFINITY, 0.5;
  a_nan << 0.1, NAN, 0.5;
  vector<int> n3{1, 2, 1}, n_neg{0, -2, 1};
  matrix_cuda<int> n3_d(n3), n_neg_d(n_neg);
  matrix_cuda<double> a_pinf_d(a_pinf), a_ninf_d(a_ninf), a_nan_d(a_nan), a3_d(a_pinf);
  EXPECT_EQ(stan::math::poisson_log_lpmf(n3_d, a_pinf_d), stan::math::LOG_ZERO);
  EXPECT_EQ(stan::math::poisson_log_lpmf(n3_d, a_ninf_d), stan::math::LOG_ZERO);
  EXPECT_THROW(stan::math::poisson_log_lpmf(n3_d, a_nan_d), std::domain_error);
  EXPECT_THROW(stan::math::poisson_log_lpmf(n_neg_d, a3_d), std::domain_error);
  matrix_cuda<int> n4_d(n);
  EXPECT_THROW(stan::math::poisson_log_lpmf(n4_d, a3_d), std::invalid_argument);
}

TEST(CudaUnfused, neg_binomial_2_log_lpmf) {
  vector<int> n{14, 0, 5, 2};
  VectorXd eta(4);
  eta << 0.3, -2.5, 1.7, 0.0;
  double phi = 2.5;
  compare_cpu_cuda_prim_rev(negb, std::make_tuple(DEV, DEV, HOST), n, eta, phi);
  compare_cpu_cuda_prim_rev(negb_propto, std::make_tuple(DEV, DEV, HOST), n, eta, phi);
  compare_cpu_cuda_prim_rev(negb, std::make_tuple(HOST, DEV, HOST), 3, eta, phi);
  int N = 4099;
  vector<int> nb(N);
  for (int i = 0; i < N; ++i) nb[i] = (i * 7) % 9;
  compare_cpu_cuda_prim_rev(negb, std::make_tuple(DEV, DEV, HOST), nb,
                            random_theta(N, 2.0, 3), 0.7);
  compare_cpu_cuda_prim_rev(negb_propto, std::make_tuple(DEV, DEV, HOST), nb,
                            random_theta(N, 2.0, 3), 0.7);
  VectorXd e_inf(4);
  e_inf << 0.1, INFINITY, 0.5, 0.2;
  matrix_cuda<int> n_d(n);
  matrix_cuda<double> e_inf_d(e_inf), eta_d(eta);
  EXPECT_THROW(stan::math::neg_binomial_2_log_lpmf(n_d, e_inf_d, phi), std::domain_error);
  EXPECT_THROW(stan::math::neg_binomial_2_log_lpmf(n_d, eta_d, -1.0), std::domain_error);
  vector<int> n_neg{0, -2, 1, 3};
  matrix_cuda<int> n_neg_d(n_neg);
  EXPECT_THROW(stan::math::neg_binomial_2_log_lpmf(n_neg_d, eta_d, phi), std::domain_error);
}

TEST(CudaUnfused, ordered_logistic_lpmf) {
  vector<int> y{1, 1, 2, 4, 4, 3};
  VectorXd lambda(6);
  lambda << 1.9, 4.9, 7.9, 10.9, 3.6, -2.0;
  VectorXd c(3);
  c << 0.9, 1.1, 7;
  compare_cpu_cuda_prim_rev(ordl, std::make_tuple(DEV, DEV, HOST), y, lambda, c);
  compare_cpu_cuda_prim_rev(ordl_propto, std::make_tuple(DEV, DEV, HOST), y, lambda, c);
  compare_cpu_cuda_prim_rev(ordl, std::make_tuple(HOST, DEV, HOST), 2, lambda, c);
  int N = 4099;
  vector<int> yb(N);
  for (int i = 0; i < N; ++i) yb[i] = 1 + (i * 7) % 4;
  compare_cpu_cuda_prim_rev(ordl, std::make_tuple(DEV, DEV, HOST), yb,
                            random_theta(N, 3.0, 4), c);
  matrix_cuda<int> y_d(y);
  vector<int> y_bad{1, 1, 2, 5, 4, 3};
  matrix_cuda<int> y_bad_d(y_bad);
  matrix_cuda<double> l_d(lambda);
  VectorXd l_inf = lambda, c_unordered(3);
  l_inf[2] = INFINITY;
  c_unordered << 0.9, 0.1, 7;
  matrix_cuda<double> l_inf_d(l_inf);
  EXPECT_THROW(stan::math::ordered_logistic_lpmf(y_bad_d, l_d, c), std::domain_error);
  EXPECT_THROW(stan::math::ordered_logistic_lpmf(y_d, l_inf_d, c), std::domain_error);
  EXPECT_THROW(stan::math::ordered_logistic_lpmf(y_d, l_d, c_unordered), std::domain_error);
}

TEST(CudaUnfused, categorical_logit_lpmf_row_per_outcome) {
  vector<int> y{1, 3, 1, 2, 2};
  MatrixXd lin(5, 3);
  lin << 0.5, -2, 4, 1.3, 0.2, -0.7, -30, 2, 55, 0, 0, 0, 700, -700, 1e-3;
  compare_cpu_cuda_prim_rev(cat_rows<false>{}, std::make_tuple(DEV, DEV), y, lin);
  compare_cpu_cuda_prim_rev(cat_rows<true>{}, std::make_tuple(DEV, DEV), y, lin);
  compare_cpu_cuda_prim_rev(cat_rows<false>{}, std::make_tuple(HOST, DEV), y, lin);
  compare_cpu_cuda_prim_rev(cat_rows<false>{}, std::make_tuple(HOST, DEV), 2, lin);
  // register-resident rows (C <= 8, C <= 32) and the re-reading kernel (C = 43,
  // the class count of the reference's OpenCL categorical test)
  for (int C : {1, 2, 8, 9, 32, 43}) {
    const int N = C == 43 ? 1153 : 4099;
    srand(100 + C);
    MatrixXd big = MatrixXd::Random(N, C) * 4.0;
    vector<int> yb(N);
    for (int i = 0; i < N; ++i) yb[i] = 1 + (i * 7) % C;
    compare_cpu_cuda_prim_rev(cat_rows<false>{}, std::make_tuple(DEV, DEV), yb, big);
  }
  // the same call on a device matrix that came out of a device computation
  matrix_cuda<int> y_d(y);
  vector<int> y_bad{1, 3, 1, 4, 2}, y_short{1, 2, 3};
  matrix_cuda<int> y_bad_d(y_bad), y_short_d(y_short);
  matrix_cuda<double> lin_d(lin);
  MatrixXd lin_inf = lin, lin_nan = lin;
  lin_inf(2, 1) = INFINITY;
  lin_nan(4, 2) = NAN;
  matrix_cuda<double> lin_inf_d(lin_inf), lin_nan_d(lin_nan);
  EXPECT_THROW(stan::math::categorical_logit_lpmf(y_bad_d, lin_d), std::domain_error);
  EXPECT_THROW(stan::math::categorical_logit_lpmf(0, lin_d), std::domain_error);
  EXPECT_THROW(stan::math::categorical_logit_lpmf(y_d, lin_inf_d), std::domain_error);
  EXPECT_THROW(stan::math::categorical_logit_lpmf(y_d, lin_nan_d), std::domain_error);
  EXPECT_THROW(stan::math::categorical_logit_lpmf<true>(y_d, lin_nan_d), std::domain_error);
  EXPECT_THROW(stan::math::categorical_logit_lpmf(y_short_d, lin_d), std::invalid_argument);
  EXPECT_EQ(stan::math::categorical_logit_lpmf<true>(y_d, lin_d), 0.0);
  vector<int> e{};
  matrix_cuda<int> e_d(e);
  matrix_cuda<double> lin0_d(MatrixXd(0, 3));
  EXPECT_EQ(stan::math::categorical_logit_lpmf(e_d, lin0_d), 0.0);
}

TEST(CudaUnfused, normal_lpdf) {
  auto norm = [](const auto& y, const auto& mu, const auto& sigma) {
    return stan::math::normal_lpdf(y, mu, sigma);
  };
  auto norm_propto = [](const auto& y, const auto& mu, const auto& sigma) {
    return stan::math::normal_lpdf<true>(y, mu, sigma);
  };
  VectorXd y(5), mu(5);
  y << 14, 32, 21, -3.5, 0.25;
  mu << 12.5, 30, 22, -3, 0;
  double sigma = 1.7;
  compare_cpu_cuda_prim_rev(norm, std::make_tuple(DEV, DEV, HOST), y, mu, sigma);
  compare_cpu_cuda_prim_rev(norm_propto, std::make_tuple(DEV, DEV, HOST), y, mu, sigma);
  compare_cpu_cuda_prim_rev(norm, std::make_tuple(DEV, HOST, HOST), y, 1.5, sigma);
  compare_cpu_cuda_prim_rev(norm, std::make_tuple(HOST, DEV, HOST), 2.5, mu, sigma);
  int N = 4099;
  compare_cpu_cuda_prim_rev(norm, std::make_tuple(DEV, DEV, HOST), random_theta(N, 3.0, 7),
                            random_theta(N, 2.0, 8), 0.6);
  matrix_cuda<double> y_d(y), mu_d(mu);
  VectorXd y_nan = y, mu_inf = mu, y_inf = y;
  y_nan[1] = NAN;
  mu_inf[2] = INFINITY;
  y_inf[3] = INFINITY;
  matrix_cuda<double> y_nan_d(y_nan), mu_inf_d(mu_inf), y_inf_d(y_inf), short_d(VectorXd(VectorXd::Zero(4)));
  EXPECT_THROW(stan::math::normal_lpdf(y_nan_d, mu_d, sigma), std::domain_error);
  EXPECT_THROW(stan::math::normal_lpdf(y_d, mu_inf_d, sigma), std::domain_error);
  EXPECT_THROW(stan::math::normal_lpdf(y_d, mu_d, 0.0), std::domain_error);
  EXPECT_THROW(stan::math::normal_lpdf(y_d, short_d, sigma), std::invalid_argument);
  // an infinite observation is legal (check_not_nan only): -inf, as prim returns
  EXPECT_EQ(stan::math::normal_lpdf(y_inf_d, mu_d, sigma),
            stan::math::normal_lpdf(y_inf, mu, sigma));
}

TEST(CudaUnfused, multiply_add_then_density_matches_the_fused_glm) {
  // y ~ bernoulli_logit(alpha + x * beta) built step by step on the device, against
  // the same model on the host (the un-fused composition the reference's own GLM
  // tests compare with) and against the fused device GLM
  int N = 1531, K = 71;
  srand(5);
  MatrixXd x = MatrixXd::Random(N, K);
  VectorXd beta = VectorXd::Random(K) / std::sqrt(K);
  vector<int> y(N);
  for (int i = 0; i < N; ++i) y[i] = (i * 11) % 2;
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> y_d(y);

  // all data
  {
    auto theta_d = stan::math::add(stan::math::multiply(x_d, beta), 0.3);
    double dev = stan::math::bernoulli_logit_lpmf(y_d, theta_d);
    double cpu = stan::math::bernoulli_logit_lpmf(y, ((x * beta).array() + 0.3).matrix());
    expect_close("value", dev, cpu, kRelLogp, 0);
  }
  // alpha, beta autodiff
  {
    var a1 = 0.3, a2 = 0.3, a3 = 0.3;
    Matrix<var, Dynamic, 1> b1 = beta, b2 = beta, b3 = beta;
    var lp_dev = stan::math::bernoulli_logit_lpmf(
        y_d, stan::math::add(stan::math::multiply(x_d, b1), a1));
    Matrix<var, Dynamic, 1> th2 = stan::math::add(stan::math::multiply(x, b2), a2);
    var lp_cpu = stan::math::bernoulli_logit_lpmf(y, th2);
    var lp_glm = stan::math::bernoulli_logit_glm_lpmf(y_d, x_d, a3, b3);
    (lp_dev + lp_cpu + lp_glm).grad();
    expect_close("value vs host", lp_dev.val(), lp_cpu.val(), kRelLogp, 0);
    expect_close("value vs fused", lp_dev.val(), lp_glm.val(), kRelLogp, 0);
    expect_close("d_alpha", a1.adj(), a2.adj(), kRelGrad, std::fabs(a2.adj()));
    expect_close("d_alpha vs fused", a1.adj(), a3.adj(), kRelGrad, std::fabs(a3.adj()));
    const double scale = b2.adj().cwiseAbs().maxCoeff();
    for (int k = 0; k < K; ++k) {
      expect_close("d_beta", b1[k].adj(), b2[k].adj(), kRelGrad, scale);
      expect_close("d_beta vs fused", b1[k].adj(), b3[k].adj(), kRelGrad, scale);
    }
    stan::math::recover_memory();
  }
  // var_value<vector> beta, poisson
  {
    vector<int> yp(N);
    for (int i = 0; i < N; ++i) yp[i] = (i * 5) % 7;
    matrix_cuda<int> yp_d(yp);
    stan::math::var_value<VectorXd> b1(beta), b2(beta);
    var lp_dev = stan::math::poisson_log_lpmf(
        yp_d, stan::math::add(stan::math::multiply(x_d, b1), 0.1));
    var lp_glm = stan::math::poisson_log_glm_lpmf(yp_d, x_d, 0.1, b2);
    (lp_dev + lp_glm).grad();
    expect_close("poisson value vs fused", lp_dev.val(), lp_glm.val(), kRelLogp, 0);
    const double scale = b2.adj().cwiseAbs().maxCoeff();
    for (int k = 0; k < K; ++k)
      expect_close("poisson d_beta", b1.adj()[k], b2.adj()[k], kRelGrad, scale);
    stan::math::recover_memory();
  }
  EXPECT_THROW(stan::math::multiply(x_d, VectorXd(VectorXd::Zero(K + 1))),
               std::invalid_argument);
}

TEST(CudaUnfused, matrix_product_then_categorical_matches_the_fused_glm) {
  // lin = x beta + alpha^T on the device (FP64 tensor-core sweep), the un-fused
  // categorical density on it, and the reverse sweep x^T adj / column sums:
  // against prim's categorical_logit_glm_lpmf on the host and the fused device GLM
  for (auto shape : {std::array<int, 3>{1531, 37, 5}, std::array<int, 3>{4099, 128, 32},
                     std::array<int, 3>{777, 21, 70}, std::array<int, 3>{5, 2, 3}}) {
    const int N = shape[0], K = shape[1], C = shape[2];
    srand(11 + C);
    MatrixXd x = MatrixXd::Random(N, K);
    MatrixXd beta = MatrixXd::Random(K, C) / std::sqrt(double(K));
    VectorXd alpha = VectorXd::Random(C);
    vector<int> y(N);
    for (int i = 0; i < N; ++i) y[i] = 1 + (i * 7) % C;
    matrix_cuda<double> x_d(x);
    matrix_cuda<int> y_d(y);
    // all data: the product itself, then the density
    {
      MatrixXd lin_dev = stan::math::from_matrix_cuda(stan::math::multiply(x_d, beta));
      MatrixXd lin_cpu = x * beta;
      ASSERT_EQ(lin_dev.rows(), N);
      ASSERT_EQ(lin_dev.cols(), C);
      const double scale = lin_cpu.cwiseAbs().maxCoeff();
      for (int c = 0; c < C; ++c)
        for (int i = 0; i < N; i += 13)
          expect_close("x * beta", lin_dev(i, c), lin_cpu(i, c), kRelGrad, scale);
      double dev = stan::math::categorical_logit_lpmf(
          y_d, stan::math::linear_predictor(x_d, beta, alpha));
      double cpu = stan::math::categorical_logit_glm_lpmf(y, x, alpha, beta);
      expect_close("value", dev, cpu, kRelLogp, 0);
    }
    // alpha, beta autodiff (Eigen matrices of var)
    {
      Matrix<var, Dynamic, Dynamic> b1 = beta, b2 = beta, b3 = beta;
      Matrix<var, Dynamic, 1> a1 = alpha, a2 = alpha, a3 = alpha;
      var lp_dev = stan::math::categorical_logit_lpmf(
          y_d, stan::math::linear_predictor(x_d, b1, a1));
      var lp_cpu = stan::math::categorical_logit_glm_lpmf(y, x, a2, b2);
      var lp_glm = C <= 64 ? stan::math::categorical_logit_glm_lpmf(y_d, x_d, a3, b3)
                           : var(lp_cpu.val());
      (lp_dev + lp_cpu + lp_glm).grad();
      expect_close("value vs host", lp_dev.val(), lp_cpu.val(), kRelLogp, 0);
      expect_close("value vs fused", lp_dev.val(), lp_glm.val(), kRelLogp, 0);
      const double sa = a2.adj().cwiseAbs().maxCoeff(), sb = b2.adj().cwiseAbs().maxCoeff();
      for (int c = 0; c < C; ++c) {
        expect_close("d_alpha", a1[c].adj(), a2[c].adj(), kRelGrad, sa);
        if (C <= 64) expect_close("d_alpha vs fused", a1[c].adj(), a3[c].adj(), kRelGrad, sa);
        for (int k = 0; k < K; ++k) {
          expect_close("d_beta", b1(k, c).adj(), b2(k, c).adj(), kRelGrad, sb);
          if (C <= 64)
            expect_close("d_beta vs fused", b1(k, c).adj(), b3(k, c).adj(), kRelGrad, sb);
        }
      }
      stan::math::recover_memory();
    }
    // var_value<MatrixXd> beta through multiply (no intercept), data alpha dropped
    {
      stan::math::var_value<MatrixXd> b1(beta);
      Matrix<var, Dynamic, Dynamic> b2 = beta;
      var lp_dev = stan::math::categorical_logit_lpmf(y_d, stan::math::multiply(x_d, b1));
      var lp_cpu = stan::math::categorical_logit_glm_lpmf(y, x, VectorXd(VectorXd::Zero(C)), b2);
      (lp_dev + lp_cpu).grad();
      expect_close("multiply value", lp_dev.val(), lp_cpu.val(), kRelLogp, 0);
      const double sb = b2.adj().cwiseAbs().maxCoeff();
      for (int c = 0; c < C; ++c)
        for (int k = 0; k < K; ++k)
          expect_close("multiply d_beta", b1.adj()(k, c), b2(k, c).adj(), kRelGrad, sb);
      stan::math::recover_memory();
    }
  }
  // the other host containers: var_value<MatrixXd> weights, a row vector of var and a
  // var_value<VectorXd> as intercepts, data weights with autodiff intercepts
  {
    const int N = 257, K = 9, C = 4;
    srand(21);
    MatrixXd x = MatrixXd::Random(N, K), beta = MatrixXd::Random(K, C);
    VectorXd alpha = VectorXd::Random(C);
    vector<int> y(N);
    for (int i = 0; i < N; ++i) y[i] = 1 + (i * 3) % C;
    matrix_cuda<double> x_d(x);
    matrix_cuda<int> y_d(y);
    stan::math::var_value<MatrixXd> b1(beta);
    Matrix<var, 1, Dynamic> a1 = alpha.transpose();
    stan::math::var_value<VectorXd> a2(alpha);
    Matrix<var, Dynamic, 1> a3 = alpha, a_ref = alpha;
    Matrix<var, Dynamic, Dynamic> b_ref = beta;
    var lp1 = stan::math::categorical_logit_lpmf(y_d, stan::math::linear_predictor(x_d, b1, a1));
    var lp2 = stan::math::categorical_logit_lpmf(y_d, stan::math::linear_predictor(x_d, beta, a2));
    var lp3 = stan::math::categorical_logit_lpmf<true>(
        y_d, stan::math::linear_predictor(x_d, beta, a3));
    var lp_ref = stan::math::categorical_logit_glm_lpmf(y, x, a_ref, b_ref);
    (lp1 + lp2 + lp3 + lp_ref).grad();
    expect_close("value 1", lp1.val(), lp_ref.val(), kRelLogp, 0);
    expect_close("value 2", lp2.val(), lp_ref.val(), kRelLogp, 0);
    expect_close("value 3", lp3.val(), lp_ref.val(), kRelLogp, 0);
    const double sa = a_ref.adj().cwiseAbs().maxCoeff(), sb = b_ref.adj().cwiseAbs().maxCoeff();
    for (int c = 0; c < C; ++c) {
      expect_close("row-vector d_alpha", a1[c].adj(), a_ref[c].adj(), kRelGrad, sa);
      expect_close("var_value d_alpha", a2.adj()[c], a_ref[c].adj(), kRelGrad, sa);
      expect_close("propto d_alpha", a3[c].adj(), a_ref[c].adj(), kRelGrad, sa);
      for (int k = 0; k < K; ++k)
        expect_close("var_value d_beta", b1.adj()(k, c), b_ref(k, c).adj(), kRelGrad, sb);
    }
    stan::math::recover_memory();
  }
  matrix_cuda<double> x_d(MatrixXd(MatrixXd::Zero(4, 3)));
  EXPECT_THROW(stan::math::multiply(x_d, MatrixXd(MatrixXd::Zero(4, 2))),
               std::invalid_argument);
  EXPECT_THROW(stan::math::linear_predictor(x_d, MatrixXd(MatrixXd::Zero(3, 2)),
                                            VectorXd(VectorXd::Zero(3))),
               std::invalid_argument);
}

TEST(CudaUnfused, bernoulli_logit_glm_rng_draws_what_prim_draws) {
  // same generator state in, same variates out
  int N = 2003, K = 37;
  srand(6);
  MatrixXd x = MatrixXd::Random(N, K);
  VectorXd beta = VectorXd::Random(K);
  VectorXd alpha = VectorXd::Random(N);
  matrix_cuda<double> x_d(x);
  {
    boost::random::mt19937 rng_cpu(1234), rng_dev(1234);
    vector<int> cpu = stan::math::bernoulli_logit_glm_rng(x, alpha, beta, rng_cpu);
    vector<int> dev = stan::math::bernoulli_logit_glm_rng(x_d, alpha, beta, rng_dev);
    ASSERT_EQ(cpu.size(), dev.size());
    int differ = 0, ones = 0;
    for (int i = 0; i < N; ++i) {
      differ += cpu[i] != dev[i];
      ones += dev[i];
    }
    EXPECT_EQ(differ, 0);
    EXPECT_GT(ones, N / 4);
    EXPECT_LT(ones, 3 * N / 4);
    EXPECT_EQ(rng_cpu(), rng_dev());  // the generators advanced in step
  }
  {  // std::vector intercepts and weights
    boost::random::mt19937 rng_cpu(99), rng_dev(99);
    vector<double> a(alpha.data(), alpha.data() + N), b(beta.data(), beta.data() + K);
    EXPECT_EQ(stan::math::bernoulli_logit_glm_rng(x, a, b, rng_cpu),
              stan::math::bernoulli_logit_glm_rng(x_d, a, b, rng_dev));
  }
  {  // error behaviour, prim L55-62
    boost::random::mt19937 rng(1);
    VectorXd beta_bad = beta, alpha_short = alpha.head(N - 1), beta_long(K + 1);
    beta_bad[3] = INFINITY;
    beta_long.setZero();
    MatrixXd x_bad = x;
    x_bad(5, 5) = NAN;
    matrix_cuda<double> x_bad_d(x_bad);
    EXPECT_THROW(stan::math::bernoulli_logit_glm_rng(x_d, alpha, beta_long, rng),
                 std::invalid_argument);
    EXPECT_THROW(stan::math::bernoulli_logit_glm_rng(x_d, alpha_short, beta, rng),
                 std::invalid_argument);
    EXPECT_THROW(stan::math::bernoulli_logit_glm_rng(x_d, alpha, beta_bad, rng),
                 std::domain_error);
    EXPECT_THROW(stan::math::bernoulli_logit_glm_rng(x_bad_d, alpha, beta, rng),
                 std::domain_error);
  }
}

TEST(CudaUnfused, hierarchical_intercept_through_device_indexing) {
  // y ~ poisson_log_glm(x, z[group], beta): the N-vector intercept is gathered on
  // the device from G group effects and its partials are summed back per group
  // there; oracle: the same model with a host Matrix<var> alpha
  int N = 6007, K = 23, G = 37;
  srand(9);
  MatrixXd x = MatrixXd::Random(N, K);
  VectorXd beta = VectorXd::Random(K) / std::sqrt(K), z = VectorXd::Random(G);
  vector<int> y(N), group(N);
  for (int i = 0; i < N; ++i) {
    y[i] = (i * 5) % 7;
    group[i] = (i * 31 + (i / 97)) % G;
  }
  matrix_cuda<double> x_d(x);
  matrix_cuda<int> y_d(y), group_d(group);

  // data only
  VectorXd alpha_host(N);
  for (int i = 0; i < N; ++i) alpha_host[i] = z[group[i]];
  expect_close("value (data)",
               stan::math::poisson_log_glm_lpmf(y_d, x_d, stan::math::indexing(z, group_d), beta),
               stan::math::poisson_log_glm_lpmf(y, x, alpha_host, beta), kRelLogp, 0);

  Matrix<var, Dynamic, 1> z1 = z, z2 = z, b1 = beta, b2 = beta;
  var lp_dev = stan::math::poisson_log_glm_lpmf(y_d, x_d, stan::math::indexing(z1, group_d), b1);
  Matrix<var, Dynamic, 1> alpha2(N);
  for (int i = 0; i < N; ++i) alpha2[i] = z2[group[i]];
  var lp_cpu = stan::math::poisson_log_glm_lpmf(y, x, alpha2, b2);
  (lp_dev + lp_cpu).grad();
  expect_close("value", lp_dev.val(), lp_cpu.val(), kRelLogp, 0);
  const double zs = z2.adj().cwiseAbs().maxCoeff(), bs = b2.adj().cwiseAbs().maxCoeff();
  for (int g = 0; g < G; ++g) expect_close("d_z", z1[g].adj(), z2[g].adj(), kRelGrad, zs);
  for (int k = 0; k < K; ++k) expect_close("d_beta", b1[k].adj(), b2[k].adj(), kRelGrad, bs);
  stan::math::recover_memory();

  // more groups than the shared-memory accumulators hold (host route), and a bad index
  {
    int G2 = 5000;
    VectorXd zz = VectorXd::Random(G2);
    vector<int> grp(N);
    for (int i = 0; i < N; ++i) grp[i] = (i * 7919) % G2;
    matrix_cuda<int> grp_d(grp);
    stan::math::var_value<VectorXd> zv1(zz), zv2(zz);
    var a = stan::math::bernoulli_logit_lpmf(
        matrix_cuda<int>(vector<int>(N, 1)), stan::math::indexing(zv1, grp_d));
    Matrix<var, Dynamic, 1> th(N);
    for (int i = 0; i < N; ++i) th[i] = zv2.coeff(grp[i]);
    var b = stan::math::bernoulli_logit_lpmf(vector<int>(N, 1), th);
    (a + b).grad();
    expect_close("value (many groups)", a.val(), b.val(), kRelLogp, 0);
    const double s = zv2.adj().cwiseAbs().maxCoeff();
    for (int g = 0; g < G2; ++g)
      expect_close("d_z (many groups)", zv1.adj()[g], zv2.adj()[g], kRelGrad, s);
    stan::math::recover_memory();
  }
  vector<int> bad = group;
  bad[100] = G;
  matrix_cuda<int> bad_d(bad);
  EXPECT_THROW(stan::math::indexing(z, bad_d), std::out_of_range);
}
