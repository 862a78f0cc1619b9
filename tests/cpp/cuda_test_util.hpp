// Device-vs-CPU parity protocol for the CUDA GLM overloads, modelled on the
// reference's OpenCL protocol (test/unit/math/opencl/util.hpp L129-191,
// compare_cpu_opencl_prim_rev): for EVERY prim / var combination of the
// arguments, evaluate the functor with host (Eigen) arguments -- that is the
// reference's own prim implementation, the oracle -- and with the arguments
// placed on the GPU, compare the values, then grad() the sum and compare every
// adjoint.  Tolerances are BASELINE.json's: relative 1e-10 on the log density,
// 1e-9 on gradients, with an absolute floor scaled by the largest entry.
#ifndef TESTS_CPP_CUDA_TEST_UTIL_HPP
#define TESTS_CPP_CUDA_TEST_UTIL_HPP

#include <stan/math.hpp>
#include <stan/math/cuda.hpp>
#include <gtest/gtest.h>

#include <cmath>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace cuda_test {

using stan::math::var;

constexpr double kRelLogp = 1e-10;
constexpr double kRelGrad = 1e-9;
constexpr double kAbsFloor = 1e-12;

/** Where an argument lives for the device evaluation. */
struct host_t {};
struct dev_t {};
struct shard_t {};  // on the device, row-sharded over the shard set
constexpr host_t HOST{};
constexpr dev_t DEV{};
constexpr shard_t SHARD{};

inline void expect_close(const std::string& what, double dev, double cpu, double rel,
                         double scale) {
  if (std::isnan(cpu) || std::isnan(dev)) {
    EXPECT_TRUE(std::isnan(cpu) && std::isnan(dev)) << what;
    return;
  }
  const double tol = rel * std::fabs(cpu) + kAbsFloor * (scale > 1.0 ? scale : 1.0);
  EXPECT_NEAR(dev, cpu, tol) << what;
}

// ---- promotion of a prim argument to its autodiff form ---------------------
template <typename T>
struct promotable : std::false_type {};
template <>
struct promotable<double> : std::true_type {};
template <int R, int C>
struct promotable<Eigen::Matrix<double, R, C>> : std::true_type {};

template <bool V, typename T>
auto promote(const T& a) {
  if constexpr (!V) {
    return a;
  } else if constexpr (std::is_same<T, double>::value) {
    return var(a);
  } else {
    return Eigen::Matrix<var, T::RowsAtCompileTime, T::ColsAtCompileTime>(a);
  }
}

template <typename P, typename T>
decltype(auto) place(P, const T& a) {
  if constexpr (std::is_same<P, dev_t>::value) {
    return stan::math::to_matrix_cuda(a);
  } else if constexpr (std::is_same<P, shard_t>::value) {
    return stan::math::to_matrix_cuda_sharded(a);
  } else {
    return (a);
  }
}
// scalars always stay on the host
template <typename P>
double place(P, const double& a) { return a; }
template <typename P>
var place(P, const var& a) { return a; }
template <typename P>
int place(P, const int& a) { return a; }

template <typename T>
void compare_adj(const std::string&, const T&, const T&) {}
inline void compare_adj(const std::string& what, const var& dev, const var& cpu) {
  expect_close(what, dev.adj(), cpu.adj(), kRelGrad, std::fabs(cpu.adj()));
}
template <int R, int C>
void compare_adj(const std::string& what, const Eigen::Matrix<var, R, C>& dev,
                 const Eigen::Matrix<var, R, C>& cpu) {
  ASSERT_EQ(dev.rows(), cpu.rows()) << what;
  ASSERT_EQ(dev.cols(), cpu.cols()) << what;
  const Eigen::MatrixXd a = dev.adj(), b = cpu.adj();
  const double scale = b.size() ? b.cwiseAbs().maxCoeff() : 0.0;
  for (Eigen::Index j = 0; j < b.cols(); ++j) {
    for (Eigen::Index i = 0; i < b.rows(); ++i) {
      expect_close(what + " [" + std::to_string(i) + "," + std::to_string(j) + "]",
                   a(i, j), b(i, j), kRelGrad, scale);
    }
  }
}

template <std::size_t Mask, typename F, typename Places, typename Args,
          std::size_t... Is>
void run_combo(const F& f, const Places& places, const Args& args,
               std::index_sequence<Is...>) {
  constexpr bool valid
      = (((((Mask >> Is) & 1u) == 0)
          || promotable<std::decay_t<std::tuple_element_t<Is, Args>>>::value)
         && ...);
  if constexpr (valid) {
    const std::string sig = "var mask " + std::to_string(Mask);
    auto cpu = std::make_tuple(promote<((Mask >> Is) & 1u) != 0>(std::get<Is>(args))...);
    auto dev = std::make_tuple(promote<((Mask >> Is) & 1u) != 0>(std::get<Is>(args))...);
    auto res_cpu = f(std::get<Is>(cpu)...);
    auto res_dev = f(place(std::get<Is>(places), std::get<Is>(dev))...);
    const double vc = stan::math::value_of(res_cpu), vd = stan::math::value_of(res_dev);
    expect_close("log density, " + sig, vd, vc, kRelLogp, 0.0);
    if constexpr (Mask != 0) {
      var total = res_cpu + res_dev;
      total.grad();
      (compare_adj("adjoint of argument " + std::to_string(Is) + ", " + sig,
                   std::get<Is>(dev), std::get<Is>(cpu)),
       ...);
    }
    stan::math::recover_memory();
  }
}

template <typename F, typename Places, typename Args, std::size_t... Ms>
void run_all(const F& f, const Places& places, const Args& args,
             std::index_sequence<Ms...>) {
  (run_combo<Ms>(f, places, args,
                 std::make_index_sequence<std::tuple_size<Args>::value>{}),
   ...);
}

/** compare_cpu_cuda_prim_rev(f, places(...), args...): see the file comment. */
template <typename F, typename... Ps, typename... Ts>
void compare_cpu_cuda_prim_rev(const F& f, const std::tuple<Ps...>& places,
                               const Ts&... args) {
  static_assert(sizeof...(Ps) == sizeof...(Ts), "one placement per argument");
  run_all(f, places, std::make_tuple(args...),
          std::make_index_sequence<(std::size_t{1} << sizeof...(Ts))>{});
}

}  // namespace cuda_test
#endif
