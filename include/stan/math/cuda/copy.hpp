#ifndef STAN_MATH_CUDA_COPY_HPP
#define STAN_MATH_CUDA_COPY_HPP
// Host <-> device copies for matrix_cuda: the analogue of to_matrix_cl
// (stan/math/opencl/copy.hpp L45) and from_matrix_cl<T_dst> (L61-235) for Eigen
// objects, std::vector and scalars.
#include <stan/math/cuda/matrix_cuda.hpp>

#include <vector>

namespace stan {
namespace math {

/** Eigen matrix / vector / expression -> device. */
template <typename Mat, require_eigen_t<Mat>* = nullptr,
          require_st_arithmetic<Mat>* = nullptr>
inline matrix_cuda<value_type_t<Mat>> to_matrix_cuda(const Mat& m) {
  return matrix_cuda<value_type_t<Mat>>(m);
}

/** std::vector<double|int> -> n x 1 device column. */
template <typename T, require_arithmetic_t<T>* = nullptr>
inline matrix_cuda<T> to_matrix_cuda(const std::vector<T>& v) {
  return matrix_cuda<T>(v);
}

/** std::vector of Eigen vectors -> device matrix, one vector per column
 * (opencl/copy.hpp L57-60 through matrix_cl.hpp L246-268). */
template <typename Vec, require_std_vector_vt<is_eigen, Vec>* = nullptr,
          require_st_arithmetic<Vec>* = nullptr>
inline matrix_cuda<scalar_type_t<Vec>> to_matrix_cuda(const Vec& v) {
  return matrix_cuda<scalar_type_t<Vec>>(v);
}

/** Scalar -> 1 x 1 device matrix (matrix_cl.hpp L349-356). */
template <typename T, require_arithmetic_t<T>* = nullptr>
inline matrix_cuda<T> to_matrix_cuda(T v) {
  matrix_cuda<T> m(1, 1);
  check_cuda_status("to_matrix_cuda", smc_matrix_upload(m.handle(), &v, 1));
  return m;
}

/** Eigen / std::vector -> device matrix partitioned row-wise over the GPUs of the
 * shard set (cuda_shard_init): scattered once, then accepted by every GLM overload
 * in place of a plain matrix_cuda. */
template <typename Mat, require_eigen_t<Mat>* = nullptr,
          require_st_arithmetic<Mat>* = nullptr>
inline matrix_cuda<value_type_t<Mat>> to_matrix_cuda_sharded(const Mat& m) {
  return matrix_cuda<value_type_t<Mat>>::sharded(m);
}
template <typename T, require_arithmetic_t<T>* = nullptr>
inline matrix_cuda<T> to_matrix_cuda_sharded(const std::vector<T>& v) {
  return matrix_cuda<T>::sharded(v);
}

/** matrix_cuda_sharded<T>: a matrix_cuda<T> whose constructors scatter over the shard
 * set -- the multi-GPU form of the reference's one-device matrix_cl
 * (opencl/opencl_context.hpp L75-76).  It IS a matrix_cuda<T>: every overload and
 * every helper takes it unchanged. */
template <typename T>
class matrix_cuda_sharded : public matrix_cuda<T> {
 public:
  matrix_cuda_sharded() = default;
  matrix_cuda_sharded(int64_t rows, int64_t cols)
      : matrix_cuda<T>(matrix_cuda<T>::sharded(rows, cols)) {}
  template <typename Mat, require_eigen_t<Mat>* = nullptr,
            require_same_t<value_type_t<Mat>, T>* = nullptr>
  explicit matrix_cuda_sharded(const Mat& m) : matrix_cuda<T>(matrix_cuda<T>::sharded(m)) {}
  explicit matrix_cuda_sharded(const std::vector<T>& v)
      : matrix_cuda<T>(matrix_cuda<T>::sharded(v)) {}
};

/** The shard set: one shard per visible GPU (n_shards <= 0) or the first n_shards;
 * returns the number of shards.  Call once, before the first sharded matrix. */
inline int cuda_shard_init(int n_shards = 0) {
  check_cuda_status("cuda_shard_init", smc_shard_init(n_shards, nullptr));
  int n = 0;
  check_cuda_status("cuda_shard_init", smc_shard_count(&n));
  return n;
}
inline int cuda_shard_count() {
  int n = 0;
  check_cuda_status("cuda_shard_count", smc_shard_count(&n));
  return n;
}
inline void cuda_shard_shutdown() {
  check_cuda_status("cuda_shard_shutdown", smc_shard_shutdown());
}

/** Already on the device: pass through. */
template <typename T>
inline const matrix_cuda<T>& to_matrix_cuda(const matrix_cuda<T>& m) {
  return m;
}
template <typename T>
inline matrix_cuda<T> to_matrix_cuda(matrix_cuda<T>&& m) {
  return std::move(m);
}

/** Device -> Eigen (default: dynamic matrix of the element type). */
template <typename T_dst, typename T, require_eigen_t<T_dst>* = nullptr>
inline T_dst from_matrix_cuda(const matrix_cuda<T>& src) {
  static_assert(std::is_same<value_type_t<T_dst>, T>::value,
                "from_matrix_cuda: element types differ");
  if (T_dst::RowsAtCompileTime == 1 || T_dst::ColsAtCompileTime == 1) {
    // vectors: accept n x 1 or 1 x n on the device, as the reference does
    check_size_match("from_matrix_cuda", "vector dimension of src",
                     src.rows() == 1 || src.cols() == 1 || src.size() == 0, "1", 1);
  }
  Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> tmp(src.rows(), src.cols());
  if (src.size() > 0) {
    check_cuda_status("from_matrix_cuda",
                      smc_matrix_download(src.handle(), tmp.data(), src.rows()));
  }
  T_dst dst;
  if (T_dst::RowsAtCompileTime == 1 && T_dst::ColsAtCompileTime != 1) {
    dst = Eigen::Map<const Eigen::Matrix<T, 1, Eigen::Dynamic>>(tmp.data(),
                                                                 tmp.size());
  } else if (T_dst::ColsAtCompileTime == 1 && T_dst::RowsAtCompileTime != 1) {
    dst = Eigen::Map<const Eigen::Matrix<T, Eigen::Dynamic, 1>>(tmp.data(),
                                                                 tmp.size());
  } else {
    dst = tmp;
  }
  return dst;
}

// (the leading non-type parameter keeps from_matrix_cuda<double>(m) from matching
// this overload)
template <int Unused = 0, typename T>
inline Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> from_matrix_cuda(
    const matrix_cuda<T>& src) {
  return from_matrix_cuda<Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>>(src);
}

/** Device -> std::vector of Eigen vectors, one per column (opencl/copy.hpp L213-224). */
template <typename T_dst, typename T,
          require_std_vector_vt<is_eigen_vector, T_dst>* = nullptr>
inline T_dst from_matrix_cuda(const matrix_cuda<T>& src) {
  Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> tmp = from_matrix_cuda(src);
  T_dst dst;
  dst.reserve(src.cols());
  for (int64_t i = 0; i < src.cols(); ++i) {
    dst.emplace_back(tmp.col(i));
  }
  return dst;
}

/** Device -> std::vector (column-major order). */
template <typename T_dst, typename T, require_std_vector_t<T_dst>* = nullptr,
          require_not_std_vector_vt<is_eigen_vector, T_dst>* = nullptr>
inline T_dst from_matrix_cuda(const matrix_cuda<T>& src) {
  Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> tmp = from_matrix_cuda(src);
  return T_dst(tmp.data(), tmp.data() + tmp.size());
}

/** 1 x 1 device matrix -> scalar. */
template <typename T_dst, typename T, require_arithmetic_t<T_dst>* = nullptr>
inline T_dst from_matrix_cuda(const matrix_cuda<T>& src) {
  check_size_match("from_matrix_cuda<scalar>", "src.rows()", src.rows(),
                   "dst.rows()", 1);
  check_size_match("from_matrix_cuda<scalar>", "src.cols()", src.cols(),
                   "dst.cols()", 1);
  T v;
  check_cuda_status("from_matrix_cuda",
                    smc_matrix_download(src.handle(), &v, 1));
  return static_cast<T_dst>(v);
}

}  // namespace math
}  // namespace stan
#endif
