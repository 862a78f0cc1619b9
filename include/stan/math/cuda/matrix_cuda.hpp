#ifndef STAN_MATH_CUDA_MATRIX_CUDA_HPP
#define STAN_MATH_CUDA_MATRIX_CUDA_HPP
// matrix_cuda<T>: a column-major matrix resident in B200 HBM, the analogue of
// matrix_cl<T> (stan/math/opencl/matrix_cl.hpp L46-55: buffer + rows_/cols_,
// double and int element types).  It owns an opaque smc_matrix handle of
// libstanmath_cuda.so; the design matrix is uploaded once (model constructor) and
// reused by every log-density evaluation.
//
// Semantics follow matrix_cl: copy construction / assignment is a deep
// device-to-device copy (matrix_cl.hpp L198-210), moves transfer the buffer,
// construction from host containers uploads (L319-370), (rows, cols) allocates
// only (L284-300).  A non-owning view of another matrix (or of a raw device
// pointer, cf. L190-192) is made with matrix_cuda::view().
#include <stan/math/cuda/err.hpp>
#include <stan/math/prim/fun/Eigen.hpp>
#include <stan/math/prim/meta.hpp>

#include <cstdint>
#include <type_traits>
#include <utility>
#include <vector>

namespace stan {
namespace math {

namespace internal {
template <typename T>
struct cuda_dtype;
template <>
struct cuda_dtype<double> {
  static constexpr int value = SMC_F64;
};
template <>
struct cuda_dtype<int> {
  static constexpr int value = SMC_I32;
};
}  // namespace internal

class matrix_cuda_base {};

template <typename T>
class matrix_cuda : public matrix_cuda_base {
 public:
  using Scalar = T;
  using type = T;

  matrix_cuda() = default;

  /** Allocates rows x cols on the device; contents are unspecified. */
  matrix_cuda(int64_t rows, int64_t cols) { allocate(rows, cols); }

  /** Deep copy on the device (sharded like the source). */
  matrix_cuda(const matrix_cuda& o) {
    if (o.handle_) {
      check_cuda_status("matrix_cuda(copy)",
                        smc_matrix_create_like(o.handle_, -1, -1, &handle_));
      check_cuda_status("matrix_cuda(copy)", smc_matrix_copy(handle_, o.handle_));
    }
  }
  matrix_cuda(matrix_cuda&& o) noexcept : handle_(o.handle_) { o.handle_ = nullptr; }

  /** Uploads a dense Eigen object (matrix, vector, row vector or expression). */
  template <typename Mat, require_eigen_t<Mat>* = nullptr,
            require_same_t<value_type_t<Mat>, T>* = nullptr>
  explicit matrix_cuda(const Mat& m) {
    const auto& ref = m.eval();
    allocate(ref.rows(), ref.cols());
    if (ref.size() > 0) {
      // column-major source: leading dimension = rows (a RowMajor vector is
      // contiguous too, so the same call covers row vectors)
      Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> cm;
      const T* src = ref.data();
      if (std::decay_t<decltype(ref)>::IsRowMajor && ref.rows() > 1) {
        cm = ref;
        src = cm.data();
      }
      check_cuda_status("matrix_cuda(Eigen)",
                        smc_matrix_upload(handle_, src, ref.rows()));
    }
  }

  /** Uploads a std::vector as an n x 1 column (matrix_cl.hpp L359+). */
  explicit matrix_cuda(const std::vector<T>& v) {
    allocate(static_cast<int64_t>(v.size()), 1);
    if (!v.empty()) {
      check_cuda_status("matrix_cuda(std::vector)",
                        smc_matrix_upload(handle_, v.data(), v.size()));
    }
  }

  /** Uploads a std::vector of Eigen vectors of equal length, one per COLUMN
   * (matrix_cl.hpp L246-268): the form in which per-outcome cut points reach
   * ordered_logistic_lpmf. */
  template <typename Vec, require_std_vector_vt<is_eigen, Vec>* = nullptr,
            require_st_same<Vec, T>* = nullptr>
  explicit matrix_cuda(const Vec& A) {
    const int64_t rows = A.empty() ? 0 : static_cast<int64_t>(A[0].size());
    const int64_t cols = static_cast<int64_t>(A.size());
    Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> cm(rows, cols);
    for (int64_t i = 0; i < cols; ++i) {
      check_size_match("matrix constructor", "input rows", A[i].size(),
                       "matrix_cuda rows", rows);
      cm.col(i) = Eigen::Map<const Eigen::Matrix<T, Eigen::Dynamic, 1>>(
          A[i].eval().data(), rows);
    }
    allocate(rows, cols);
    if (cm.size() > 0) {
      check_cuda_status("matrix_cuda(std::vector<Eigen>)",
                        smc_matrix_upload(handle_, cm.data(), rows));
    }
  }

  ~matrix_cuda() { release(); }

  matrix_cuda& operator=(const matrix_cuda& o) {
    if (this != &o) {
      matrix_cuda tmp(o);
      std::swap(handle_, tmp.handle_);
    }
    return *this;
  }
  matrix_cuda& operator=(matrix_cuda&& o) noexcept {
    if (this != &o) {
      release();
      handle_ = o.handle_;
      o.handle_ = nullptr;
    }
    return *this;
  }

  /** Non-owning view of device memory laid out column-major with stride ld. */
  static matrix_cuda view(T* device_ptr, int64_t rows, int64_t cols, int64_t ld) {
    matrix_cuda m;
    check_cuda_status("matrix_cuda::view",
                      smc_matrix_wrap(device_ptr, rows, cols, ld,
                                      internal::cuda_dtype<T>::value, &m.handle_));
    return m;
  }
  /** Non-owning view of another matrix_cuda (which must outlive the view); a view
   * of a row-sharded matrix is sharded the same way. */
  static matrix_cuda view(const matrix_cuda& o) {
    matrix_cuda m;
    if (o.handle_) {
      check_cuda_status("matrix_cuda::view", smc_matrix_view(o.handle_, &m.handle_));
    }
    return m;
  }
  /** Non-owning view of the matrix behind a C-ABI handle (which must outlive it). */
  static matrix_cuda view_of_handle(const smc_matrix* h) {
    matrix_cuda m;
    if (h) {
      check_cuda_status("matrix_cuda::view", smc_matrix_view(h, &m.handle_));
    }
    return m;
  }
  /** Allocates rows x cols partitioned row-wise over the GPUs of the shard set
   * (smc_shard_init / stan::math::cuda_shard_init): shard g holds rows
   * [g N / G, (g+1) N / G).  A sharded matrix goes wherever a GLM takes a plain one
   * -- x and its per-row operands sharded alike -- and the evaluation then runs on
   * every GPU with one NCCL all-reduce of the packed partials. */
  static matrix_cuda sharded(int64_t rows, int64_t cols) {
    matrix_cuda m;
    check_cuda_status("matrix_cuda::sharded",
                      smc_sharded_matrix_create(rows, cols, internal::cuda_dtype<T>::value,
                                                &m.handle_));
    return m;
  }
  /** Scatters a dense column-major Eigen object over the shard set (done once, in
   * the model constructor). */
  template <typename Mat, require_eigen_t<Mat>* = nullptr,
            require_same_t<value_type_t<Mat>, T>* = nullptr>
  static matrix_cuda sharded(const Mat& m) {
    const Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic> cm = m;
    matrix_cuda out = sharded(cm.rows(), cm.cols());
    if (cm.size() > 0) {
      check_cuda_status("matrix_cuda::sharded(Eigen)",
                        smc_matrix_upload(out.handle_, cm.data(), cm.rows()));
    }
    return out;
  }
  static matrix_cuda sharded(const std::vector<T>& v) {
    matrix_cuda out = sharded(static_cast<int64_t>(v.size()), 1);
    if (!v.empty()) {
      check_cuda_status("matrix_cuda::sharded(std::vector)",
                        smc_matrix_upload(out.handle_, v.data(), v.size()));
    }
    return out;
  }
  /** Allocates a matrix with the rows -- and the row partition, if `like` is sharded
   * -- of `like` and `cols` columns (contents unspecified). */
  template <typename U>
  static matrix_cuda like(const matrix_cuda<U>& like, int64_t cols) {
    return like_handle(like.handle(), like.rows(), cols);
  }
  static matrix_cuda like_handle(const smc_matrix* h, int64_t rows, int64_t cols) {
    matrix_cuda m;
    if (h) {
      check_cuda_status("matrix_cuda::like",
                        smc_matrix_create_like(h, cols, internal::cuda_dtype<T>::value,
                                               &m.handle_));
    } else {
      m.allocate(rows, cols);
    }
    return m;
  }
  /** Number of shards (0: a plain single-GPU matrix). */
  int shard_count() const noexcept { return handle_ ? smc_matrix_shard_count(handle_) : 0; }
  /** Takes ownership of a handle made through the C ABI. */
  static matrix_cuda adopt(smc_matrix* h) {
    matrix_cuda m;
    m.handle_ = h;
    return m;
  }
  /** Gives up ownership of the handle (the caller frees it). */
  smc_matrix* release_handle() noexcept {
    smc_matrix* h = handle_;
    handle_ = nullptr;
    return h;
  }

  int64_t rows() const noexcept { return handle_ ? smc_matrix_rows(handle_) : 0; }
  int64_t cols() const noexcept { return handle_ ? smc_matrix_cols(handle_) : 0; }
  int64_t size() const noexcept { return rows() * cols(); }
  /** The C-ABI handle (the analogue of matrix_cl::buffer(), L177-178). */
  smc_matrix* handle() const noexcept { return handle_; }

  void zero() {
    if (handle_) {
      check_cuda_status("matrix_cuda::zero", smc_matrix_zero(handle_));
    }
  }

  /** Deterministic synthetic contents (see smc_matrix_fill_synthetic). */
  void fill_synthetic(uint64_t seed, int64_t row0, int kind, double scale, int lo,
                      int hi) {
    check_cuda_status("matrix_cuda::fill_synthetic",
                      smc_matrix_fill_synthetic(handle_, seed, row0, kind, scale,
                                                lo, hi));
  }

 private:
  void allocate(int64_t rows, int64_t cols) {
    static_assert(std::is_same<T, double>::value || std::is_same<T, int>::value,
                  "matrix_cuda holds double or int");
    check_cuda_status("matrix_cuda",
                      smc_matrix_create(rows, cols, internal::cuda_dtype<T>::value,
                                        &handle_));
  }
  void release() noexcept {
    if (handle_) {
      smc_matrix_free(handle_);
      handle_ = nullptr;
    }
  }
  smc_matrix* handle_{nullptr};
};

}  // namespace math

// ---------------------------------------------------------------- type traits
/** True for matrix_cuda<T> (cf. prim/meta/is_matrix_cl.hpp). */
template <typename T>
struct is_matrix_cuda
    : bool_constant<std::is_base_of<math::matrix_cuda_base, std::decay_t<T>>::value> {
};
template <typename... Types>
using require_all_matrix_cuda_t = require_all_t<is_matrix_cuda<Types>...>;
template <typename T>
using require_matrix_cuda_t = require_t<is_matrix_cuda<T>>;
template <typename T>
using require_not_matrix_cuda_t = require_not_t<is_matrix_cuda<T>>;

// A device matrix is a non-scalar "kernel expression" in the sense of
// prim/meta/is_kernel_expression.hpp L33-41: that is the hook with which the
// reference switches its host implementations OFF for device operands
// (require_all_not_nonscalar_prim_or_rev_kernel_expression_t on every prim
// density, e.g. prim/prob/bernoulli_logit_lpmf.hpp L34-36), leaving the device
// overload as the only candidate.
template <typename T>
struct is_kernel_expression_and_not_scalar<T, require_matrix_cuda_t<T>>
    : std::true_type {};

// scalar_type / value_type of a device matrix is its element type, so that
// return_type_t, partials_return_t and include_summand treat matrix_cuda<double>
// as `double` data.
template <typename T>
struct scalar_type<T, require_matrix_cuda_t<T>> {
  using type = typename std::decay_t<T>::Scalar;
};
template <typename T>
struct value_type<T, require_matrix_cuda_t<T>> {
  using type = typename std::decay_t<T>::Scalar;
};
// a device matrix of data is a constant for autodiff (cf. opencl/is_constant.hpp)
template <typename T>
struct is_constant<T, require_matrix_cuda_t<T>> : std::true_type {};

}  // namespace stan
#endif
