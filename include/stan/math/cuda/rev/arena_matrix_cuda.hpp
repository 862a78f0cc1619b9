#ifndef STAN_MATH_CUDA_REV_ARENA_MATRIX_CUDA_HPP
#define STAN_MATH_CUDA_REV_ARENA_MATRIX_CUDA_HPP
// arena_matrix_cuda<T>: a device matrix whose lifetime is that of the autodiff
// arena, the analogue of arena_matrix_cl (stan/math/opencl/rev/arena_matrix_cl.hpp
// L15-22).  The object itself is a trivially destructible handle (it is captured
// by value in reverse_pass_callback lambdas, which are never destroyed:
// rev/core/callback_vari.hpp L32-33); the device buffer is owned by a
// chainable_alloc on var_alloc_stack_ (rev/core/chainable_alloc.hpp L16-22) and
// is released by recover_memory() (rev/core/recover_memory.hpp).
#include <stan/math/cuda/matrix_cuda.hpp>
#include <stan/math/rev/core/chainable_alloc.hpp>

#include <utility>

namespace stan {
namespace math {

namespace internal {
class cuda_arena_owner : public chainable_alloc {
 public:
  explicit cuda_arena_owner(smc_matrix* h) : h_(h) {}
  ~cuda_arena_owner() override {
    if (h_) {
      smc_matrix_free(h_);
    }
  }

 private:
  smc_matrix* h_;
};
}  // namespace internal

template <typename T>
class arena_matrix_cuda : public matrix_cuda_base {
 public:
  using Scalar = T;
  arena_matrix_cuda() = default;

  /** rows x cols of zeros owned by the arena.  The zeros are declared, not written
   * (smc_matrix_zero_lazy): the adjoint of a device var is usually produced whole by
   * its one consumer, which then stores into it without a memset or a read. */
  arena_matrix_cuda(int64_t rows, int64_t cols) {
    matrix_cuda<T> m(rows, cols);
    if (m.handle()) {
      check_cuda_status("arena_matrix_cuda(zeros)", smc_matrix_zero_lazy(m.handle()));
    }
    take(std::move(m));
  }
  /** Zeros (declared lazily) with the rows and the row partition of `like`. */
  static arena_matrix_cuda zeros_like(const smc_matrix* like, int64_t rows, int64_t cols) {
    arena_matrix_cuda a;
    matrix_cuda<T> m = matrix_cuda<T>::like_handle(like, rows, cols);
    if (m.handle()) {
      check_cuda_status("arena_matrix_cuda(zeros)", smc_matrix_zero_lazy(m.handle()));
    }
    a.take(std::move(m));
    return a;
  }
  /** Uninitialised, with the rows and the row partition of `like`. */
  static arena_matrix_cuda uninitialized_like(const smc_matrix* like, int64_t rows,
                                              int64_t cols) {
    arena_matrix_cuda a;
    a.take(matrix_cuda<T>::like_handle(like, rows, cols));
    return a;
  }
  /** rows x cols owned by the arena, contents unspecified (the producer
   * overwrites every element). */
  static arena_matrix_cuda uninitialized(int64_t rows, int64_t cols) {
    arena_matrix_cuda a;
    a.take(matrix_cuda<T>(rows, cols));
    return a;
  }
  /** Moves an owning matrix into the arena. */
  explicit arena_matrix_cuda(matrix_cuda<T>&& m) { take(std::move(m)); }
  /** Arena-owned view: `m` must outlive the reverse sweep (data matrices do). */
  static arena_matrix_cuda view(const matrix_cuda<T>& m) {
    arena_matrix_cuda a;
    a.take(matrix_cuda<T>::view(m));
    return a;
  }

  int64_t rows() const noexcept { return h_ ? smc_matrix_rows(h_) : 0; }
  int64_t cols() const noexcept { return h_ ? smc_matrix_cols(h_) : 0; }
  int64_t size() const noexcept { return rows() * cols(); }
  smc_matrix* handle() const noexcept { return h_; }

  /** Overwrites the contents with those of a device matrix of the same shape
   * (`x.adj() = expr` of the OpenCL backend, opencl/rev/arena_matrix_cl.hpp L60-75). */
  arena_matrix_cuda& operator=(const matrix_cuda<T>& m) {
    check_size_match("arena_matrix_cuda::operator=", "rows", rows(), "rows of the source",
                     m.rows());
    check_size_match("arena_matrix_cuda::operator=", "cols", cols(), "cols of the source",
                     m.cols());
    if (h_ && size() > 0) {
      check_cuda_status("arena_matrix_cuda::operator=", smc_matrix_copy(h_, m.handle()));
    }
    return *this;
  }
  /** this += m on the device (`a.adj() += to_matrix_cl(...)`, opencl/rev/copy.hpp L118-125). */
  arena_matrix_cuda& operator+=(const matrix_cuda<T>& m) {
    check_size_match("arena_matrix_cuda::operator+=", "rows", rows(),
                     "rows of the source", m.rows());
    check_size_match("arena_matrix_cuda::operator+=", "cols", cols(),
                     "cols of the source", m.cols());
    if (h_ && size() > 0) {
      check_cuda_status("arena_matrix_cuda::operator+=",
                        smc_matrix_axpy(h_, 1.0, m.handle()));
    }
    return *this;
  }

  /** Copy to an owning matrix (device-to-device). */
  matrix_cuda<T> to_matrix_cuda() const {
    matrix_cuda<T> out = matrix_cuda<T>::like_handle(h_, rows(), cols());
    if (h_) {
      check_cuda_status("arena_matrix_cuda::to_matrix_cuda",
                        smc_matrix_copy(out.handle(), h_));
    }
    return out;
  }

 private:
  void take(matrix_cuda<T>&& m) {
    h_ = m.release_handle();
    if (h_) {
      new internal::cuda_arena_owner(h_);  // registered on var_alloc_stack_
    }
  }
  smc_matrix* h_{nullptr};
};

}  // namespace math
}  // namespace stan
#endif
