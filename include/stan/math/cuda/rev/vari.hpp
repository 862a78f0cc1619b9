#ifndef STAN_MATH_CUDA_REV_VARI_HPP
#define STAN_MATH_CUDA_REV_VARI_HPP
// var_value<matrix_cuda<double>>: an autodiff matrix whose value AND adjoint
// live in HBM, the analogue of var_value<matrix_cl<double>> /
// vari_value<matrix_cl<T>> (stan/math/opencl/rev/vari.hpp L14, L178-290).  This
// is the type the design matrix takes when x is an autodiff variable
// (BASELINE config 4: a host Matrix<var> with 1.28e9 varis is infeasible, the
// N x K adjoint stays on the device).
//
//   - the vari sits on var_nochain_stack_ (its chain() is empty: producers
//     propagate into adj_ through update_adjoints),
//   - set_zero_adjoint() re-zeros adj_ on the device (opencl/rev/vari.hpp L287);
//     zeros are declared lazily (smc_matrix_zero_lazy): the first producer to
//     propagate into the adjoint stores its contribution instead of adding to a
//     freshly written block of zeros,
//   - value and adjoint buffers are arena-owned (freed by recover_memory()).
#include <stan/math/cuda/rev/arena_matrix_cuda.hpp>
#include <stan/math/rev/core/var.hpp>
#include <stan/math/rev/core/vari.hpp>

#include <utility>

namespace stan {
namespace math {

template <>
class vari_value<matrix_cuda<double>, void> : public vari_base {
 public:
  using value_type = matrix_cuda<double>;
  arena_matrix_cuda<double> val_;
  arena_matrix_cuda<double> adj_;

  /** Takes ownership of the value buffer. */
  explicit vari_value(matrix_cuda<double>&& v)
      : val_(std::move(v)),
        adj_(arena_matrix_cuda<double>::zeros_like(val_.handle(), val_.rows(),
                                                   val_.cols())) {
    ChainableStack::instance_->var_nochain_stack_.push_back(this);
  }
  /** The constructor rev/core/callback_vari.hpp L16-19 calls, so that the reference's
   * own make_callback_var(device value, functor) yields a device vari whose chain()
   * runs the functor: `stacked` puts it on the chaining stack. */
  vari_value(matrix_cuda<double>&& v, bool stacked)
      : val_(std::move(v)),
        adj_(arena_matrix_cuda<double>::zeros_like(val_.handle(), val_.rows(),
                                                   val_.cols())) {
    if (stacked) {
      ChainableStack::instance_->var_stack_.push_back(this);
    } else {
      ChainableStack::instance_->var_nochain_stack_.push_back(this);
    }
  }
  /** Views the value buffer: `v` must outlive the reverse sweep. */
  explicit vari_value(const matrix_cuda<double>& v)
      : val_(arena_matrix_cuda<double>::view(v)),
        adj_(arena_matrix_cuda<double>::zeros_like(val_.handle(), val_.rows(),
                                                   val_.cols())) {
    ChainableStack::instance_->var_nochain_stack_.push_back(this);
  }

  inline const arena_matrix_cuda<double>& val() const noexcept { return val_; }
  inline arena_matrix_cuda<double>& adj() noexcept { return adj_; }
  inline const arena_matrix_cuda<double>& adj() const noexcept { return adj_; }
  inline int64_t rows() const noexcept { return val_.rows(); }
  inline int64_t cols() const noexcept { return val_.cols(); }
  inline int64_t size() const noexcept { return val_.size(); }

  void chain() override {}
  void set_zero_adjoint() override {
    if (adj_.handle()) {
      check_cuda_status("vari_value<matrix_cuda>::set_zero_adjoint",
                        smc_matrix_zero_lazy(adj_.handle()));
    }
  }
};

template <>
class var_value<matrix_cuda<double>, void> {
 public:
  using value_type = matrix_cuda<double>;
  using vari_type = vari_value<matrix_cuda<double>>;
  vari_type* vi_;

  var_value() : vi_(nullptr) {}
  var_value(vari_type* vi) : vi_(vi) {}  // NOLINT
  /** Moves the device buffer onto the tape (no copy). */
  var_value(matrix_cuda<double>&& v) : vi_(new vari_type(std::move(v))) {}  // NOLINT
  /** Zero-copy view of a device matrix that outlives the sweep. */
  var_value(const matrix_cuda<double>& v) : vi_(new vari_type(v)) {}  // NOLINT

  inline bool is_uninitialized() noexcept { return vi_ == nullptr; }
  inline const arena_matrix_cuda<double>& val() const noexcept { return vi_->val_; }
  inline arena_matrix_cuda<double>& adj() const noexcept { return vi_->adj_; }
  inline int64_t rows() const noexcept { return vi_->rows(); }
  inline int64_t cols() const noexcept { return vi_->cols(); }
  inline int64_t size() const noexcept { return vi_->size(); }
  inline vari_type& operator*() { return *vi_; }
  inline vari_type* operator->() { return vi_; }
};

}  // namespace math

template <typename T>
struct is_var_matrix_cuda
    : std::is_same<std::decay_t<T>, math::var_value<math::matrix_cuda<double>>> {};

}  // namespace stan
#endif
