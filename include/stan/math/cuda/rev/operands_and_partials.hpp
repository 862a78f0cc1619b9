#ifndef STAN_MATH_CUDA_REV_OPERANDS_AND_PARTIALS_HPP
#define STAN_MATH_CUDA_REV_OPERANDS_AND_PARTIALS_HPP
// Partials edge for a device-resident autodiff operand, the analogue of
// ops_partials_edge<double, var_value<Op>, require_kernel_expression_lhs_t<Op>>
// (stan/math/opencl/rev/operands_and_partials.hpp L16-31) and of the OpenCL
// update_adjoints overload (rev/functor/operands_and_partials.hpp L28-38).
//
// The partial of a device operand takes one of two forms, chosen by the function
// that fills the edge:
//   full      an arena-owned device matrix with the shape of the operand (as every
//             host edge is, rev/functor/operands_and_partials.hpp L100-184); the
//             kernel writes EVERY element of it, so it is not zero-filled first.
//             Reverse sweep: one device axpy  x.adj() += ret.adj() * partial.
//   factored  the partial of a design matrix in a GLM is the rank-one product
//             d beta^T (prim/prob/bernoulli_logit_glm_lpmf.hpp L158-159 and its
//             siblings): the edge keeps the N-vector d on the device and beta in the
//             arena, and the reverse sweep is ONE pass over the adjoint,
//             x.adj() += ret.adj() * d beta^T (smc_matrix_rank1_update) -- the
//             N x K product is never written, read back or allocated.  With the
//             adjoint still lazily zero (the usual case: one GLM consumes x) that
//             pass is a pure store.
// Nothing is allocated until the function asks for one of the two.
#include <stan/math/cuda/rev/vari.hpp>
#include <stan/math/prim/functor/partials_propagator.hpp>
#include <stan/math/rev/core/chainablestack.hpp>
#include <stan/math/rev/functor/operands_and_partials.hpp>
#include <stan/math/rev/functor/partials_propagator.hpp>

#include <algorithm>

namespace stan {
namespace math {

/** The partial of a device operand: trivially destructible (captured by value in
 * the reverse-pass callback), buffers owned by the arena. */
class cuda_edge_partial {
 public:
  cuda_edge_partial() = default;
  /** `like`: the operand's value handle; the partial follows its row partition. */
  cuda_edge_partial(const smc_matrix* like, int64_t rows, int64_t cols)
      : like_(like), rows_(rows), cols_(cols) {}

  /** The full rows x cols partial (allocated on first use, contents unspecified). */
  smc_matrix* handle() {
    if (!full_.handle()) {
      full_ = arena_matrix_cuda<double>::uninitialized_like(like_, rows_, cols_);
    }
    return full_.handle();
  }
  /** The rank-one form: returns the rows x 1 device vector the kernel writes d
   * into; beta (cols doubles) is copied into the arena. */
  smc_matrix* factored(const double* beta) {
    d_ = arena_matrix_cuda<double>::uninitialized_like(like_, rows_, 1);
    beta_ = ChainableStack::instance_->memalloc_.alloc_array<double>(
        static_cast<size_t>(cols_ > 0 ? cols_ : 1));
    std::copy(beta, beta + cols_, beta_);
    return d_.handle();
  }
  bool is_factored() const noexcept { return beta_ != nullptr; }
  bool is_full() const noexcept { return full_.handle() != nullptr; }
  const arena_matrix_cuda<double>& full() const noexcept { return full_; }
  const arena_matrix_cuda<double>& factor() const noexcept { return d_; }
  const double* beta() const noexcept { return beta_; }
  int64_t rows() const noexcept { return rows_; }
  int64_t cols() const noexcept { return cols_; }

 private:
  arena_matrix_cuda<double> full_;
  arena_matrix_cuda<double> d_;
  double* beta_{nullptr};
  const smc_matrix* like_{nullptr};
  int64_t rows_{0}, cols_{0};
};

inline void update_adjoints(var_value<matrix_cuda<double>>& x,
                            const arena_matrix_cuda<double>& y, const vari& z) {
  if (x.size() > 0) {
    check_cuda_status("update_adjoints(matrix_cuda)",
                      smc_matrix_axpy(x.adj().handle(), z.adj(), y.handle()));
  }
}
inline void update_adjoints(var_value<matrix_cuda<double>>& x,
                            const arena_matrix_cuda<double>& y, const var& z) {
  update_adjoints(x, y, *z.vi_);
}
inline void update_adjoints(var_value<matrix_cuda<double>>& x,
                            const cuda_edge_partial& y, const vari& z) {
  if (x.size() == 0) {
    return;
  }
  if (y.is_factored()) {
    check_cuda_status("update_adjoints(matrix_cuda, d beta^T)",
                      smc_matrix_rank1_update(x.adj().handle(), z.adj(),
                                              y.factor().handle(), y.beta()));
  }
  if (y.is_full()) {
    update_adjoints(x, y.full(), z);
  }
}
inline void update_adjoints(var_value<matrix_cuda<double>>& x,
                            const cuda_edge_partial& y, const var& z) {
  update_adjoints(x, y, *z.vi_);
}

namespace internal {

template <>
class ops_partials_edge<double, var_value<matrix_cuda<double>>, void> {
 public:
  using partials_t = cuda_edge_partial;
  partials_t partials_;
  broadcast_array<partials_t> partials_vec_;
  explicit ops_partials_edge(const var_value<matrix_cuda<double>>& ops)
      : partials_(ops.val().handle(), ops.rows(), ops.cols()),
        partials_vec_(partials_),
        operands_(ops) {}
  inline auto& partial() noexcept { return partials_; }
  inline auto& operand() const noexcept { return operands_; }
  var_value<matrix_cuda<double>> operands_;
  static constexpr int size() { return 0; }
};

}  // namespace internal
}  // namespace math
}  // namespace stan
#endif
