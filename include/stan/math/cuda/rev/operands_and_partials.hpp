#ifndef STAN_MATH_CUDA_REV_OPERANDS_AND_PARTIALS_HPP
#define STAN_MATH_CUDA_REV_OPERANDS_AND_PARTIALS_HPP
// Partials edge for a device-resident autodiff operand, the analogue of
// ops_partials_edge<double, var_value<Op>, require_kernel_expression_lhs_t<Op>>
// (stan/math/opencl/rev/operands_and_partials.hpp L16-31) and of the OpenCL
// update_adjoints overload (rev/functor/operands_and_partials.hpp L28-38).
//
// The partial is an arena-owned device matrix with the shape of the operand (as
// every host edge is, rev/functor/operands_and_partials.hpp L100-184); the GLM
// kernel writes EVERY element of it directly (so, unlike the host edges, it is
// not zero-filled first: at N x K = 1e7 x 128 that memset alone would cost a
// quarter of the evaluation), and the reverse sweep runs one device axpy:
//   x.adj() += ret.adj() * partial.
#include <stan/math/cuda/rev/vari.hpp>
#include <stan/math/prim/functor/partials_propagator.hpp>
#include <stan/math/rev/functor/operands_and_partials.hpp>
#include <stan/math/rev/functor/partials_propagator.hpp>

namespace stan {
namespace math {

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

namespace internal {

template <>
class ops_partials_edge<double, var_value<matrix_cuda<double>>, void> {
 public:
  using partials_t = arena_matrix_cuda<double>;
  partials_t partials_;
  broadcast_array<partials_t> partials_vec_;
  explicit ops_partials_edge(const var_value<matrix_cuda<double>>& ops)
      : partials_(arena_matrix_cuda<double>::uninitialized(ops.rows(), ops.cols())),
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
