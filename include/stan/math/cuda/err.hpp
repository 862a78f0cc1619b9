#ifndef STAN_MATH_CUDA_ERR_HPP
#define STAN_MATH_CUDA_ERR_HPP
// Maps the C ABI's status codes onto the exception types the reference throws
// on this path (SURVEY.md 8(b) "Error conventions"):
//   SMC_ERR_INVALID_ARGUMENT -> std::invalid_argument  (check_consistent_size /
//                                check_size_match, prim/err/check_size_match.hpp)
//   SMC_ERR_DOMAIN           -> std::domain_error      (check_bounded, check_finite ...)
//   SMC_ERR_CUDA / other     -> std::system_error      (cf. check_opencl_error,
//                                stan/math/opencl/err/check_opencl.hpp)
// The reference's tests pin the exception type only, never the message text.
#include <stanmath_cuda.h>

#include <stdexcept>
#include <string>
#include <system_error>

namespace stan {
namespace math {

/** Error category of backend failures (the codes are smc_status values, not errno). */
class cuda_backend_category_t : public std::error_category {
 public:
  const char* name() const noexcept override { return "stanmath_cuda"; }
  std::string message(int status) const override {
    switch (status) {
      case SMC_ERR_CUDA:
        return "CUDA backend failure";
      case SMC_ERR_UNSUPPORTED:
        return "not supported by the CUDA backend";
      default:
        return "stanmath_cuda status " + std::to_string(status);
    }
  }
};
inline const std::error_category& cuda_backend_category() {
  static const cuda_backend_category_t c;
  return c;
}

inline void check_cuda_status(const char* function, int status) {
  if (status == SMC_OK) {
    return;
  }
  const char* txt = smc_last_error();
  std::string msg = std::string(function) + ": " + (txt ? txt : "");
  switch (status) {
    case SMC_ERR_INVALID_ARGUMENT:
      throw std::invalid_argument(msg);
    case SMC_ERR_DOMAIN:
      throw std::domain_error(msg);
    default:
      throw std::system_error(status, cuda_backend_category(), msg);
  }
}

}  // namespace math
}  // namespace stan
#endif
