// categorical_logit_glm_lpmf on the device (placeholder until the DMMA kernels land).
#include "smc_internal.h"
using namespace smc;
extern "C" int smc_categorical_logit_glm(const smc_matrix*, int, const smc_matrix*,
                                         const double*, const double*, int64_t,
                                         unsigned, double*, double*, double*,
                                         smc_matrix*) {
  return fail(SMC_ERR_UNSUPPORTED, "categorical_logit_glm_lpmf: not built yet");
}
