"""math_b200: B200-native (sm_100a, FP64) backend for Stan Math's GLM
log-density + gradient hot path.  See DESIGN.md and include/stanmath_cuda.h."""
from ._lib import (DomainError, BackendError, PROPTO, VAR_X, VAR_ALPHA, VAR_BETA,
                   VAR_AUX, VAR_Y, LIB_PATH)
from .matrix_cuda import (MatrixCuda, to_matrix_cuda, from_matrix_cuda,
                          synthetic_host)
from .glm import (GlmResult, bernoulli_logit_glm_lpmf, poisson_log_glm_lpmf,
                  normal_id_glm_lpdf, neg_binomial_2_log_glm_lpmf,
                  ordered_logistic_glm_lpmf, categorical_logit_glm_lpmf,
                  binomial_logit_glm_lpmf)
from . import lpmf
from . import runtime
