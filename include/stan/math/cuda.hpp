#ifndef STAN_MATH_CUDA_HPP
#define STAN_MATH_CUDA_HPP
// Umbrella header of the B200 (CUDA, sm_100a) backend for the GLM hot path.
// Include after <stan/math.hpp> (or let stan/math/rev.hpp pull it in under
// `#ifdef STAN_CUDA`, the way STAN_OPENCL hooks stan/math/opencl/rev.hpp in
// stan/math/rev.hpp L6-8 -- see INTEGRATION.md).  Link with -lstanmath_cuda.
#include <stan/math/rev.hpp>

#include <stan/math/cuda/err.hpp>
#include <stan/math/cuda/matrix_cuda.hpp>
#include <stan/math/cuda/copy.hpp>
#include <stan/math/cuda/rev/arena_matrix_cuda.hpp>
#include <stan/math/cuda/rev/vari.hpp>
#include <stan/math/cuda/rev/copy.hpp>
#include <stan/math/cuda/rev/operands_and_partials.hpp>
#include <stan/math/cuda/prim/bernoulli_logit_glm_lpmf.hpp>
#include <stan/math/cuda/prim/binomial_logit_glm_lpmf.hpp>
#include <stan/math/cuda/prim/poisson_log_glm_lpmf.hpp>
#include <stan/math/cuda/prim/normal_id_glm_lpdf.hpp>
#include <stan/math/cuda/prim/neg_binomial_2_log_glm_lpmf.hpp>
#include <stan/math/cuda/prim/ordered_logistic_glm_lpmf.hpp>
#include <stan/math/cuda/prim/categorical_logit_glm_lpmf.hpp>
#include <stan/math/cuda/rev/multiply.hpp>
#include <stan/math/cuda/rev/indexing.hpp>
#include <stan/math/cuda/prim/unfused_lpmf.hpp>
#include <stan/math/cuda/prim/bernoulli_logit_glm_rng.hpp>

#endif
