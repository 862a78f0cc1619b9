// One of the reference's own device tests of a GLM
// (test/unit/math/opencl/rev/<family>_test.cpp, named by -DREF_TEST_FILE), compiled
// UNMODIFIED against the CUDA backend.  The reference's headers are read first, without
// STAN_OPENCL, so nothing of the OpenCL backend is pulled in; the define that follows
// only opens the body of the test file, whose <stan/math/opencl/rev.hpp> resolves to
// tests/cpp/ref_shim/ (first on the include path).
#include <stan/math.hpp>
#include <stan/math/cuda.hpp>
#include <test/unit/math/expect_near_rel.hpp>
#include <test/unit/pretty_print_types.hpp>
#define STAN_OPENCL
#include REF_TEST_FILE
