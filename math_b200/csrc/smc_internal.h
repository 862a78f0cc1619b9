// Internal declarations shared by the translation units of libstanmath_cuda.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/stanmath_cuda.h"

namespace smc {

constexpr int kMaxFusedK = 256;   // columns one fused CTA tile can hold
constexpr int kColsPerThread = 32;
constexpr int kMaxParamDoubles = 256;  // by-value parameter block (beta)
constexpr int kMaxCuts = 512;
// internal flag bit (above the public SMC_* flags): un-fused density semantics
constexpr unsigned kFlagUnfused = 1u << 16;

enum Family {
  kNormal = 0,
  kBernoulli = 1,
  kPoisson = 2,
  kNegBinomial = 3,
  kOrdered = 4,
  kBinomial = 5,
  // Not a density: theta = x beta + alpha written out per row, and / or
  // x^T v for a per-row vector v (the matrix-vector products either side of an
  // un-fused log density, lpmf.cu).
  kLinear = 6
};

// Per-host-thread state: device, stream, reusable workspace.
struct Context {
  uint64_t id = 0;  // process-unique, assigned when the thread binds a device
  int device = -1;
  bool inited = false;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // own_stream or the user's
  int sm_count = 0;
  // workspace
  double* partials = nullptr;  // [grid][stride] per-CTA partial sums
  size_t partials_bytes = 0;
  unsigned int* counter = nullptr;  // last-block-done ticket (zero between launches)
  double* params_dev = nullptr;     // small device parameter staging
  size_t params_bytes = 0;
  double* out_host = nullptr;  // pinned + mapped result buffer
  size_t out_bytes = 0;
  double* scratch = nullptr;   // generic device scratch (N-vectors for fallbacks)
  size_t scratch_bytes = 0;
  int64_t launches = 0;
  cudaEvent_t timer_ev[2] = {nullptr, nullptr};  // smc_timer_start / smc_timer_stop
  unsigned long long sync_seq = 0;  // value the next polled completion flag takes
  bool flag_armed = false;          // the last launch will store its completion flag
  // Recycled device blocks, keyed by exact size.  An HMC run allocates the same
  // arena buffers (N-vector partials, the N x K d_x of an autodiff x) on every
  // evaluation: a freed block goes here and the next create of that size takes
  // it back without touching the driver.  Blocks are reused only by this
  // thread, on this thread's stream, so reuse is stream-ordered.
  std::unordered_map<size_t, std::vector<void*>> block_cache;
  size_t cached_bytes = 0;
  std::string last_error;
  ~Context();
};

Context& ctx();
Context& own_ctx();  // the calling thread's own context, whatever is current
// Makes `c` (a shard's context) the one this thread's launches go to; NULL restores
// the thread's own.  Returns the previous shard context (NULL: the thread's own).
Context* swap_current_context(Context* c);
int init_context(Context& c, int device);  // binds `c` to a device: stream + workspace
// The deferred memset of a lazily zeroed matrix; call before reading `m`.
int realize(const smc_matrix* m);
// Guards the caches that live ON a matrix (TMA descriptor, y statistics, binomial
// pair statistics): several host threads -- chains -- share one read-only x / y,
// and the first of them to need a cached item fills it in.
std::mutex& cache_mutex();
int ensure_ctx();  // lazily binds device 0 / creates the stream; returns smc_status
int fail(int status, const char* fmt, ...);
int ensure_partials(size_t bytes);
int ensure_params(size_t bytes);
int ensure_out(size_t bytes);
int ensure_scratch(size_t bytes);
int cache_alloc(void** p, size_t bytes);  // recycled block or cudaMalloc
void cache_free(void* p, size_t bytes);
void cache_trim();  // return every cached block to the driver

#define SMC_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess)                                                   \
      return ::smc::fail(SMC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,        \
                         cudaGetErrorString(e__), __FILE__, __LINE__);        \
  } while (0)

// Arguments of one evaluation of a memory-bound GLM family; filled by the C API.
struct GlmCall {
  int family = 0;
  const smc_matrix* x = nullptr;
  const smc_matrix* y = nullptr;
  double y_scalar = 0;
  const smc_matrix* alpha_vec = nullptr;
  double alpha = 0;
  const smc_matrix* aux_vec = nullptr;  // sigma / phi vector (f64); binomial: trials (i32)
  double aux = 0;                       // sigma / phi / trials scalar
  const double* beta_host = nullptr;    // K (by-value path)
  const double* cuts_host = nullptr;    // ncuts
  const double* params_dev = nullptr;   // beta[K] (+cuts) already on device
  int64_t ncuts = 0;
  unsigned flags = 0;
  // un-fused density on a device linear predictor (lpmf.cu): the constant terms
  // of a broadcast scalar y follow prim/prob/<family>_lpmf.hpp, not the GLM
  bool unfused = false;
  // row-sharded evaluation: false on every shard but the first, for the terms the
  // reference adds once per call rather than once per row (the poisson GLM's
  // lgamma(y + 1) of a broadcast scalar y)
  bool once_terms = true;
  double* out = nullptr;  // packed result, device-accessible
  unsigned long long* done_flag = nullptr;  // see FusedArgs::done_flag
  unsigned long long done_val = 0;
  // column chunk of a wider evaluation (see FusedArgs): 0 / 0 / false = the whole x
  int out_beta_off = 0, out_K_total = 0;
  bool out_skip_header = false;
  smc_matrix* d_alpha_vec = nullptr;
  smc_matrix* d_aux_vec = nullptr;
  smc_matrix* d_y_vec = nullptr;
  smc_matrix* d_x = nullptr;
};

// Launches the evaluation on ctx().stream; does not synchronise.
int launch_glm(const GlmCall& c);
// launch_glm with the packed result (n_out doubles) in the thread's pinned,
// device-mapped buffer, then a stream synchronise; *out points at the result.
int run_sync(GlmCall& c, int n_out, const double** out);
// True when the single-pass TMA kernel can take this x.
bool fused_supported(const smc_matrix* x);
// The same without the bound on the number of columns (wide x: column chunks).
bool fused_layout_ok(const smc_matrix* x);
int launch_glm_fused(const GlmCall& c);
// out (N x K, leading dimension ld) = d beta^T as a pure store stream (context.cu)
int launch_outer(double* out, int64_t ld, const double* d, int64_t N, int K,
                 const double* beta_host, const double* beta_dev);
int launch_glm_generic(const GlmCall& c);

// sum_i lgamma(y_i + 1) over an i32 device vector (cached on the matrix).
int y_lgamma_sum(const smc_matrix* y, double* sum);
// min / max of an i32 device vector (cached on the matrix).
int y_range(const smc_matrix* y, int* lo, int* hi);
// binomial_logit_glm_lpmf's data-only pieces over the successes n and the trials
// N (each a device vector or a broadcast scalar; `count` = number of pairs):
// *in_support = every 0 <= n_i <= N_i, *coef_sum = sum_i
// binomial_coefficient_log(N_i, n_i).  Cached on the n (else N) handle.
int binom_stats(const smc_matrix* n, int n_scalar, const smc_matrix* trials,
                int trials_scalar, int64_t count, bool* in_support,
                double* coef_sum);

// Once-initialised A/B switches (environment variables read on first use, never on
// the evaluation path afterwards).
struct Knobs {
  bool force_generic, dx_fused, cat_dx_fma, cat_no_tma;
  int cat_ks;
};
const Knobs& knobs();
// How long a synchronous call polls its completion flag before it blocks on the stream
// (SMC_SPIN_US, default 60 us).
int64_t spin_budget_us();
uint64_t next_matrix_id();

// ---- row-sharded matrices (sharded.cu) ------------------------------------------
int shard_count_of(const smc_matrix* m);  // 0: a plain matrix
bool same_partition(const smc_matrix* a, const smc_matrix* b);
// fn(g, shard, first global row) for every shard, with that shard's device, stream
// and workspace current; stops at the first non-zero status.
int for_each_shard(const smc_matrix* m,
                   const std::function<int(int, smc_matrix*, int64_t)>& fn);
int synchronize_shards();
// smc_timer_*: the same two events on every shard's stream (stop: max over shards)
int shards_timer_start();
int shards_timer_stop(double* ms_max, bool* any);
// One evaluation over the shards of c.x; *out = the reduced packed result (host).
int run_sharded(GlmCall& c, int n_out, const double** out);
// categorical_logit_glm_lpmf over the shards of x; params_host = [beta K x C, alpha C],
// *out = [logp, #non-finite, d_alpha[C], d_beta[K x C]] reduced, in host memory.
int run_sharded_categorical(const smc_matrix* y, int y_scalar, const smc_matrix* x,
                            const double* params_host, int64_t C, unsigned flags,
                            smc_matrix* d_x, const double** out);

int launch_categorical(const smc_matrix* y, int y_scalar, const smc_matrix* x,
                       const double* alpha_host, const double* beta_host,
                       int64_t C, unsigned flags, double* out_host_logp,
                       double* d_alpha, double* d_beta, smc_matrix* d_x);

}  // namespace smc

struct smc_matrix {
  // process-unique identity (a freed block is recycled at the same address and a
  // fresh handle restarts at version 1: caches that involve a second matrix key on
  // (id, version), never on the data pointer)
  uint64_t id = 0;
  // the Context (host thread) whose block cache / stream the owned block belongs to
  uint64_t home_ctx = 0;
  void* data = nullptr;
  int64_t rows = 0, cols = 0, ld = 0;
  int dtype = SMC_F64;
  bool owned = false;
  size_t alloc_bytes = 0;  // size of the owned block (key of the block cache)
  int device = 0;
  // cached TMA descriptor for the fused kernel (depends on the tile shape)
  CUtensorMap tmap;
  int tmap_rows = 0, tmap_cols = 0;
  // cached sum_i lgamma(y_i + 1) for integer data vectors (y is data: the
  // value only changes on upload)
  bool lgamma_valid = false;
  double lgamma_sum = 0;
  // cached min / max of an integer data vector (eager y-range checks)
  bool range_valid = false;
  int imin = 0, imax = 0;
  // bumped by every call that writes the matrix: keys caches that involve a
  // second matrix (the binomial pair statistics below)
  uint64_t version = 0;
  // smc_matrix_zero_lazy: every element is zero by declaration, the memset has not
  // run (and never will if the next writer overwrites the whole matrix)
  bool zero_pending = false;
  // cached binom_stats() of (this, partner): valid while both versions match
  bool binom_valid = false;
  uint64_t binom_partner_id = 0;
  uint64_t binom_partner_version = 0, binom_self_version = 0;
  int binom_partner_scalar = 0;
  bool binom_in_support = false;
  double binom_coef_sum = 0;
  // cached group index of an i32 index vector (indexing.cu, reverse sweep with more
  // groups than the shared-memory accumulators hold): the rows stably sorted by
  // group and the start of every group, on the device; valid while grp_version ==
  // version and grp_G is the caller's group count.  Owned by the matrix.
  int* grp_perm = nullptr;
  int* grp_off = nullptr;
  size_t grp_perm_bytes = 0, grp_off_bytes = 0;
  int64_t grp_G = 0;
  uint64_t grp_version = 0;
  // A row-sharded matrix (sharded.cu): data is NULL and shard g -- a plain matrix on
  // GPU g -- holds rows [shard_row0[g], shard_row0[g] + shards[g]->rows).
  std::vector<smc_matrix*> shards;
  std::vector<int64_t> shard_row0;
};

namespace smc {
// A per-row operand (N x 1 or 1 x N) the kernels index as data[i]: contiguous.
// smc_matrix_create lays vectors out that way; a wrapped strided row is refused.
inline bool is_sharded(const smc_matrix* m) { return m && !m->shards.empty(); }
// Entry points that have no row-sharded form yet refuse sharded handles up front.
inline int refuse_sharded(const char* fn, std::initializer_list<const smc_matrix*> ms) {
  for (const smc_matrix* m : ms)
    if (is_sharded(m))
      return fail(SMC_ERR_UNSUPPORTED, "%s: row-sharded operands are not supported here",
                  fn);
  return SMC_OK;
}
inline bool vec_contiguous(const smc_matrix* m) {
  return m->cols <= 1 || m->rows * m->cols == 0 || (m->rows == 1 && m->ld == 1);
}
}  // namespace smc
