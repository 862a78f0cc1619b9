/* stanmath_cuda.h -- C ABI of libstanmath_cuda.so, the B200 (sm_100a) backend for
 * Stan Math's GLM log-density + gradient hot path.
 *
 * This is the boundary a `stan/math/cuda` backend binds, the slot the
 * reference's OpenCL backend occupies today:
 *   - smc_matrix            <-> matrix_cl<T>        stan/math/opencl/matrix_cl.hpp L46-55
 *   - smc_matrix_upload/... <-> to_matrix_cl / from_matrix_cl   stan/math/opencl/copy.hpp L45, L61-235
 *   - smc_<family>_glm      <-> the matrix_cl overloads of the same names in
 *                               stan/math/opencl/prim/<family>_glm_l{pdf,pmf}.hpp
 *                               (value + partials of prim/prob/<family>_glm_*.hpp)
 * The C++ header overloads that call it live in include/stan/math/cuda/.
 *
 * Conventions
 *   - Plain C: opaque handles, pointers and sizes; no exceptions cross the ABI.
 *   - Matrices are column-major (Eigen / matrix_cl default), dtype f64 or i32.
 *   - Every call returns smc_status; smc_last_error() gives the thread-local text.
 *     SMC_ERR_INVALID_ARGUMENT <-> std::invalid_argument (sizes),
 *     SMC_ERR_DOMAIN           <-> std::domain_error (values),
 *     SMC_ERR_CUDA             <-> std::system_error (backend failure).
 *   - Re-entrant: each host thread owns its device, stream and workspace, so
 *     several chains may evaluate against the same read-only device x.
 *   - Caller owns host buffers; the library owns device buffers behind handles.
 *   - There is no CPU fallback: without a CUDA device every compute call fails
 *     with SMC_ERR_CUDA.
 *
 * `flags` mirror the reference's compile-time switches
 * (include_summand<propto, ...>, is_constant_all<...>): a term or a partial is
 * computed only when the reference would compute it.
 */
#ifndef STANMATH_CUDA_H
#define STANMATH_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smc_matrix smc_matrix;

enum smc_status {
  SMC_OK = 0,
  SMC_ERR_INVALID_ARGUMENT = 1,
  SMC_ERR_DOMAIN = 2,
  SMC_ERR_CUDA = 3,
  SMC_ERR_UNSUPPORTED = 4
};

enum smc_dtype { SMC_F64 = 0, SMC_I32 = 1 };

/* flags */
#define SMC_PROPTO 1u     /* propto = true                                   */
#define SMC_VAR_X 2u      /* x is an autodiff variable -> d_x is produced    */
#define SMC_VAR_ALPHA 4u  /* alpha is a var                                  */
#define SMC_VAR_BETA 8u   /* beta is a var                                   */
#define SMC_VAR_AUX 16u   /* sigma / phi / cuts is a var                     */
#define SMC_VAR_Y 32u     /* y is a var (normal_id_glm only)                 */
/* With SMC_VAR_X: d_x receives the N x 1 factor d of the rank-one partial
 * d_x = d beta^T instead of the N x K product.  The caller keeps beta and runs its
 * reverse sweep x.adj += lp.adj * d beta^T with smc_matrix_rank1_update, so the
 * product is never materialised (memory-bound families; the categorical partial
 * T beta^T has rank C and ignores this flag). */
#define SMC_DX_FACTORED 64u

/* Layout of the packed result vector written by the *_device entry points
 * (and all-reduced across GPUs by the row-sharded driver):
 *   [0] log density   [1] sum_i d_i   [2] family aux sum   [3] #rows with a
 *   non-finite linear predictor   [4] second aux sum   [5..7] reserved
 *   [8 .. 8+K)            d_beta
 *   [8+K .. 8+K+ncuts)    d_cuts (ordered_logistic only)                    */
#define SMC_OUT_HEADER 8
#define SMC_OUT_LOGP 0
#define SMC_OUT_SUM_D 1
#define SMC_OUT_AUX 2
#define SMC_OUT_NONFINITE 3
#define SMC_OUT_AUX2 4

/* ---- runtime ------------------------------------------------------------- */
int smc_device_count(int* count);
int smc_set_device(int device);        /* for the calling thread               */
int smc_get_device(int* device);
/* Use an existing CUDA stream (cudaStream_t) for the calling thread's launches;
 * NULL restores the library's own non-blocking stream. */
int smc_set_stream(void* cuda_stream);
int smc_synchronize(void);
/* Device blocks of freed matrices are kept by the calling thread for reuse (an
 * HMC run re-creates the same arena buffers every evaluation); this returns
 * them to the driver.  Done automatically when an allocation fails. */
int smc_trim_cache(void);
int smc_device_info(int* sm_count, int* cc_major, int* cc_minor,
                    size_t* free_bytes, size_t* total_bytes);
/* Device-side timing (CUDA events on the calling thread's stream, and on every
 * shard's stream: the maximum is reported) of the work queued between the calls. */
int smc_timer_start(void);
int smc_timer_stop(double* ms);
/* The FP64 tensor-core (DMMA) rate this GPU sustains from registers, in TFLOP/s: the
 * measured ceiling the categorical GLM's contraction rate is quoted against. */
int smc_measure_dmma_peak(double* tflops);
const char* smc_last_error(void);
/* Number of GLM kernel launches issued by the calling thread since the last
 * smc_reset_launch_count (bench.py's gpu_launches). */
int64_t smc_launch_count(void);
void smc_reset_launch_count(void);

/* ---- device matrix (matrix_cl analogue) ---------------------------------- */
/* rows x cols, column-major; the library pads the leading dimension so every
 * column starts 128-byte aligned (TMA requirement and full-sector DRAM reads). */
int smc_matrix_create(int64_t rows, int64_t cols, int dtype, smc_matrix** out);
/* Wrap an existing device buffer (not owned), cf. matrix_cl.hpp L190-192. */
int smc_matrix_wrap(void* device_ptr, int64_t rows, int64_t cols, int64_t ld,
                    int dtype, smc_matrix** out);
int smc_matrix_free(smc_matrix* m);
int64_t smc_matrix_rows(const smc_matrix* m);
int64_t smc_matrix_cols(const smc_matrix* m);
int64_t smc_matrix_ld(const smc_matrix* m);
int smc_matrix_dtype(const smc_matrix* m);
void* smc_matrix_data(const smc_matrix* m);
/* Host <-> device copies; ld_host is the host leading dimension in elements. */
int smc_matrix_upload(smc_matrix* m, const void* host, int64_t ld_host);
/* Upload the row block [row0, row0+nrows) from a host matrix whose element
 * (row0, 0) is at `host` (row-sharding: one call per GPU, ld_host = N). */
int smc_matrix_upload_rows(smc_matrix* m, int64_t row0, int64_t nrows,
                           const void* host, int64_t ld_host);
int smc_matrix_download(const smc_matrix* m, void* host, int64_t ld_host);
int smc_matrix_download_rows(const smc_matrix* m, int64_t row0, int64_t nrows,
                             void* host, int64_t ld_host);
int smc_matrix_zero(smc_matrix* m);
/* Declares every element zero WITHOUT touching memory (the adjoint of a device var,
 * opencl/rev/vari.hpp L287 `adj_ = constant(0, ...)`): the memset is deferred until a
 * library call reads the matrix or writes part of it (including smc_matrix_data,
 * which hands out the raw pointer), and never runs when the next writer overwrites
 * or accumulates into the whole matrix (smc_matrix_rank1_update, smc_matrix_axpy,
 * uploads, copies): those store their result directly. */
int smc_matrix_zero_lazy(smc_matrix* m);
/* For the owner of WRAPPED memory (smc_matrix_wrap) that changed the contents behind
 * the library's back: drops what the handle caches about them (range and lgamma
 * sums of an integer vector, binomial pair statistics). */
int smc_matrix_invalidate(smc_matrix* m);
/* Device-to-device copy of a same-shaped matrix (matrix_cl copy construction,
 * matrix_cl.hpp L198-210); asynchronous on the thread's stream. */
int smc_matrix_copy(smc_matrix* dst, const smc_matrix* src);
/* out[i, k] = beta[k] * d[i] for an N x 1 f64 device vector d and K host doubles
 * (K <= 256): the N x K partial beta (x) d of an autodiff design matrix written as a
 * pure store stream (16-byte stores when `out` is 16-byte aligned with an even leading
 * dimension -- what smc_matrix_create lays out -- else element by element). */
int smc_matrix_outer(smc_matrix* out, const smc_matrix* d, const double* beta);
/* y += a * x on the device: update_adjoints for a device-resident operand
 * (rev/functor/operands_and_partials.hpp L28-38). */
int smc_matrix_axpy(smc_matrix* y, double a, const smc_matrix* x);
/* y[i,k] += a * (d[i] * beta[k]) for an N x K f64 device matrix y, an N x 1 f64
 * device vector d and K host doubles: the reverse sweep of an autodiff design
 * matrix, x.adj() += lp.adj() * partial with partial = d beta^T
 * (rev/functor/operands_and_partials.hpp L28-38;
 * prim/prob/neg_binomial_2_log_glm_lpmf.hpp L221-222) as ONE read-modify-write of the
 * adjoint -- or one pure store when y is lazily zero -- without ever forming the
 * N x K partial. */
int smc_matrix_rank1_update(smc_matrix* y, double a, const smc_matrix* d,
                            const double* beta);
/* Lazy value checks (cf. check_cl in opencl/prim/ *_glm_*.hpp). */
/* y += a (every element) */
int smc_matrix_add_scalar(smc_matrix* y, double a);
int smc_matrix_all_finite(const smc_matrix* m, int* all_finite);
int smc_matrix_int_range(const smc_matrix* m, int* min_out, int* max_out);
/* Deterministic synthetic fill, identical on every GPU and reproducible on the
 * host (counter-based hash of (seed, row0+i, k); exact integer arithmetic and a
 * single correctly-rounded multiply, so host and device agree bit for bit):
 *   kind 0: f64, zero mean, unit variance (sum of four 16-bit uniforms), *scale
 *   kind 1: i32 uniform in [lo, hi]                                          */
int smc_matrix_fill_synthetic(smc_matrix* m, uint64_t seed, int64_t row0,
                              int kind, double scale, int lo, int hi);

/* ---- row-sharded matrices: the GPUs of one box behind the same entry points -- */
/* A sharded smc_matrix partitions its rows contiguously over the shard set (shard g
 * holds rows [g N / G, (g+1) N / G) on its GPU; uploaded once).  It is accepted
 * wherever the GLM entry points below -- all seven families, smc_linear_predictor and
 * its adjoint -- and the matrix calls above (upload / download incl. row blocks, zero,
 * zero_lazy, copy, axpy, rank1_update, outer, add_scalar, all_finite, int_range,
 * fill_synthetic, free) take a plain one, provided x and every per-row operand of a
 * call are sharded alike (smc_matrix_create_like).  One evaluation launches the fused
 * kernel on every GPU (small parameters travel as kernel arguments: that is the
 * broadcast) and sums the packed results: K + O(1) doubles per GPU go straight from
 * every kernel into its slot of one pinned host buffer, a completion flag behind them,
 * and are added in shard order by the calling thread ("direct": no collective, no copy,
 * no stream synchronise; bit-reproducible); larger results (the categorical K x C
 * gradient) are all-reduced in place over NCCL on the compute streams and read back
 * once.  N-vector partials and the N x K adjoint of an autodiff x stay sharded.  This is the reference's scatter-once / broadcast-
 * parameters / reduce pattern (prim/functor/mpi_parallel_call.hpp L332-392,
 * L408-449) in one process, beyond the one device of the OpenCL backend
 * (opencl/opencl_context.hpp L75-76).  Entry points without a sharded form
 * (un-fused densities, indexing, the matrix product with a K x C weight matrix,
 * smc_matrix_wrap / smc_matrix_data) return SMC_ERR_UNSUPPORTED or NULL.
 *
 * smc_shard_init: n_shards <= 0 takes every visible GPU; `devices` (n_shards ids)
 * or NULL for 0, 1, ...  NCCL is loaded at run time (libnccl.so.2).  If it is
 * missing, if SMC_SHARD_REDUCE=host is set, or if several shards share a GPU
 * (single-GPU tests of the sharded logic), the packed results are read back per
 * shard and summed on the host in shard order instead.  SMC_SHARD_REDUCE=nccl sends
 * the small results through the NCCL all-reduce too (smc_shard_reduce_mode() tells:
 * "direct+nccl", "direct+host", "nccl", "host").  Re-initialising invalidates
 * existing sharded matrices.  One sharded evaluation runs at a time per process. */
int smc_shard_init(int n_shards, const int* devices);
int smc_shard_count(int* n_shards);
const char* smc_shard_reduce_mode(void); /* "nccl" | "host" | "" before init */
int smc_shard_synchronize(void);
int smc_shard_shutdown(void);
int smc_sharded_matrix_create(int64_t rows, int64_t cols, int dtype, smc_matrix** out);
/* A matrix with the rows -- and the row partition -- of `like`: `cols` columns
 * (< 0: as many as `like`) of `dtype` (< 0: the same).  Plain when `like` is plain. */
int smc_matrix_create_like(const smc_matrix* like, int64_t cols, int dtype,
                           smc_matrix** out);
/* A non-owning alias of `src` (plain or sharded), which must outlive it. */
int smc_matrix_view(const smc_matrix* src, smc_matrix** out);
int smc_matrix_shard_count(const smc_matrix* m); /* 0: a plain matrix */
/* Shard g of a sharded matrix (borrowed: owned by `m`), its first global row and GPU. */
int smc_matrix_shard(const smc_matrix* m, int g, smc_matrix** shard, int64_t* row0,
                     int* device);

/* ---- GLM log density + gradient ------------------------------------------ */
/* Common argument meaning (N = x rows, K = x cols):
 *   y            N x 1 device vector (i32; f64 for normal_id) or NULL to
 *                broadcast the scalar y_scalar
 *   x            N x K f64 device matrix (uploaded once, reused every call)
 *   alpha_vec    N x 1 f64 device vector, or NULL -> scalar `alpha`
 *   beta         host pointer, K doubles
 *   logp         host out: the log density the reference returns
 *   d_alpha      host out (1 double): sum_i d_i, the partial of a scalar alpha
 *   d_alpha_vec  device out N x 1: the partial of a vector alpha (or NULL)
 *   d_beta       host out, K doubles
 *   d_x          device out N x K (required when SMC_VAR_X is set)
 * Outputs whose flag is not set are not written and may be NULL.             */

/* prim/prob/bernoulli_logit_glm_lpmf.hpp L49-167 */
int smc_bernoulli_logit_glm(const smc_matrix* y, int y_scalar,
                            const smc_matrix* x, const smc_matrix* alpha_vec,
                            double alpha, const double* beta, unsigned flags,
                            double* logp, double* d_alpha,
                            smc_matrix* d_alpha_vec, double* d_beta,
                            smc_matrix* d_x);

/* prim/prob/binomial_logit_glm_lpmf.hpp L54-160 (device twin:
 * opencl/prim/binomial_logit_glm_lpmf.hpp L28-139)
 *   n       N x 1 i32 device vector of successes, or NULL -> n_scalar
 *   trials  N x 1 i32 device vector of population sizes, or NULL -> trials_scalar */
int smc_binomial_logit_glm(const smc_matrix* n, int n_scalar,
                           const smc_matrix* trials, int trials_scalar,
                           const smc_matrix* x, const smc_matrix* alpha_vec,
                           double alpha, const double* beta, unsigned flags,
                           double* logp, double* d_alpha,
                           smc_matrix* d_alpha_vec, double* d_beta,
                           smc_matrix* d_x);

/* prim/prob/poisson_log_glm_lpmf.hpp L51-163 */
int smc_poisson_log_glm(const smc_matrix* y, int y_scalar, const smc_matrix* x,
                        const smc_matrix* alpha_vec, double alpha,
                        const double* beta, unsigned flags, double* logp,
                        double* d_alpha, smc_matrix* d_alpha_vec,
                        double* d_beta, smc_matrix* d_x);

/* prim/prob/normal_id_glm_lpdf.hpp L54-216
 *   sigma_vec N x 1 device or NULL -> scalar sigma;  d_sigma host (scalar sigma)
 *   or d_sigma_vec device (vector sigma);  d_y host (scalar y) / d_y_vec device */
int smc_normal_id_glm(const smc_matrix* y, double y_scalar, const smc_matrix* x,
                      const smc_matrix* alpha_vec, double alpha,
                      const double* beta, const smc_matrix* sigma_vec,
                      double sigma, unsigned flags, double* logp,
                      double* d_alpha, smc_matrix* d_alpha_vec, double* d_beta,
                      double* d_sigma, smc_matrix* d_sigma_vec, double* d_y,
                      smc_matrix* d_y_vec, smc_matrix* d_x);

/* prim/prob/neg_binomial_2_log_glm_lpmf.hpp L64-248 */
int smc_neg_binomial_2_log_glm(const smc_matrix* y, int y_scalar,
                               const smc_matrix* x, const smc_matrix* alpha_vec,
                               double alpha, const double* beta,
                               const smc_matrix* phi_vec, double phi,
                               unsigned flags, double* logp, double* d_alpha,
                               smc_matrix* d_alpha_vec, double* d_beta,
                               double* d_phi, smc_matrix* d_phi_vec,
                               smc_matrix* d_x);

/* prim/prob/ordered_logistic_glm_lpmf.hpp L46-210; cuts host, ncuts = C-1 */
int smc_ordered_logistic_glm(const smc_matrix* y, int y_scalar,
                             const smc_matrix* x, const double* beta,
                             const double* cuts, int64_t ncuts, unsigned flags,
                             double* logp, double* d_beta, double* d_cuts,
                             smc_matrix* d_x);

/* prim/prob/categorical_logit_glm_lpmf.hpp L43-195
 *   alpha host (C), beta host column-major K x C, d_alpha host (C),
 *   d_beta host column-major K x C */
int smc_categorical_logit_glm(const smc_matrix* y, int y_scalar,
                              const smc_matrix* x, const double* alpha,
                              const double* beta, int64_t n_classes,
                              unsigned flags, double* logp, double* d_alpha,
                              double* d_beta, smc_matrix* d_x);

/* ---- device-resident results (pipelining / multi-GPU) -------------------- */
/* Same evaluation, but asynchronous on the thread's stream: parameters are read
 * from DEVICE memory (`params_dev`: beta[K], then cuts[ncuts] for ordered) and
 * the packed result (layout above, SMC_OUT_HEADER + K (+ ncuts) doubles) is
 * left in `out_dev` for an NCCL all-reduce; no host synchronisation and no
 * value checks (the caller inspects out[SMC_OUT_NONFINITE]).  `family`:
 * 0 normal_id, 1 bernoulli_logit, 2 poisson_log, 3 neg_binomial_2_log,
 * 4 ordered_logistic, 5 binomial_logit (aux_vec = the i32 trials vector, aux =
 * the scalar number of trials).  Scalars alpha/aux are passed by value; out[0]
 * already contains every term of the reference's logp for this rank's rows. */
int smc_glm_eval_device(int family, const smc_matrix* y, double y_scalar,
                        const smc_matrix* x, const smc_matrix* alpha_vec,
                        double alpha, const smc_matrix* aux_vec, double aux,
                        const double* params_dev, int64_t ncuts, unsigned flags,
                        double* out_dev, smc_matrix* d_alpha_vec,
                        smc_matrix* d_aux_vec, smc_matrix* d_y_vec,
                        smc_matrix* d_x);

/* The categorical GLM in the same asynchronous form: `params_dev` holds beta
 * (column-major K x C) followed by alpha (C) on the device; `out_dev` receives
 * [logp, #rows with a non-finite term, d_alpha[C], d_beta[K x C]] (2 + C + K*C
 * doubles) for the all-reduce.  No y-range / value checks (the caller checks y
 * once at upload and inspects out[1]). */
int smc_categorical_logit_glm_device(const smc_matrix* y, int y_scalar,
                                     const smc_matrix* x, const double* params_dev,
                                     int64_t n_classes, unsigned flags,
                                     double* out_dev, smc_matrix* d_x);

/* ---- the step either side of the GLMs ------------------ */
/* Models that add terms to the linear predictor before the likelihood form
 * theta on the device and call the un-fused density on it.  These replace what
 * the OpenCL backend builds from its kernel generator: the matrix-vector
 * product of opencl/prim/multiply.hpp + opencl/rev/multiply.hpp, and
 * opencl/prim/{bernoulli_logit,poisson_log,neg_binomial_2_log,ordered_logistic}_lpmf.hpp. */

/* theta_out[i] = sum_k x[i,k] beta[k] + alpha_i  (alpha_vec N x 1 or NULL ->
 * the scalar alpha).  One sweep over x through the fused TMA kernel. */
int smc_linear_predictor(const smc_matrix* x, const double* beta,
                         const smc_matrix* alpha_vec, double alpha,
                         smc_matrix* theta_out);
/* Reverse sweep of the product: xt_v[k] = sum_i x[i,k] v[i] (K host doubles, may
 * be NULL) and *sum_v = sum_i v[i] (may be NULL).  One sweep over x. */
int smc_linear_predictor_adjoint(const smc_matrix* x, const smc_matrix* v,
                                 double* xt_v, double* sum_v);
/* The matrix form, for a K x C weight matrix (the product in front of an un-fused
 * categorical_logit_lpmf; opencl/prim/multiply.hpp): lin_out[i,c] = sum_k x[i,k]
 * beta[k,c] + alpha[c].  beta host column-major K x C, alpha host C doubles or NULL
 * (no intercept), lin_out an N x C f64 device matrix.  One sweep over x per block of
 * 64 classes through the categorical GLM's FP64 tensor-core (DMMA) kernel.
 * Asynchronous on the thread's stream. */
int smc_linear_predictor_matrix(const smc_matrix* x, const double* beta,
                                int64_t n_classes, const double* alpha,
                                smc_matrix* lin_out);
/* Reverse sweep of the matrix form (opencl/rev/multiply.hpp L25-60): xt_adj[k,c] =
 * sum_i x[i,k] adj[i,c] (host column-major K x C, may be NULL) and colsum[c] =
 * sum_i adj[i,c] (host C doubles, may be NULL) for an N x C f64 device matrix adj.
 * Fixed-order sums: bit-reproducible. */
int smc_linear_predictor_matrix_adjoint(const smc_matrix* x, const smc_matrix* adj,
                                        double* xt_adj, double* colsum);
/* *sum = sum_i v[i] of an f64 device vector (fixed-order, deterministic). */
int smc_vector_sum(const smc_matrix* v, double* sum);

/* Hierarchical intercepts: out[i] = z[idx[i]] for a small host
 * vector z (G doubles) and a resident i32 index vector (0-based, like
 * opencl/kernel_generator/indexing.hpp), and the reverse sweep
 * adj_z[g] += sum_{i : idx[i] == g} res_adj[i] (opencl/indexing_rev.hpp L24-60;
 * deterministic here, fixed-order instead of atomics).  adj_z is ACCUMULATED
 * into.  An index outside [0, G) is SMC_ERR_DOMAIN. */
int smc_indexing(const double* z, int64_t G, const smc_matrix* idx, smc_matrix* out);
int smc_indexing_rev(const smc_matrix* idx, const smc_matrix* res_adj, int64_t G,
                     double* adj_z);

/* Un-fused densities on a device N-vector parameter.  `n`/`y`: N x 1 i32 device
 * vector or NULL -> the broadcast scalar.  SMC_VAR_ALPHA in `flags` marks the
 * vector parameter (theta / alpha / eta / lambda) as an autodiff variable: its
 * partial is written to the N x 1 device vector `d_*`; SMC_VAR_AUX marks phi /
 * cuts.  Value and error semantics follow prim/prob/<name>.hpp. */
/* prim/prob/bernoulli_logit_lpmf.hpp L33-98 */
int smc_bernoulli_logit_lpmf(const smc_matrix* n, int n_scalar,
                             const smc_matrix* theta, unsigned flags, double* logp,
                             smc_matrix* d_theta);
/* prim/prob/poisson_log_lpmf.hpp L27-100 */
int smc_poisson_log_lpmf(const smc_matrix* n, int n_scalar, const smc_matrix* alpha,
                         unsigned flags, double* logp, smc_matrix* d_alpha);
/* prim/prob/neg_binomial_2_log_lpmf.hpp L24-134 (phi: N x 1 device vector or
 * NULL -> scalar; d_phi host for a scalar phi, d_phi_vec device for a vector) */
int smc_neg_binomial_2_log_lpmf(const smc_matrix* n, int n_scalar,
                                const smc_matrix* eta, const smc_matrix* phi_vec,
                                double phi, unsigned flags, double* logp,
                                smc_matrix* d_eta, double* d_phi,
                                smc_matrix* d_phi_vec);
/* prim/prob/normal_lpdf.hpp L41-104.  y and mu: N x 1 f64 device vectors or NULL ->
 * the broadcast scalars (at least one is a vector); sigma a host scalar.  Flags:
 * SMC_VAR_Y (d_y_vec device / d_y host for a scalar y), SMC_VAR_ALPHA marks mu
 * (d_mu_vec / d_mu), SMC_VAR_AUX marks sigma (d_sigma host). */
int smc_normal_lpdf(const smc_matrix* y, double y_scalar, const smc_matrix* mu,
                    double mu_scalar, double sigma, unsigned flags, double* logp,
                    smc_matrix* d_y_vec, double* d_y, smc_matrix* d_mu_vec,
                    double* d_mu, double* d_sigma);
/* prim/prob/ordered_logistic_lpmf.hpp L72-214 (one cut-point vector, host) */
int smc_ordered_logistic_lpmf(const smc_matrix* y, int y_scalar,
                              const smc_matrix* lambda, const double* cuts,
                              int64_t ncuts, unsigned flags, double* logp,
                              smc_matrix* d_lambda, double* d_cuts);
/* The same density with ONE CUT-POINT VECTOR PER OUTCOME (prim/prob/
 * ordered_logistic_lpmf.hpp L72-200 called with a std::vector of cut vectors;
 * opencl/prim/ordered_logistic_lpmf.hpp L68-160): `cuts` is a (C-1) x N f64 device
 * matrix, column i the cut points of outcome i (a (C-1) x 1 matrix is the one-vector
 * form above with the cut points resident).  SMC_VAR_AUX: the partial of the cut
 * points is written whole into `d_cuts`, a device matrix with the shape of `cuts`.
 * Columns that are not strictly increasing or whose first / last cut is not finite
 * are SMC_ERR_DOMAIN (check_ordered / check_finite, L112-122). */
int smc_ordered_logistic_lpmf_rows(const smc_matrix* y, int y_scalar,
                                   const smc_matrix* lambda, const smc_matrix* cuts,
                                   unsigned flags, double* logp, smc_matrix* d_lambda,
                                   smc_matrix* d_cuts);
/* prim/prob/categorical_logit_lpmf.hpp L16-32, one row of log odds per outcome:
 * `lin` is an N x C f64 device matrix and the result is
 * sum_i categorical_logit_lpmf(y_i, lin.row(i)^T) -- what a model that adds terms
 * to x * beta writes as a loop over the rows.  SMC_VAR_ALPHA marks lin as an
 * autodiff variable; its partial (one-hot(y_i) - softmax(lin_i)) is written to the
 * N x C device matrix `d_lin`.  y outside [1, C] and non-finite log odds are
 * SMC_ERR_DOMAIN (check_bounded L19, check_finite L22). */
int smc_categorical_logit_lpmf(const smc_matrix* y, int y_scalar,
                               const smc_matrix* lin, unsigned flags, double* logp,
                               smc_matrix* d_lin);

#ifdef __cplusplus
}
#endif
#endif /* STANMATH_CUDA_H */
