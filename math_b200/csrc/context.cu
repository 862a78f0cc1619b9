// Per-thread runtime state and the device matrix type (matrix_cl analogue,
// reference: stan/math/opencl/matrix_cl.hpp L46-55, opencl/copy.hpp).
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "smc_internal.h"

namespace smc {

// t_ctx is the calling thread's own context; t_cur is the one launches go to: t_ctx,
// or -- while a row-sharded call walks over the GPUs (sharded.cu) -- the context of
// the shard's device.  Error text always lands in the thread's own context.
static thread_local Context t_ctx;
static thread_local Context* t_cur = &t_ctx;

Context& ctx() { return *t_cur; }
Context& own_ctx() { return t_ctx; }
Context* swap_current_context(Context* c) {
  Context* prev = t_cur;
  t_cur = c ? c : &t_ctx;
  return prev == &t_ctx ? nullptr : prev;
}

Context::~Context() {
  // Runs at thread exit; the CUDA runtime may already be gone at process exit,
  // so errors are ignored.
  if (!inited) return;
  if (cudaSetDevice(device) != cudaSuccess) return;
  for (auto& kv : block_cache)
    for (void* b : kv.second) cudaFree(b);
  block_cache.clear();
  if (partials) cudaFree(partials);
  if (counter) cudaFree(counter);
  if (params_dev) cudaFree(params_dev);
  if (scratch) cudaFree(scratch);
  if (out_host) cudaFreeHost(out_host);
  for (cudaEvent_t e : timer_ev)
    if (e) cudaEventDestroy(e);
  if (own_stream) cudaStreamDestroy(own_stream);
}

std::mutex& cache_mutex() {
  static std::mutex m;
  return m;
}

static std::atomic<uint64_t> g_next_id{1};
uint64_t next_matrix_id() { return g_next_id.fetch_add(1, std::memory_order_relaxed); }

const Knobs& knobs() {
  static const Knobs k = [] {
    auto on = [](const char* name) {
      const char* e = getenv(name);
      return e && e[0] == '1';
    };
    Knobs v;
    v.force_generic = on("SMC_FORCE_GENERIC");
    v.dx_fused = on("SMC_DX_FUSED");
    v.cat_dx_fma = on("SMC_CAT_DX_FMA");
    v.cat_no_tma = on("SMC_CAT_NO_TMA");
    const char* ks = getenv("SMC_CAT_KS");
    v.cat_ks = ks ? atoi(ks) : 0;
    return v;
  }();
  return k;
}

int fail(int status, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_ctx.last_error = buf;
  return status;
}

int init_context(Context& c, int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(SMC_ERR_CUDA,
                "no CUDA device available (%s); libstanmath_cuda has no CPU "
                "fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n)
    return fail(SMC_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)",
                device, n);
  if (c.inited && c.device == device) {
    SMC_CUDA(cudaSetDevice(device));
    return SMC_OK;
  }
  if (c.inited) {
    // moving the context to another device: drop the old workspace
    c.~Context();
    new (&c) Context();
  }
  SMC_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SMC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(SMC_ERR_UNSUPPORTED,
                "device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  c.device = device;
  c.id = g_next_id.fetch_add(1, std::memory_order_relaxed);
  c.sm_count = prop.multiProcessorCount;
  SMC_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
  c.stream = c.own_stream;
  SMC_CUDA(cudaMalloc(&c.counter, 64));
  SMC_CUDA(cudaMemset(c.counter, 0, 64));
  c.inited = true;
  return SMC_OK;
}

static int bind_device(int device) { return init_context(t_ctx, device); }

int ensure_ctx() {
  if (t_cur->inited) {
    SMC_CUDA(cudaSetDevice(t_cur->device));
    return SMC_OK;
  }
  return bind_device(0);
}

static int grow(void** p, size_t* have, size_t want, bool pinned) {
  if (*have >= want) return SMC_OK;
  size_t n = want < 4096 ? 4096 : want;
  if (*p) {
    // make sure nothing in flight still uses the old buffer
    SMC_CUDA(cudaStreamSynchronize(t_cur->stream));
    if (pinned)
      SMC_CUDA(cudaFreeHost(*p));
    else
      SMC_CUDA(cudaFree(*p));
    *p = nullptr;
    *have = 0;
  }
  if (pinned)
    SMC_CUDA(cudaHostAlloc(p, n, cudaHostAllocMapped | cudaHostAllocPortable));
  else
    SMC_CUDA(cudaMalloc(p, n));
  *have = n;
  return SMC_OK;
}

int cache_alloc(void** p, size_t bytes) {
  Context& c = *t_cur;
  auto it = c.block_cache.find(bytes);
  if (it != c.block_cache.end() && !it->second.empty()) {
    *p = it->second.back();
    it->second.pop_back();
    c.cached_bytes -= bytes;
    return SMC_OK;
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation && c.cached_bytes > 0) {
    (void)cudaGetLastError();
    cache_trim();  // give the cached blocks back and retry once
    e = cudaMalloc(p, bytes);
  }
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(SMC_ERR_CUDA, "cudaMalloc(%zu bytes) failed: %s", bytes,
                cudaGetErrorString(e));
  }
  return SMC_OK;
}

void cache_free(void* p, size_t bytes) {
  // work queued on this thread's stream may still touch the block; the next
  // owner is this thread again, on the same stream, so no synchronisation
  Context& c = *t_cur;
  c.block_cache[bytes].push_back(p);
  c.cached_bytes += bytes;
}

void cache_trim() {
  Context& c = *t_cur;
  if (c.cached_bytes == 0) return;
  cudaStreamSynchronize(c.stream);
  for (auto& kv : c.block_cache)
    for (void* b : kv.second) cudaFree(b);
  c.block_cache.clear();
  c.cached_bytes = 0;
}

int ensure_partials(size_t bytes) {
  return grow(reinterpret_cast<void**>(&t_cur->partials), &t_cur->partials_bytes,
              bytes, false);
}
int ensure_params(size_t bytes) {
  return grow(reinterpret_cast<void**>(&t_cur->params_dev), &t_cur->params_bytes,
              bytes, false);
}
int ensure_out(size_t bytes) {
  return grow(reinterpret_cast<void**>(&t_cur->out_host), &t_cur->out_bytes, bytes,
              true);
}
int ensure_scratch(size_t bytes) {
  return grow(reinterpret_cast<void**>(&t_cur->scratch), &t_cur->scratch_bytes,
              bytes, false);
}

// ---------------------------------------------------------------- kernels
// Element-wise sweeps over a column-major matrix with a leading dimension.  The
// read-modify-write ones use a flat index: its 64-bit division hides under three
// memory accesses per element (the N x K reverse sweep x.adj += lp.adj * d_x of an
// autodiff design matrix: 4.9 ms = 6.3 TB/s at N=1e7, K=128; a 2-D grid measured
// 5 % slower).
__global__ void axpy_kernel(double* __restrict__ y, int64_t ldy,
                            const double* __restrict__ x, int64_t ldx,
                            int64_t rows, int64_t cols, double a) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / rows, r = i - c * rows;
    y[c * ldy + r] += a * x[c * ldx + r];
  }
}

__global__ void scale_copy_kernel(double* __restrict__ y, int64_t ldy,
                                  const double* __restrict__ x, int64_t ldx,
                                  int64_t rows, int64_t cols, double a) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / rows, r = i - c * rows;
    y[c * ldy + r] = a * x[c * ldx + r];
  }
}

__global__ void add_scalar_kernel(double* __restrict__ y, int64_t ldy, int64_t rows,
                                  int64_t cols, double a) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / rows, r = i - c * rows;
    y[c * ldy + r] += a;
  }
}

// The read-only scan is instruction-bound with a flat index (2.57 ms at N=1e7,
// K=128): blockIdx.y strides the columns, the threads of blockIdx.x the rows
// (1.42 ms = 7.2 TB/s).
__global__ void all_finite_kernel(const double* __restrict__ x, int64_t ld,
                                  int64_t rows, int64_t cols, int* bad) {
  int local = 0;
  for (int64_t c = blockIdx.y; c < cols; c += gridDim.y) {
    const double* xc = x + c * ld;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows;
         r += (int64_t)gridDim.x * blockDim.x)
      if (!isfinite(xc[r])) local = 1;
  }
  if (__any_sync(0xffffffffu, local) && (threadIdx.x & 31) == 0) *bad = 1;
}

// out[i, k] (+)= a * (beta[k] * d[i]).  RMW = false: the N x K partial beta (x) d of an
// autodiff design matrix -- or its whole reverse sweep x.adj = lp.adj * beta (x) d when
// the adjoint is known to be zero -- as a pure write stream (two rows per thread: one
// 16-byte store per column).  RMW = true: x.adj += lp.adj * beta (x) d as one
// read-modify-write of the adjoint; the product is never materialised
// (rev/functor/operands_and_partials.hpp L28-38 with partial = beta (x) d,
// prim/prob/neg_binomial_2_log_glm_lpmf.hpp L221-222).  SCALED = false skips the
// multiplication by a (a = 1 would be exact anyway; it keeps the plain partial's
// instruction stream as it was).
struct OuterArgs {
  double* out;
  int64_t ld;
  const double* d;
  int64_t N;
  int K;
  double a;
  const double* beta_dev;  // NULL: beta[] below
  double beta[kMaxParamDoubles];
};
template <bool RMW, bool SCALED>
__global__ void __launch_bounds__(256)
    outer_kernel(const __grid_constant__ OuterArgs a) {
  const int64_t npairs = (a.N + 1) / 2;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = 2 * p;
    const bool two = i + 1 < a.N;
    const double d0 = a.d[i], d1 = two ? a.d[i + 1] : 0.0;
    double* o = a.out + i;
#pragma unroll 8
    for (int k = 0; k < a.K; ++k) {
      const double b = a.beta_dev ? __ldg(a.beta_dev + k) : a.beta[k];
      double v0 = b * d0, v1 = b * d1;
      if (SCALED) {
        v0 *= a.a;
        v1 *= a.a;
      }
      double* ok = o + (int64_t)k * a.ld;
      if (two) {
        double2 cur = make_double2(0.0, 0.0);
        if (RMW) cur = *reinterpret_cast<const double2*>(ok);
        *reinterpret_cast<double2*>(ok) = make_double2(cur.x + v0, cur.y + v1);
      } else {
        ok[0] = (RMW ? ok[0] : 0.0) + v0;
      }
    }
  }
}

// Four rows per thread: one 32-BYTE store per column (st.global.v4.f64, sm_100).  The
// store order is the same; the wider access alone takes the pure store stream from 6.1 to
// 6.9 TB/s at N = 1e7, K = 128 (profiles/micro/outer_store.cu: a fill kernel reaches 7.4,
// longer per-CTA runs, column panels, st.cs and TMA bulk stores from shared-memory tiles
// all stay at 5.0 - 6.2).  Needs a 32-byte aligned base and ld % 4 == 0.
__device__ __forceinline__ void st_v4(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d)
               : "memory");
}
__device__ __forceinline__ void ld_v4(const double* p, double& a, double& b, double& c,
                                      double& d) {
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];"
               : "=d"(a), "=d"(b), "=d"(c), "=d"(d)
               : "l"(p)
               : "memory");
}
template <bool RMW, bool SCALED>
__global__ void __launch_bounds__(256)
    outer_quad_kernel(const __grid_constant__ OuterArgs a) {
  const int64_t nquads = (a.N + 3) / 4;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nquads;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = 4 * p;
    const int live = a.N - i >= 4 ? 4 : (int)(a.N - i);
    double dv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) dv[j] = j < live ? a.d[i + j] : 0.0;
    double* o = a.out + i;
#pragma unroll 4
    for (int k = 0; k < a.K; ++k) {
      double b = a.beta_dev ? __ldg(a.beta_dev + k) : a.beta[k];
      double v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = b * dv[j];
        if (SCALED) v[j] *= a.a;
      }
      double* ok = o + (int64_t)k * a.ld;
      if (live == 4) {
        if (RMW) {
          double c0, c1, c2, c3;
          ld_v4(ok, c0, c1, c2, c3);
          st_v4(ok, c0 + v[0], c1 + v[1], c2 + v[2], c3 + v[3]);
        } else {
          st_v4(ok, v[0], v[1], v[2], v[3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < live) ok[j] = (RMW ? ok[j] : 0.0) + v[j];
      }
    }
  }
}

// The same update for layouts the paired kernel cannot take (odd leading dimension,
// unaligned base: one-row shards, wrapped buffers): one element per thread.
template <bool RMW>
__global__ void outer_scalar_kernel(const __grid_constant__ OuterArgs a) {
  const int64_t total = a.N * a.K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / a.N, r = i - k * a.N;
    const double b = a.beta_dev ? __ldg(a.beta_dev + k) : a.beta[k];
    const double v = (b * a.d[r]) * a.a;
    double* o = a.out + k * a.ld + r;
    *o = (RMW ? *o : 0.0) + v;
  }
}

__host__ __device__ inline uint64_t synth_hash(uint64_t seed, uint64_t row,
                                               uint64_t col) {
  uint64_t z = seed + row * 0x9E3779B97F4A7C15ull + col * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void fill_synth_kernel(void* data, int64_t ld, int64_t rows,
                                  int64_t cols, int dtype, uint64_t seed,
                                  int64_t row0, int kind, double c, int lo,
                                  int hi) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t col = i / rows, r = i - col * rows;
    const uint64_t z = synth_hash(seed, (uint64_t)(row0 + r), (uint64_t)col);
    if (kind == 0) {
      const uint32_t u = (uint32_t)(z & 0xffff) + (uint32_t)((z >> 16) & 0xffff)
                         + (uint32_t)((z >> 32) & 0xffff) + (uint32_t)(z >> 48);
      static_cast<double*>(data)[col * ld + r] = ((double)u - 131070.0) * c;
    } else {
      const uint64_t span = (uint64_t)((int64_t)hi - (int64_t)lo + 1);
      const int v = (int)((int64_t)lo + (int64_t)(z % span));
      if (dtype == SMC_I32)
        static_cast<int*>(data)[col * ld + r] = v;
      else
        static_cast<double*>(data)[col * ld + r] = (double)v;
    }
  }
}

static inline int grid_for(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)t_cur->sm_count * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// Grid of the 2-D element-wise kernels: row chunks x columns, at most 16 CTAs per SM.
static inline dim3 grid2d_for(int64_t rows, int64_t cols, int threads) {
  const int64_t cap = (int64_t)t_cur->sm_count * 16;
  int64_t gx = (rows + threads - 1) / threads;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  int64_t gy = cap / gx;
  if (gy > cols) gy = cols;
  if (gy > 65535) gy = 65535;
  if (gy < 1) gy = 1;
  return dim3((unsigned)gx, (unsigned)gy);
}

// out (N x K, ld) (+)= a * d beta^T on this thread's stream; beta from the host or,
// when beta_dev is set, from device memory (the asynchronous multi-GPU path).
static int launch_rank1(double* out, int64_t ld, const double* d, int64_t N, int K,
                        const double* beta_host, const double* beta_dev, double scale,
                        bool rmw) {
  if (N == 0 || K == 0) return SMC_OK;
  if (K > kMaxParamDoubles)
    return fail(SMC_ERR_UNSUPPORTED, "outer: needs K <= %d", kMaxParamDoubles);
  const bool paired = !(reinterpret_cast<uintptr_t>(out) & 15) && !(K > 1 && (ld & 1));
  OuterArgs a;
  a.out = out;
  a.ld = ld;
  a.d = d;
  a.N = N;
  a.K = K;
  a.a = scale;
  a.beta_dev = beta_dev;
  if (!beta_dev) memcpy(a.beta, beta_host, sizeof(double) * K);
  const bool quad = !(reinterpret_cast<uintptr_t>(out) & 31) && !(K > 1 && (ld & 3));
  const int grid = grid_for(quad ? (N + 3) / 4 : paired ? (N + 1) / 2 : N * K, 256);
  if (quad && rmw)
    outer_quad_kernel<true, true><<<grid, 256, 0, t_cur->stream>>>(a);
  else if (quad && scale != 1.0)
    outer_quad_kernel<false, true><<<grid, 256, 0, t_cur->stream>>>(a);
  else if (quad)
    outer_quad_kernel<false, false><<<grid, 256, 0, t_cur->stream>>>(a);
  else if (!paired && rmw)
    outer_scalar_kernel<true><<<grid, 256, 0, t_cur->stream>>>(a);
  else if (!paired)
    outer_scalar_kernel<false><<<grid, 256, 0, t_cur->stream>>>(a);
  else if (rmw)
    outer_kernel<true, true><<<grid, 256, 0, t_cur->stream>>>(a);
  else if (scale != 1.0)
    outer_kernel<false, true><<<grid, 256, 0, t_cur->stream>>>(a);
  else
    outer_kernel<false, false><<<grid, 256, 0, t_cur->stream>>>(a);
  SMC_CUDA(cudaGetLastError());
  t_cur->launches += 1;
  return SMC_OK;
}

int launch_outer(double* out, int64_t ld, const double* d, int64_t N, int K,
                 const double* beta_host, const double* beta_dev) {
  return launch_rank1(out, ld, d, N, K, beta_host, beta_dev, 1.0, false);
}

// The deferred memset of a lazily zeroed matrix (smc_matrix_zero_lazy): runs before
// anything reads the matrix or writes only part of it.
int realize(const smc_matrix* mc) {
  smc_matrix* m = const_cast<smc_matrix*>(mc);
  if (!m || !m->zero_pending) return SMC_OK;
  if (int rc = ensure_ctx()) return rc;
  m->zero_pending = false;
  if (m->rows == 0 || m->cols == 0) return SMC_OK;
  const size_t es = m->dtype == SMC_F64 ? 8 : 4;
  SMC_CUDA(cudaMemset2DAsync(m->data, (size_t)m->ld * es, 0, (size_t)m->rows * es,
                             (size_t)m->cols, t_cur->stream));
  return SMC_OK;
}

static inline size_t elem_size(int dtype) { return dtype == SMC_F64 ? 8 : 4; }

}  // namespace smc

using namespace smc;

extern "C" {

int smc_device_count(int* count) {
  if (!count) return fail(SMC_ERR_INVALID_ARGUMENT, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(SMC_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *count = n;
  return SMC_OK;
}

int smc_set_device(int device) { return bind_device(device); }

int smc_get_device(int* device) {
  if (int rc = ensure_ctx()) return rc;
  *device = ctx().device;
  return SMC_OK;
}

int smc_trim_cache(void) {
  if (int rc = ensure_ctx()) return rc;
  cache_trim();
  return SMC_OK;
}

int smc_set_stream(void* s) {
  if (int rc = ensure_ctx()) return rc;
  // recycled blocks are ordered on the stream they were last used on
  SMC_CUDA(cudaStreamSynchronize(ctx().stream));
  ctx().stream = s ? static_cast<cudaStream_t>(s) : ctx().own_stream;
  return SMC_OK;
}

int smc_synchronize(void) {
  if (int rc = ensure_ctx()) return rc;
  SMC_CUDA(cudaStreamSynchronize(ctx().stream));
  return synchronize_shards();  // (no shards: nothing to wait for)
}

int smc_device_info(int* sm_count, int* cc_major, int* cc_minor,
                    size_t* free_bytes, size_t* total_bytes) {
  if (int rc = ensure_ctx()) return rc;
  cudaDeviceProp prop;
  SMC_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  size_t f = 0, t = 0;
  SMC_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return SMC_OK;
}

// Device-side timing of whatever the calling thread queues between the two calls:
// CUDA events on the launching stream(s), as bench.py's roofline figures need them.
int smc_timer_start(void) {
  if (int rc = ensure_ctx()) return rc;
  Context& c = ctx();
  for (cudaEvent_t& e : c.timer_ev)
    if (!e) SMC_CUDA(cudaEventCreate(&e));
  SMC_CUDA(cudaEventRecord(c.timer_ev[0], c.stream));
  return shards_timer_start();
}

int smc_timer_stop(double* ms) {
  if (!ms) return fail(SMC_ERR_INVALID_ARGUMENT, "ms is NULL");
  if (int rc = ensure_ctx()) return rc;
  Context& c = ctx();
  if (!c.timer_ev[0] || !c.timer_ev[1])
    return fail(SMC_ERR_INVALID_ARGUMENT, "smc_timer_stop without smc_timer_start");
  SMC_CUDA(cudaEventRecord(c.timer_ev[1], c.stream));
  double shard_ms = 0.0;
  bool any = false;
  if (int rc = shards_timer_stop(&shard_ms, &any)) return rc;
  SMC_CUDA(cudaSetDevice(c.device));
  SMC_CUDA(cudaEventSynchronize(c.timer_ev[1]));
  float own = 0.f;
  SMC_CUDA(cudaEventElapsedTime(&own, c.timer_ev[0], c.timer_ev[1]));
  *ms = any && shard_ms > own ? shard_ms : (double)own;
  return SMC_OK;
}

const char* smc_last_error(void) { return t_ctx.last_error.c_str(); }
int64_t smc_launch_count(void) { return ctx().launches; }
void smc_reset_launch_count(void) { ctx().launches = 0; }

// ------------------------------------------------------------------ matrix
int smc_matrix_create(int64_t rows, int64_t cols, int dtype, smc_matrix** out) {
  if (!out) return fail(SMC_ERR_INVALID_ARGUMENT, "out is NULL");
  if (rows < 0 || cols < 0 || (dtype != SMC_F64 && dtype != SMC_I32))
    return fail(SMC_ERR_INVALID_ARGUMENT, "bad matrix shape/dtype %lld x %lld",
                (long long)rows, (long long)cols);
  if (int rc = ensure_ctx()) return rc;
  smc_matrix* m = new smc_matrix();
  m->rows = rows;
  m->cols = cols;
  m->dtype = dtype;
  m->device = ctx().device;
  m->id = next_matrix_id();
  m->home_ctx = ctx().id;
  const int64_t align = 128 / (int64_t)elem_size(dtype);
  // vectors (N x 1 and 1 x N) are contiguous: the kernels index them as data[i]
  m->ld = (cols > 1 && rows > 1) ? (rows + align - 1) / align * align : rows;
  if (m->ld < 1) m->ld = 1;
  const size_t bytes = (size_t)m->ld * (size_t)(cols > 0 ? cols : 1) * elem_size(dtype);
  m->alloc_bytes = bytes < 256 ? 256 : (bytes + 255) & ~(size_t)255;
  if (int rc = cache_alloc(&m->data, m->alloc_bytes)) {
    delete m;
    return rc;
  }
  m->owned = true;
  *out = m;
  return SMC_OK;
}

int smc_matrix_wrap(void* device_ptr, int64_t rows, int64_t cols, int64_t ld,
                    int dtype, smc_matrix** out) {
  if (!out || (!device_ptr && rows * cols > 0) || rows < 0 || cols < 0
      || ld < rows || (dtype != SMC_F64 && dtype != SMC_I32))
    return fail(SMC_ERR_INVALID_ARGUMENT, "bad arguments to smc_matrix_wrap");
  if (int rc = ensure_ctx()) return rc;
  smc_matrix* m = new smc_matrix();
  m->data = device_ptr;
  m->rows = rows;
  m->cols = cols;
  m->ld = ld < 1 ? 1 : ld;
  m->dtype = dtype;
  m->owned = false;
  m->device = ctx().device;
  m->id = next_matrix_id();
  m->home_ctx = ctx().id;
  *out = m;
  return SMC_OK;
}

int smc_matrix_free(smc_matrix* m) {
  if (!m) return SMC_OK;
  if (is_sharded(m)) {
    // every piece goes back under its own shard's context (its recycling cache); if
    // the shard set is gone the pieces fall through to the foreign-context path
    const int rc = for_each_shard(m, [](int, smc_matrix* p, int64_t) {
      return smc_matrix_free(p);
    });
    if (rc)
      for (smc_matrix* p : m->shards) smc_matrix_free(p);
    delete m;
    return SMC_OK;
  }
  const bool has_blocks = m->grp_perm || m->grp_off || (m->owned && m->data);
  if (has_blocks) {
    if (int rc = ensure_ctx()) return rc;
    // Recycling is stream-ordered only inside the thread (Context) that created the
    // block and queued work on it.  Freed from anywhere else -- a finalizer thread,
    // a TBB worker dropping the last reference, a thread bound to another GPU --
    // the block goes straight back to the driver: cudaFree waits for the device.
    const bool home = m->home_ctx == ctx().id && m->device == ctx().device;
    auto give_back = [&](void* p, size_t bytes) {
      if (!p) return;
      if (home) {
        cache_free(p, bytes);
      } else {
        cudaSetDevice(m->device);
        cudaFree(p);
        cudaSetDevice(ctx().device);
      }
    };
    give_back(m->grp_perm, m->grp_perm_bytes);
    give_back(m->grp_off, m->grp_off_bytes);
    if (m->owned) give_back(m->data, m->alloc_bytes);
  }
  delete m;
  return SMC_OK;
}

int smc_matrix_invalidate(smc_matrix* m) {
  if (!m) return fail(SMC_ERR_INVALID_ARGUMENT, "NULL matrix");
  for (smc_matrix* p : m->shards) smc_matrix_invalidate(p);
  std::lock_guard<std::mutex> lock(cache_mutex());
  m->lgamma_valid = false;
  m->range_valid = false;
  m->binom_valid = false;
  m->version += 1;
  return SMC_OK;
}

int64_t smc_matrix_rows(const smc_matrix* m) { return m ? m->rows : 0; }
int64_t smc_matrix_cols(const smc_matrix* m) { return m ? m->cols : 0; }
int64_t smc_matrix_ld(const smc_matrix* m) { return m ? m->ld : 0; }
int smc_matrix_dtype(const smc_matrix* m) { return m ? m->dtype : -1; }
void* smc_matrix_data(const smc_matrix* m) {
  if (!m || is_sharded(m)) return nullptr;  // (a sharded matrix has no one buffer)
  realize(m);  // the raw pointer escapes: a deferred memset has to happen now
  return m->data;
}

static int copy_rows(smc_matrix* m, int64_t row0, int64_t nrows, void* host,
                     int64_t ld_host, bool to_device) {
  if (!m || (!host && nrows * m->cols > 0))
    return fail(SMC_ERR_INVALID_ARGUMENT, "NULL matrix or host pointer");
  if (row0 < 0 || nrows < 0 || row0 + nrows > m->rows || ld_host < nrows)
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "row block [%lld,%lld) outside matrix with %lld rows (ld_host "
                "%lld)",
                (long long)row0, (long long)(row0 + nrows), (long long)m->rows,
                (long long)ld_host);
  if (is_sharded(m)) {
    // scatter / gather: every shard moves its part of the row block (the copies of
    // the shards run back to back; each waits for its own stream)
    const size_t es = m->dtype == SMC_F64 ? 8 : 4;
    return for_each_shard(m, [&](int, smc_matrix* p, int64_t p0) {
      const int64_t lo = row0 > p0 ? row0 : p0;
      const int64_t hi = row0 + nrows < p0 + p->rows ? row0 + nrows : p0 + p->rows;
      if (hi <= lo) return (int)SMC_OK;
      return copy_rows(p, lo - p0, hi - lo, static_cast<char*>(host) + (size_t)(lo - row0) * es,
                       ld_host, to_device);
    });
  }
  if (int rc = ensure_ctx()) return rc;
  if (nrows == 0 || m->cols == 0) return SMC_OK;
  const size_t es = elem_size(m->dtype);
  char* dev = static_cast<char*>(m->data) + (size_t)row0 * es;
  if (to_device && row0 == 0 && nrows == m->rows)
    m->zero_pending = false;  // every element is overwritten
  else if (int rc = realize(m))
    return rc;
  if (to_device) {
    m->lgamma_valid = false;
    m->range_valid = false;
    m->version += 1;
    SMC_CUDA(cudaMemcpy2DAsync(dev, (size_t)m->ld * es, host,
                               (size_t)ld_host * es, (size_t)nrows * es,
                               (size_t)m->cols, cudaMemcpyHostToDevice,
                               ctx().stream));
  } else {
    SMC_CUDA(cudaMemcpy2DAsync(host, (size_t)ld_host * es, dev,
                               (size_t)m->ld * es, (size_t)nrows * es,
                               (size_t)m->cols, cudaMemcpyDeviceToHost,
                               ctx().stream));
  }
  SMC_CUDA(cudaStreamSynchronize(ctx().stream));
  return SMC_OK;
}

int smc_matrix_upload(smc_matrix* m, const void* host, int64_t ld_host) {
  return copy_rows(m, 0, m ? m->rows : 0, const_cast<void*>(host), ld_host, true);
}
int smc_matrix_upload_rows(smc_matrix* m, int64_t row0, int64_t nrows,
                           const void* host, int64_t ld_host) {
  // `host` addresses element (row0, 0) of the host matrix; the device copy
  // holds the block at rows [0, nrows) when the matrix is a shard, so row0
  // here indexes the DEVICE matrix.
  return copy_rows(m, row0, nrows, const_cast<void*>(host), ld_host, true);
}
int smc_matrix_download(const smc_matrix* m, void* host, int64_t ld_host) {
  return copy_rows(const_cast<smc_matrix*>(m), 0, m ? m->rows : 0, host, ld_host,
                   false);
}
int smc_matrix_download_rows(const smc_matrix* m, int64_t row0, int64_t nrows,
                             void* host, int64_t ld_host) {
  return copy_rows(const_cast<smc_matrix*>(m), row0, nrows, host, ld_host, false);
}

int smc_matrix_zero(smc_matrix* m) {
  if (!m) return fail(SMC_ERR_INVALID_ARGUMENT, "NULL matrix");
  if (is_sharded(m))
    return for_each_shard(m, [](int, smc_matrix* p, int64_t) { return smc_matrix_zero(p); });
  if (int rc = ensure_ctx()) return rc;
  if (m->rows == 0 || m->cols == 0) return SMC_OK;
  m->lgamma_valid = false;
  m->range_valid = false;
  m->version += 1;
  m->zero_pending = false;
  const size_t es = elem_size(m->dtype);
  SMC_CUDA(cudaMemset2DAsync(m->data, (size_t)m->ld * es, 0, (size_t)m->rows * es,
                             (size_t)m->cols, ctx().stream));
  return SMC_OK;
}

int smc_matrix_zero_lazy(smc_matrix* m) {
  if (!m) return fail(SMC_ERR_INVALID_ARGUMENT, "NULL matrix");
  for (smc_matrix* p : m->shards) smc_matrix_zero_lazy(p);
  m->lgamma_valid = false;
  m->range_valid = false;
  m->version += 1;
  m->zero_pending = true;
  return SMC_OK;
}

int smc_matrix_copy(smc_matrix* dst, const smc_matrix* src) {
  if (!dst || !src || dst->rows != src->rows || dst->cols != src->cols
      || dst->dtype != src->dtype)
    return fail(SMC_ERR_INVALID_ARGUMENT, "copy: shape/dtype mismatch");
  if (is_sharded(dst) || is_sharded(src)) {
    if (!is_sharded(dst) || !is_sharded(src) || !same_partition(dst, src))
      return fail(SMC_ERR_INVALID_ARGUMENT, "copy: the operands are not sharded alike");
    return for_each_shard(dst, [&](int g, smc_matrix* p, int64_t) {
      return smc_matrix_copy(p, src->shards[g]);
    });
  }
  if (int rc = ensure_ctx()) return rc;
  if (dst->rows == 0 || dst->cols == 0) return SMC_OK;
  if (src->zero_pending) return dst == src ? SMC_OK : smc_matrix_zero(dst);
  dst->zero_pending = false;
  dst->lgamma_valid = false;
  dst->range_valid = false;
  dst->version += 1;
  const size_t es = elem_size(dst->dtype);
  SMC_CUDA(cudaMemcpy2DAsync(dst->data, (size_t)dst->ld * es, src->data,
                             (size_t)src->ld * es, (size_t)src->rows * es,
                             (size_t)src->cols, cudaMemcpyDeviceToDevice,
                             ctx().stream));
  return SMC_OK;
}

int smc_matrix_axpy(smc_matrix* y, double a, const smc_matrix* x) {
  if (!y || !x || y->rows != x->rows || y->cols != x->cols
      || y->dtype != SMC_F64 || x->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "axpy: shape/dtype mismatch");
  if (is_sharded(y) || is_sharded(x)) {
    if (!is_sharded(y) || !is_sharded(x) || !same_partition(y, x))
      return fail(SMC_ERR_INVALID_ARGUMENT, "axpy: the operands are not sharded alike");
    return for_each_shard(y, [&](int g, smc_matrix* p, int64_t) {
      return smc_matrix_axpy(p, a, x->shards[g]);
    });
  }
  if (int rc = ensure_ctx()) return rc;
  const int64_t total = y->rows * y->cols;
  if (total == 0) return SMC_OK;
  if (x->zero_pending) return SMC_OK;  // y += a * 0
  y->version++;
  if (y->zero_pending) {
    // y = a * x: the adjoint's deferred memset and its read both drop out
    y->zero_pending = false;
    scale_copy_kernel<<<grid_for(total, 256), 256, 0, ctx().stream>>>(
        static_cast<double*>(y->data), y->ld, static_cast<const double*>(x->data),
        x->ld, y->rows, y->cols, a);
  } else {
    axpy_kernel<<<grid_for(total, 256), 256, 0, ctx().stream>>>(
        static_cast<double*>(y->data), y->ld, static_cast<const double*>(x->data),
        x->ld, y->rows, y->cols, a);
  }
  SMC_CUDA(cudaGetLastError());
  ctx().launches += 1;
  return SMC_OK;
}

int smc_matrix_rank1_update(smc_matrix* y, double a, const smc_matrix* d,
                            const double* beta) {
  if (!y || !d || !beta || y->dtype != SMC_F64 || d->dtype != SMC_F64
      || d->rows * d->cols != y->rows || !vec_contiguous(d))
    return fail(SMC_ERR_INVALID_ARGUMENT, "rank1_update: shape/dtype mismatch");
  if (is_sharded(y) || is_sharded(d)) {
    if (!is_sharded(y) || !is_sharded(d) || !same_partition(y, d))
      return fail(SMC_ERR_INVALID_ARGUMENT,
                  "rank1_update: the operands are not sharded alike");
    return for_each_shard(y, [&](int g, smc_matrix* p, int64_t) {
      return smc_matrix_rank1_update(p, a, d->shards[g], beta);
    });
  }
  if (int rc = ensure_ctx()) return rc;
  if (y->rows == 0 || y->cols == 0) return SMC_OK;
  if (int rc = realize(d)) return rc;
  y->version++;
  const bool rmw = !y->zero_pending;
  y->zero_pending = false;
  // wide matrices: column blocks of at most kMaxParamDoubles (beta travels by value)
  for (int64_t c0 = 0; c0 < y->cols; c0 += kMaxParamDoubles) {
    const int kc = (int)(y->cols - c0 < kMaxParamDoubles ? y->cols - c0 : kMaxParamDoubles);
    if (int rc = launch_rank1(static_cast<double*>(y->data) + c0 * y->ld, y->ld,
                              static_cast<const double*>(d->data), y->rows, kc, beta + c0,
                              nullptr, a, rmw))
      return rc;
  }
  return SMC_OK;
}

int smc_matrix_outer(smc_matrix* out, const smc_matrix* d, const double* beta) {
  if (!out || !d || !beta || out->dtype != SMC_F64 || d->dtype != SMC_F64
      || d->rows * d->cols != out->rows)
    return fail(SMC_ERR_INVALID_ARGUMENT, "outer: shape/dtype mismatch");
  if (is_sharded(out) || is_sharded(d)) {
    if (!is_sharded(out) || !is_sharded(d) || !same_partition(out, d))
      return fail(SMC_ERR_INVALID_ARGUMENT, "outer: the operands are not sharded alike");
    return for_each_shard(out, [&](int g, smc_matrix* p, int64_t) {
      return smc_matrix_outer(p, d->shards[g], beta);
    });
  }
  if (int rc = ensure_ctx()) return rc;
  if (int rc = realize(d)) return rc;
  out->version++;
  out->zero_pending = false;
  return launch_outer(static_cast<double*>(out->data), out->ld,
                      static_cast<const double*>(d->data), out->rows, (int)out->cols, beta,
                      nullptr);
}

int smc_matrix_add_scalar(smc_matrix* y, double a) {
  if (!y || y->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "add_scalar: need an f64 matrix");
  if (is_sharded(y))
    return for_each_shard(y, [&](int, smc_matrix* p, int64_t) {
      return smc_matrix_add_scalar(p, a);
    });
  if (int rc = ensure_ctx()) return rc;
  const int64_t total = y->rows * y->cols;
  if (total == 0) return SMC_OK;
  if (int rc = realize(y)) return rc;
  y->version++;
  add_scalar_kernel<<<grid_for(total, 256), 256, 0, ctx().stream>>>(
      static_cast<double*>(y->data), y->ld, y->rows, y->cols, a);
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

int smc_matrix_all_finite(const smc_matrix* m, int* all_finite) {
  if (!m || !all_finite || m->dtype != SMC_F64)
    return fail(SMC_ERR_INVALID_ARGUMENT, "all_finite: need an f64 matrix");
  if (is_sharded(m)) {
    *all_finite = 1;
    return for_each_shard(m, [&](int, smc_matrix* p, int64_t) {
      int ok = 1;
      const int rc = smc_matrix_all_finite(p, &ok);
      if (!ok) *all_finite = 0;
      return rc;
    });
  }
  if (int rc = ensure_ctx()) return rc;
  *all_finite = 1;
  const int64_t total = m->rows * m->cols;
  if (total == 0 || m->zero_pending) return SMC_OK;
  if (int rc = ensure_out(4096)) return rc;
  int* flag = reinterpret_cast<int*>(ctx().out_host);
  *flag = 0;
  all_finite_kernel<<<grid2d_for(m->rows, m->cols, 256), 256, 0, ctx().stream>>>(
      static_cast<const double*>(m->data), m->ld, m->rows, m->cols, flag);
  SMC_CUDA(cudaGetLastError());
  SMC_CUDA(cudaStreamSynchronize(ctx().stream));
  *all_finite = (*flag == 0);
  return SMC_OK;
}

int smc_matrix_int_range(const smc_matrix* m, int* min_out, int* max_out) {
  if (!m || m->dtype != SMC_I32 || !min_out || !max_out)
    return fail(SMC_ERR_INVALID_ARGUMENT, "int_range: need an i32 matrix");
  return y_range(m, min_out, max_out);
}

int smc_matrix_fill_synthetic(smc_matrix* m, uint64_t seed, int64_t row0,
                              int kind, double scale, int lo, int hi) {
  if (!m || (kind != 0 && kind != 1) || (kind == 0 && m->dtype != SMC_F64)
      || (kind == 1 && hi < lo))
    return fail(SMC_ERR_INVALID_ARGUMENT, "fill_synthetic: bad arguments");
  if (is_sharded(m))  // the same values as one GPU would hold: keyed by the global row
    return for_each_shard(m, [&](int, smc_matrix* p, int64_t p0) {
      return smc_matrix_fill_synthetic(p, seed, row0 + p0, kind, scale, lo, hi);
    });
  if (int rc = ensure_ctx()) return rc;
  const int64_t total = m->rows * m->cols;
  if (total == 0) return SMC_OK;
  m->lgamma_valid = false;
  m->range_valid = false;
  m->version += 1;
  m->zero_pending = false;
  const double c = scale / sqrt(4294967295.0 / 3.0);
  fill_synth_kernel<<<grid_for(total, 256), 256, 0, ctx().stream>>>(
      m->data, m->ld, m->rows, m->cols, m->dtype, seed, row0, kind, c, lo, hi);
  SMC_CUDA(cudaGetLastError());
  return SMC_OK;
}

}  // extern "C"
