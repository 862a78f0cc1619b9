// Row-sharded device matrices: the GPUs of one box behind the SAME entry points.
//
// A sharded smc_matrix is a handle whose rows are partitioned contiguously over G
// GPUs (shard g holds rows [g N / G, (g+1) N / G), uploaded once).  Every GLM entry
// point of the C ABI -- and therefore every C++ overload in include/stan/math/cuda/
// -- takes such a handle wherever it takes a plain one: one evaluation then
//   1. launches the fused kernel on every GPU, each on its own stream, with the small
//      parameters travelling as kernel arguments (that IS the broadcast: G launches
//      carry beta to G devices, no copy, no collective),
//   2. all-reduces the packed K + O(1) result in place over NCCL (NVLink / NVSwitch)
//      on those same streams,
//   3. reads the reduced vector back once from shard 0 into pinned host memory.
// N-vector partials and the N x K adjoint of an autodiff x stay sharded: no exchange.
//
// This is the reference's "scatter static data once, broadcast parameters per call,
// reduce results" pattern (stan/math/prim/functor/mpi_parallel_call.hpp L332-392,
// L408-449) with one host process and one communicator rank per device, and it is
// what the OpenCL backend's single-device limit (opencl/opencl_context.hpp L75-76)
// keeps it from doing.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): programs that never shard do
// not need it.  SMC_SHARD_REDUCE=host replaces step 2 + 3 by G read-backs into pinned
// host memory and a sum there in shard order; that is also the
// fallback when NCCL is missing or when several shards share one GPU (single-GPU
// tests of the sharded logic).
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <chrono>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

#include "smc_internal.h"

namespace smc {

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t,
                            ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    // One NCCL per process: a second copy under the same soname would shadow the one
    // another component (a framework that bundles its own, newer NCCL) links against.
    // So: the copy that is already loaded, else the one SMC_NCCL_LIB names (the
    // Python binding points it at the bundled copy), else the system's.
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
    if (!lib) {
      const char* named = getenv("SMC_NCCL_LIB");
      if (named && named[0]) lib = dlopen(named, RTLD_NOW | RTLD_LOCAL);
    }
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      if (lib) break;
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    }
    if (!lib) return false;
#define SMC_NCCL_SYM(field, sym) \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, sym)); \
  if (!field) return false;
    SMC_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    SMC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    SMC_NCCL_SYM(AllReduce, "ncclAllReduce")
    SMC_NCCL_SYM(GroupStart, "ncclGroupStart")
    SMC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    SMC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SMC_NCCL_SYM
    return true;
  }
};

struct Shard {
  int device = 0;
  Context ctx;                  // stream + workspace of this shard
  ncclComm_t comm = nullptr;    // this shard's rank of the communicator
  double* out_dev = nullptr;    // packed result of the shard (all-reduced in place)
  size_t out_bytes = 0;
  // "direct" reduction: where this shard's kernel stores its completion flag (mapped
  // host memory, behind the shard's slot of the packed results) and the value to store
  unsigned long long* done_flag = nullptr;
  unsigned long long done_val = 0;
};

struct ShardComm {
  std::mutex mu;  // one sharded call at a time (chains sharing the GPUs take turns)
  std::vector<Shard*> shards;
  NcclApi nccl;
  bool use_nccl = false;
  bool direct = false;      // small results: zero-copy slots + completion flags
  std::string reduce_mode;  // "nccl" | "host" | "direct"
  unsigned long long seq = 0;
  double* host_out = nullptr;  // pinned read-back buffer of the reduced result
  size_t host_bytes = 0;
  std::vector<double> host_sum;
};

ShardComm& comm() {
  static ShardComm c;
  return c;
}

// Makes a shard's context (device, stream, workspace) the calling thread's current
// one for the lifetime of the scope.
struct ShardScope {
  Context* prev;
  explicit ShardScope(Shard* s) {
    prev = swap_current_context(&s->ctx);
    cudaSetDevice(s->device);
  }
  ~ShardScope() {
    swap_current_context(prev);
    Context& cur = ctx();
    if (cur.inited) cudaSetDevice(cur.device);
  }
};

#define SMC_NCCL(expr)                                                          \
  do {                                                                          \
    ncclResult_t r__ = (expr);                                                  \
    if (r__ != ncclSuccess)                                                     \
      return fail(SMC_ERR_CUDA, "%s failed: %s", #expr,                         \
                  comm().nccl.GetErrorString(r__));                             \
  } while (0)

void shutdown_locked(ShardComm& sc) {
  for (Shard* s : sc.shards) {
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->ctx.stream);
    if (s->comm && sc.nccl.CommDestroy) sc.nccl.CommDestroy(s->comm);
    if (s->out_dev) cudaFree(s->out_dev);
    delete s;  // ~Context releases the stream and the workspace
  }
  sc.shards.clear();
  if (sc.host_out) cudaFreeHost(sc.host_out);
  sc.host_out = nullptr;
  sc.host_bytes = 0;
  sc.use_nccl = false;
  Context& cur = ctx();
  if (cur.inited) cudaSetDevice(cur.device);
}

}  // namespace

int shard_count_of(const smc_matrix* m) { return m ? (int)m->shards.size() : 0; }

bool same_partition(const smc_matrix* a, const smc_matrix* b) {
  if (a->shards.size() != b->shards.size() || a->rows != b->rows) return false;
  for (size_t g = 0; g < a->shards.size(); ++g)
    if (a->shards[g]->rows != b->shards[g]->rows
        || a->shards[g]->device != b->shards[g]->device)
      return false;
  return true;
}

int for_each_shard(const smc_matrix* m,
                   const std::function<int(int, smc_matrix*, int64_t)>& fn) {
  ShardComm& sc = comm();
  if (m->shards.size() != sc.shards.size())
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "sharded matrix with %zu shards, but the shard set has %zu (was "
                "smc_shard_shutdown / smc_shard_init called since it was created?)",
                m->shards.size(), sc.shards.size());
  for (size_t g = 0; g < m->shards.size(); ++g) {
    ShardScope scope(sc.shards[g]);
    if (int rc = fn((int)g, m->shards[g], m->shard_row0[g])) return rc;
  }
  return SMC_OK;
}

int synchronize_shards() {
  ShardComm& sc = comm();
  for (Shard* s : sc.shards) {
    SMC_CUDA(cudaSetDevice(s->device));
    SMC_CUDA(cudaStreamSynchronize(s->ctx.stream));
  }
  Context& cur = ctx();
  if (cur.inited) SMC_CUDA(cudaSetDevice(cur.device));
  return SMC_OK;
}

int shards_timer_start() {
  ShardComm& sc = comm();
  for (Shard* s : sc.shards) {
    SMC_CUDA(cudaSetDevice(s->device));
    for (cudaEvent_t& e : s->ctx.timer_ev)
      if (!e) SMC_CUDA(cudaEventCreate(&e));
    SMC_CUDA(cudaEventRecord(s->ctx.timer_ev[0], s->ctx.stream));
  }
  Context& cur = ctx();
  if (cur.inited) SMC_CUDA(cudaSetDevice(cur.device));
  return SMC_OK;
}

int shards_timer_stop(double* ms_max, bool* any) {
  ShardComm& sc = comm();
  *ms_max = 0.0;
  *any = false;
  for (Shard* s : sc.shards) {
    if (!s->ctx.timer_ev[0] || !s->ctx.timer_ev[1]) continue;
    SMC_CUDA(cudaSetDevice(s->device));
    SMC_CUDA(cudaEventRecord(s->ctx.timer_ev[1], s->ctx.stream));
  }
  for (Shard* s : sc.shards) {
    if (!s->ctx.timer_ev[0] || !s->ctx.timer_ev[1]) continue;
    SMC_CUDA(cudaSetDevice(s->device));
    SMC_CUDA(cudaEventSynchronize(s->ctx.timer_ev[1]));
    float ms = 0.f;
    SMC_CUDA(cudaEventElapsedTime(&ms, s->ctx.timer_ev[0], s->ctx.timer_ev[1]));
    if (ms > *ms_max) *ms_max = ms;
    *any = true;
  }
  Context& cur = ctx();
  if (cur.inited) SMC_CUDA(cudaSetDevice(cur.device));
  return SMC_OK;
}

// The common skeleton of a sharded evaluation: launch(g, shard, out_g) queues shard
// g's work on its stream with the packed result (n_out doubles) going to out_g; the
// G results are then summed -- NCCL all-reduce in place on those streams and one
// read-back from shard 0, or (host mode) G results in mapped host memory added in
// shard order -- and *out points at the reduced vector in host memory.
// Direct mode (the default for results of at most kDirectMaxOut doubles, i.e. the
// memory-bound families; SMC_SHARD_REDUCE=nccl / host force the collective / the copies):
// the exchange is K + O(1) doubles per GPU, so instead of a collective every shard's
// kernel stores its packed result STRAIGHT into its slot of one
// pinned, mapped host buffer and its completion flag behind it (the mechanism of the
// one-GPU synchronous call, capi.cu run_sync); the host polls the G flags and adds the
// G slots in shard order -- bit-reproducible, no all-reduce, no copy, no stream
// synchronise on the evaluation path.
constexpr int kDirectMaxOut = 4096;

static int run_over_shards(const smc_matrix* x, int n_out,
                           const std::function<int(int, Shard*, double*)>& launch,
                           const double** out, bool allow_direct = false) {
  ShardComm& sc = comm();
  const int G = (int)x->shards.size();
  if ((int)sc.shards.size() != G)
    return fail(SMC_ERR_INVALID_ARGUMENT,
                "x has %d shards, but the shard set has %zu", G, sc.shards.size());
  std::lock_guard<std::mutex> lock(sc.mu);
  const size_t bytes = sizeof(double) * (size_t)n_out;
  int64_t launches = 0;
  if (allow_direct && sc.direct && n_out <= kDirectMaxOut) {
    const size_t stride = (size_t)n_out + 1;  // packed result + completion flag
    const size_t need = sizeof(double) * stride * (size_t)(G + 1);
    if (sc.host_bytes < need) {
      if (int rc = synchronize_shards()) return rc;
      if (sc.host_out) SMC_CUDA(cudaFreeHost(sc.host_out));
      sc.host_out = nullptr;
      const size_t want = need < 4096 ? 4096 : need;
      SMC_CUDA(cudaHostAlloc(&sc.host_out, want, cudaHostAllocPortable | cudaHostAllocMapped));
      sc.host_bytes = want;
    }
    const unsigned long long seq = ++sc.seq;
    for (int g = 0; g < G; ++g) {
      Shard* s = sc.shards[g];
      ShardScope scope(s);
      double* slot = sc.host_out + (size_t)(g + 1) * stride;
      volatile unsigned long long* flag
          = reinterpret_cast<volatile unsigned long long*>(slot + n_out);
      *flag = 0;
      s->done_flag = const_cast<unsigned long long*>(flag);
      s->done_val = seq;
      s->ctx.flag_armed = false;
      const int64_t before = s->ctx.launches;
      if (x->shards[g]->rows == 0) {  // more shards than rows: contributes zeros
        memset(slot, 0, bytes);
        *flag = seq;
        s->ctx.flag_armed = true;
      } else if (int rc = launch(g, s, slot)) {
        s->done_flag = nullptr;
        return rc;
      }
      s->done_flag = nullptr;
      launches += s->ctx.launches - before;
    }
    // shard 0 finishes first more often than not (it was launched first): poll in order
    // (the driver's own stream synchronise spins on the host too); a kernel that runs for
    // longer than the spin budget is waited for on its stream
    const auto t0 = std::chrono::steady_clock::now();
    for (int g = 0; g < G; ++g) {
      Shard* s = sc.shards[g];
      volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(
          sc.host_out + (size_t)(g + 1) * stride + n_out);
      bool done = false;
      if (s->ctx.flag_armed) {
        for (int spin = 0; !done; ++spin) {
          done = *flag == seq;
          if (!done && (spin & 63) == 63
              && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(20))
            break;
        }
      }
      if (!done) SMC_CUDA(cudaStreamSynchronize(s->ctx.stream));
    }
    for (int j = 0; j < n_out; ++j) {
      double v = 0.0;
      for (int g = 0; g < G; ++g) v += sc.host_out[(size_t)(g + 1) * stride + j];
      sc.host_out[j] = v;
    }
    own_ctx().launches += launches;
    *out = sc.host_out;
    return SMC_OK;
  }
  const bool host_reduce = !sc.use_nccl;
  for (int g = 0; g < G; ++g) {
    Shard* s = sc.shards[g];
    ShardScope scope(s);
    if (s->out_bytes < bytes) {
      SMC_CUDA(cudaStreamSynchronize(s->ctx.stream));
      if (s->out_dev) SMC_CUDA(cudaFree(s->out_dev));
      s->out_dev = nullptr;
      const size_t want = bytes < 4096 ? 4096 : bytes;
      SMC_CUDA(cudaMalloc(&s->out_dev, want));
      s->out_bytes = want;
    }
    double* out_g = s->out_dev;
    const int64_t before = s->ctx.launches;
    if (x->shards[g]->rows == 0) {  // more shards than rows: contributes zeros
      SMC_CUDA(cudaMemsetAsync(out_g, 0, bytes, s->ctx.stream));
    } else if (int rc = launch(g, s, out_g)) {
      return rc;
    }
    launches += s->ctx.launches - before;
  }
  const size_t host_need = host_reduce ? bytes * (size_t)(G + 1) : bytes;
  if (sc.host_bytes < host_need) {
    if (sc.host_out) SMC_CUDA(cudaFreeHost(sc.host_out));
    sc.host_out = nullptr;
    const size_t want = host_need < 4096 ? 4096 : host_need;
    SMC_CUDA(cudaHostAlloc(&sc.host_out, want, cudaHostAllocPortable | cudaHostAllocMapped));
    sc.host_bytes = want;
  }
  if (host_reduce) {
    // G packed results -> pinned host memory behind the reduced vector, summed in
    // shard order: bit-reproducible for a given shard count
    for (int g = 0; g < G; ++g) {
      Shard* s = sc.shards[g];
      SMC_CUDA(cudaSetDevice(s->device));
      SMC_CUDA(cudaMemcpyAsync(sc.host_out + (size_t)(g + 1) * n_out, s->out_dev, bytes,
                               cudaMemcpyDeviceToHost, s->ctx.stream));
    }
    if (int rc = synchronize_shards()) return rc;
    for (int j = 0; j < n_out; ++j) {
      double v = 0.0;
      for (int g = 0; g < G; ++g) v += sc.host_out[(size_t)(g + 1) * n_out + j];
      sc.host_out[j] = v;
    }
  } else {
    SMC_NCCL(sc.nccl.GroupStart());
    for (int g = 0; g < G; ++g) {
      Shard* s = sc.shards[g];
      SMC_NCCL(sc.nccl.AllReduce(s->out_dev, s->out_dev, (size_t)n_out, ncclDouble, ncclSum,
                                 s->comm, s->ctx.stream));
    }
    SMC_NCCL(sc.nccl.GroupEnd());
    Shard* s0 = sc.shards[0];
    SMC_CUDA(cudaSetDevice(s0->device));
    SMC_CUDA(cudaMemcpyAsync(sc.host_out, s0->out_dev, bytes, cudaMemcpyDeviceToHost,
                             s0->ctx.stream));
    SMC_CUDA(cudaStreamSynchronize(s0->ctx.stream));
    Context& cur = ctx();
    if (cur.inited) SMC_CUDA(cudaSetDevice(cur.device));
  }
  own_ctx().launches += launches + (host_reduce ? 0 : G);
  *out = sc.host_out;
  return SMC_OK;
}

static int check_partition(const smc_matrix* x,
                           std::initializer_list<const smc_matrix*> ops) {
  // every per-row operand follows the row partition of x
  for (const smc_matrix* m : ops)
    if (m && (m->shards.empty() || !same_partition(m, x)))
      return fail(SMC_ERR_INVALID_ARGUMENT,
                  "a per-row operand is not sharded like x (create it with "
                  "smc_matrix_create_like)");
  return SMC_OK;
}

// One evaluation of a memory-bound GLM family over the shards of c.x.  The small
// parameters travel as kernel arguments of every shard's launch.
int run_sharded(GlmCall& c, int n_out, const double** out) {
  if (int rc = check_partition(c.x, {c.y, c.alpha_vec, c.aux_vec, c.d_alpha_vec, c.d_aux_vec,
                                     c.d_y_vec, c.d_x}))
    return rc;
  return run_over_shards(c.x, n_out, [&](int g, Shard* sh, double* out_g) {
    GlmCall cg = c;
    auto pick = [g](const smc_matrix* m) {
      return m ? const_cast<smc_matrix*>(m->shards[g]) : nullptr;
    };
    cg.x = pick(c.x);
    cg.y = pick(c.y);
    cg.alpha_vec = pick(c.alpha_vec);
    cg.aux_vec = pick(c.aux_vec);
    cg.d_alpha_vec = pick(c.d_alpha_vec);
    cg.d_aux_vec = pick(c.d_aux_vec);
    cg.d_y_vec = pick(c.d_y_vec);
    cg.d_x = pick(c.d_x);
    cg.done_flag = sh->done_flag;  // direct mode only, else NULL
    cg.done_val = sh->done_val;
    cg.once_terms = g == 0;  // terms the reference adds once per call, not per row
    cg.out = out_g;
    return launch_glm(cg);
  }, out, /*allow_direct=*/true);
}

// The categorical GLM over the shards of x: beta (K x C) and alpha (C) are copied to
// every shard (params_host: K C + C doubles, beta first; the device form takes any
// pointer a cudaMemcpyDefault can read), the packed result is
// [logp, #non-finite, d_alpha[C], d_beta[K x C]].
int run_sharded_categorical(const smc_matrix* y, int y_scalar, const smc_matrix* x,
                            const double* params_host, int64_t C, unsigned flags,
                            smc_matrix* d_x, const double** out) {
  if (int rc = check_partition(x, {y, d_x})) return rc;
  const int n_out = (int)(2 + C + x->cols * C);
  return run_over_shards(x, n_out, [&](int g, Shard*, double* out_g) {
    return smc_categorical_logit_glm_device(y ? y->shards[g] : nullptr, y_scalar,
                                            x->shards[g], params_host, C, flags, out_g,
                                            d_x ? d_x->shards[g] : nullptr);
  }, out);
}

}  // namespace smc

using namespace smc;

extern "C" {

int smc_shard_init(int n_shards, const int* devices) {
  ShardComm& sc = comm();
  std::lock_guard<std::mutex> lock(sc.mu);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(SMC_ERR_CUDA, "no CUDA device available (%s); libstanmath_cuda has no CPU "
                "fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (n_shards <= 0) n_shards = ndev;  // every visible GPU
  if (n_shards > 64) return fail(SMC_ERR_INVALID_ARGUMENT, "at most 64 shards");
  std::vector<int> dev((size_t)n_shards);
  bool distinct = true;
  for (int g = 0; g < n_shards; ++g) {
    dev[g] = devices ? devices[g] : g % ndev;
    if (dev[g] < 0 || dev[g] >= ndev)
      return fail(SMC_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", dev[g], ndev);
    for (int h = 0; h < g; ++h) distinct = distinct && dev[h] != dev[g];
  }
  if (!sc.shards.empty()) shutdown_locked(sc);
  for (int g = 0; g < n_shards; ++g) {
    Shard* s = new Shard();
    s->device = dev[g];
    sc.shards.push_back(s);
    if (int rc = init_context(s->ctx, dev[g])) {
      shutdown_locked(sc);
      return rc;
    }
  }
  const char* mode = getenv("SMC_SHARD_REDUCE");
  const bool want_host = mode && strcmp(mode, "host") == 0;
  sc.use_nccl = false;
  if (!want_host && distinct && sc.nccl.load()) {
    std::vector<ncclComm_t> comms((size_t)n_shards);
    ncclResult_t r = sc.nccl.CommInitAll(comms.data(), n_shards, dev.data());
    if (r == ncclSuccess) {
      for (int g = 0; g < n_shards; ++g) sc.shards[g]->comm = comms[g];
      sc.use_nccl = true;
    } else if (mode && strcmp(mode, "nccl") == 0) {
      shutdown_locked(sc);
      return fail(SMC_ERR_CUDA, "ncclCommInitAll failed: %s", sc.nccl.GetErrorString(r));
    }
  } else if (mode && strcmp(mode, "nccl") == 0) {
    shutdown_locked(sc);
    return fail(SMC_ERR_CUDA, "SMC_SHARD_REDUCE=nccl: %s",
                distinct ? "libnccl.so.2 could not be loaded"
                         : "several shards share one GPU");
  }
  // small results (the memory-bound families: K + O(1) doubles) go through the direct
  // slots unless a mode is forced; larger ones (the categorical K x C gradient) through
  // NCCL, or through the host when there is no NCCL
  sc.direct = !mode || strcmp(mode, "direct") == 0;
  sc.reduce_mode = std::string(sc.direct ? "direct+" : "") + (sc.use_nccl ? "nccl" : "host");
  Context& cur = ctx();
  if (cur.inited) cudaSetDevice(cur.device);
  return SMC_OK;
}

int smc_shard_count(int* n_shards) {
  if (!n_shards) return fail(SMC_ERR_INVALID_ARGUMENT, "n_shards is NULL");
  *n_shards = (int)comm().shards.size();
  return SMC_OK;
}

const char* smc_shard_reduce_mode(void) { return comm().reduce_mode.c_str(); }

int smc_shard_shutdown(void) {
  ShardComm& sc = comm();
  std::lock_guard<std::mutex> lock(sc.mu);
  shutdown_locked(sc);
  return SMC_OK;
}

int smc_shard_synchronize(void) { return synchronize_shards(); }

static int make_sharded(int64_t rows, int64_t cols, int dtype,
                        const std::vector<int64_t>& shard_rows, smc_matrix** out) {
  ShardComm& sc = comm();
  smc_matrix* m = new smc_matrix();
  m->rows = rows;
  m->cols = cols;
  m->dtype = dtype;
  m->ld = 0;
  m->device = -1;
  m->id = next_matrix_id();
  int64_t row0 = 0;
  for (size_t g = 0; g < sc.shards.size(); ++g) {
    ShardScope scope(sc.shards[g]);
    smc_matrix* piece = nullptr;
    if (int rc = smc_matrix_create(shard_rows[g], cols, dtype, &piece)) {
      smc_matrix_free(m);
      return rc;
    }
    m->shards.push_back(piece);
    m->shard_row0.push_back(row0);
    row0 += shard_rows[g];
  }
  *out = m;
  return SMC_OK;
}

int smc_sharded_matrix_create(int64_t rows, int64_t cols, int dtype, smc_matrix** out) {
  if (!out) return fail(SMC_ERR_INVALID_ARGUMENT, "out is NULL");
  if (rows < 0 || cols < 0 || (dtype != SMC_F64 && dtype != SMC_I32))
    return fail(SMC_ERR_INVALID_ARGUMENT, "bad matrix shape/dtype %lld x %lld",
                (long long)rows, (long long)cols);
  ShardComm& sc = comm();
  if (sc.shards.empty())
    return fail(SMC_ERR_INVALID_ARGUMENT, "call smc_shard_init before creating sharded matrices");
  const int64_t G = (int64_t)sc.shards.size();
  std::vector<int64_t> shard_rows((size_t)G);
  for (int64_t g = 0; g < G; ++g)
    shard_rows[g] = rows * (g + 1) / G - rows * g / G;  // [g N / G, (g+1) N / G)
  return make_sharded(rows, cols, dtype, shard_rows, out);
}

int smc_matrix_create_like(const smc_matrix* like, int64_t cols, int dtype,
                           smc_matrix** out) {
  if (!like || !out) return fail(SMC_ERR_INVALID_ARGUMENT, "NULL argument");
  if (cols < 0) cols = like->cols;
  if (dtype < 0) dtype = like->dtype;
  if (like->shards.empty()) return smc_matrix_create(like->rows, cols, dtype, out);
  if (like->shards.size() != comm().shards.size())
    return fail(SMC_ERR_INVALID_ARGUMENT, "the shard set changed since `like` was created");
  std::vector<int64_t> shard_rows;
  for (const smc_matrix* p : like->shards) shard_rows.push_back(p->rows);
  return make_sharded(like->rows, cols, dtype, shard_rows, out);
}

int smc_matrix_view(const smc_matrix* src, smc_matrix** out) {
  if (!src || !out) return fail(SMC_ERR_INVALID_ARGUMENT, "NULL argument");
  if (src->shards.empty()) {
    if (int rc = realize(src)) return rc;
    return smc_matrix_wrap(src->data, src->rows, src->cols, src->ld, src->dtype, out);
  }
  smc_matrix* m = new smc_matrix();
  m->rows = src->rows;
  m->cols = src->cols;
  m->dtype = src->dtype;
  m->device = -1;
  m->id = next_matrix_id();
  m->shard_row0 = src->shard_row0;
  int rc = for_each_shard(src, [&](int, smc_matrix* p, int64_t) {
    smc_matrix* v = nullptr;
    if (int r = realize(p)) return r;
    if (int r = smc_matrix_wrap(p->data, p->rows, p->cols, p->ld, p->dtype, &v)) return r;
    m->shards.push_back(v);
    return (int)SMC_OK;
  });
  if (rc) {
    m->shard_row0.resize(m->shards.size());
    smc_matrix_free(m);
    return rc;
  }
  *out = m;
  return SMC_OK;
}

int smc_matrix_shard_count(const smc_matrix* m) { return shard_count_of(m); }

int smc_matrix_shard(const smc_matrix* m, int g, smc_matrix** shard, int64_t* row0,
                     int* device) {
  if (!m || g < 0 || g >= (int)m->shards.size())
    return fail(SMC_ERR_INVALID_ARGUMENT, "shard index out of range");
  if (shard) *shard = m->shards[g];
  if (row0) *row0 = m->shard_row0[g];
  if (device) *device = m->shards[g]->device;
  return SMC_OK;
}

}  // extern "C"
