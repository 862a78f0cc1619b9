// What a pure READ stream reaches on this GPU, for the headline's roofline (DESIGN.md section
// 4.1): 20.48 GB (the N=1e7 x K=256 design matrix) read once and reduced to one number per
// CTA, with 16-byte loads, 32-byte loads (ld.global.v4.f64, sm_100) and TMA tile loads
// (cp.async.bulk.tensor.2d into a shared-memory ring, tiles {64 rows x 256 columns} as the
// fused GLM kernel requests them, and {256 rows x 32 columns}).  The fused kernel streams the
// same bytes at 6.9 - 7.1 TB/s.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_ceiling read_ceiling.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

__global__ void __launch_bounds__(256) read16(const double2* __restrict__ p, int64_t n2,
                                              double* out) {
  double s = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    const double2 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
    s += (a.x + a.y) + (b.x + b.y) + (c.x + c.y) + (d.x + d.y);
  }
  for (; i < n2; i += stride) s += p[i].x + p[i].y;
  if (s == 1.2345e300) out[blockIdx.x] = s;
}

__device__ __forceinline__ void ld_v4(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__global__ void __launch_bounds__(256) read32(const double* __restrict__ p, int64_t n4,
                                              double* out) {
  double s = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {
    double a0, a1, a2, a3, b0, b1, b2, b3;
    ld_v4(p + 4 * i, a0, a1, a2, a3);
    ld_v4(p + 4 * (i + stride), b0, b1, b2, b3);
    s += (a0 + a1) + (a2 + a3) + (b0 + b1) + (b2 + b3);
  }
  for (; i < n4; i += stride) {
    double a0, a1, a2, a3;
    ld_v4(p + 4 * i, a0, a1, a2, a3);
    s += (a0 + a1) + (a2 + a3);
  }
  if (s == 1.2345e300) out[blockIdx.x] = s;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// one elected thread keeps `stages` tile loads in flight; the other warps only touch one word
// per tile (the data is not consumed: this is the feed rate of the TMA path alone)
__global__ void __launch_bounds__(128) read_tma(const __grid_constant__ CUtensorMap tm, int box_r,
                                                int box_c, int64_t ntile_r, int ntile_c, int stages,
                                                double* out) {
  extern __shared__ __align__(128) unsigned char raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw);
  double* ring = reinterpret_cast<double*>(raw + 128);
  const size_t tile_d = (size_t)box_r * box_c;
  const uint32_t bytes = (uint32_t)(tile_d * 8);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t ntiles = ntile_r * ntile_c;
  const int64_t mine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto issue = [&](int64_t k) {
    const int64_t t = blockIdx.x + k * gridDim.x;
    const int s = (int)(k % stages);
    const int r0 = (int)((t / ntile_c) * box_r), c0 = (int)((t % ntile_c) * box_c);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + s)),
                 "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, "
        "%3}], [%4];" ::"r"(smem_u32(ring + (size_t)s * tile_d)),
        "l"(reinterpret_cast<uint64_t>(&tm)), "r"(r0), "r"(c0), "r"(smem_u32(bars + s))
        : "memory");
  };
  if (threadIdx.x == 0)
    for (int64_t k = 0; k < stages && k < mine; ++k) issue(k);
  double acc = 0;
  for (int64_t k = 0; k < mine; ++k) {
    const int s = (int)(k % stages);
    mbar_wait(bars + s, (uint32_t)((k / stages) & 1));
    acc += ring[(size_t)s * tile_d + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0 && k + stages < mine) issue(k + stages);
  }
  if (acc == 1.2345e300) out[blockIdx.x] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make_map(void* base, int64_t rows, int64_t cols, int64_t ld, int br, int bc) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  CUtensorMap m;
  cuuint64_t gd[2] = {(cuuint64_t)rows, (cuuint64_t)cols}, gs[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)br, (cuuint32_t)bc}, es[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(p)(
      &m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "encode failed %d\n", (int)r);
    exit(1);
  }
  return m;
}

int main() {
  const int64_t N = 10000000, K = 256, n = N * K;
  double *x, *out;
  CK(cudaMalloc(&x, sizeof(double) * n));
  CK(cudaMalloc(&out, sizeof(double) * 65536));
  CK(cudaMemset(x, 0, sizeof(double) * n));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto time = [&](const char* name, auto launch) {
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float best = 1e9f;
    for (int r = 0; r < 10; ++r) {
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      best = ms < best ? ms : best;
    }
    CK(cudaGetLastError());
    printf("{\"variant\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f}\n", name, best, n * 8.0 / best / 1e6);
    fflush(stdout);
  };
  for (int per : {8, 16, 32})
    time(per == 8 ? "16-byte loads, 8 CTAs/SM" : per == 16 ? "16-byte loads, 16 CTAs/SM" : "16-byte loads, 32 CTAs/SM",
         [&] { read16<<<sms * per, 256>>>(reinterpret_cast<const double2*>(x), n / 2, out); });
  for (int per : {8, 16, 32})
    time(per == 8 ? "32-byte loads, 8 CTAs/SM" : per == 16 ? "32-byte loads, 16 CTAs/SM" : "32-byte loads, 32 CTAs/SM",
         [&] { read32<<<sms * per, 256>>>(x, n / 4, out); });
  {
    CUtensorMap tm = make_map(x, N, K, N, 64, 256);  // 128 KB tiles... too big for 3 stages: 64 x 128
    (void)tm;
  }
  struct Shape {
    int br, bc, stages;
    const char* name;
  } shapes[] = {{64, 128, 3, "TMA {64 x 128} tiles (64 KB), 3 stages"},
                {32, 256, 3, "TMA {32 x 256} tiles (64 KB), 3 stages"},
                {256, 32, 3, "TMA {256 x 32} tiles (64 KB), 3 stages"},
                {64, 64, 6, "TMA {64 x 64} tiles (32 KB), 6 stages"},
                {128, 256, 1, "TMA {128 x 256}... skipped"}};
  CK(cudaFuncSetAttribute(read_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (const Shape& sh : shapes) {
    if (sh.stages == 1) continue;
    CUtensorMap tm = make_map(x, N, K, N, sh.br, sh.bc);
    const int64_t ntr = (N + sh.br - 1) / sh.br;
    const int ntc = (int)((K + sh.bc - 1) / sh.bc);
    const size_t smem = 128 + (size_t)sh.stages * sh.br * sh.bc * 8;
    time(sh.name, [&] { read_tma<<<sms, 128, smem>>>(tm, sh.br, sh.bc, ntr, ntc, sh.stages, out); });
  }
  return 0;
}
