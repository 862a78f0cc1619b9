// Microbenchmark behind the d_x store stream of config 4: out[i, k] =
// beta[k] * d[i] for N = 1e7 rows, K = 128 columns, column-major with leading dimension
// ld -- 10.24 GB of pure stores.  cudaMemset / a fill kernel reach 7.4 TB/s on this GPU
// (profiles/micro/store_ceiling.py); the shipped `outer_kernel` (variant 0: two rows per
// thread, all K columns in the inner loop) reaches 6.0.  Which store ORDER gets closer?
//   0  shipped: thread = 2 rows, loop over the K columns           (CTA: 4 KB run per column)
//   1  thread = 2 rows x J row groups, loop k outer / j inner       (CTA: J * 4 KB run per column)
//   2  column panels: the grid sweeps PANEL columns at a time over all rows, d re-read
//      per panel (80 MB, L2-resident)
//   3  TMA: per-warp shared-memory tile {32 rows x 32 columns} -> one bulk tensor store
//   4  TMA: CTA tile {256 rows x 32 columns} (64 KB), double-buffered, one store per tile
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcuda -o outer_store outer_store.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

constexpr int K = 128;
struct Beta {
  double b[K];
};

__global__ void __launch_bounds__(256) v0(double* out, int64_t ld, const double* d, int64_t N,
                                          const __grid_constant__ Beta be) {
  const int64_t npairs = N / 2;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs;
       p += (int64_t)gridDim.x * blockDim.x) {
    const double2 dd = *reinterpret_cast<const double2*>(d + 2 * p);
    double* o = out + 2 * p;
#pragma unroll 8
    for (int k = 0; k < K; ++k)
      *reinterpret_cast<double2*>(o + (int64_t)k * ld) = make_double2(be.b[k] * dd.x, be.b[k] * dd.y);
  }
}

// variant 0 with streaming (evict-first) stores
__global__ void __launch_bounds__(256) v0cs(double* out, int64_t ld, const double* d, int64_t N,
                                            const __grid_constant__ Beta be) {
  const int64_t npairs = N / 2;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs;
       p += (int64_t)gridDim.x * blockDim.x) {
    const double2 dd = *reinterpret_cast<const double2*>(d + 2 * p);
    double* o = out + 2 * p;
#pragma unroll 8
    for (int k = 0; k < K; ++k)
      __stcs(reinterpret_cast<double2*>(o + (int64_t)k * ld), make_double2(be.b[k] * dd.x, be.b[k] * dd.y));
  }
}

// row-major order of the work: a warp writes ONE row pair... no: the transposed sweep --
// thread = column pair is impossible (column-major); instead 4 rows per thread (32-byte
// runs per thread, 1 KB per warp instruction pair)
__global__ void __launch_bounds__(256) v0x4(double* out, int64_t ld, const double* d, int64_t N,
                                            const __grid_constant__ Beta be) {
  const int64_t nq = N / 4;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nq;
       p += (int64_t)gridDim.x * blockDim.x) {
    const double4 dd = *reinterpret_cast<const double4*>(d + 4 * p);
    double* o = out + 4 * p;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const double b = be.b[k];
      double* ok = o + (int64_t)k * ld;
      asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(ok), "d"(b * dd.x), "d"(b * dd.y),
                   "d"(b * dd.z), "d"(b * dd.w)
                   : "memory");
    }
  }
}

template <int J>
__global__ void __launch_bounds__(256) v1(double* out, int64_t ld, const double* d, int64_t N,
                                          const __grid_constant__ Beta be) {
  const int64_t rows_per_cta = 512 * J;
  for (int64_t r0 = blockIdx.x * rows_per_cta; r0 < N; r0 += (int64_t)gridDim.x * rows_per_cta) {
    double2 dd[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int64_t i = r0 + j * 512 + 2 * threadIdx.x;
      dd[j] = i + 1 < N ? *reinterpret_cast<const double2*>(d + i) : make_double2(0, 0);
    }
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
      const double b = be.b[k];
      double* o = out + (int64_t)k * ld + r0 + 2 * threadIdx.x;
#pragma unroll
      for (int j = 0; j < J; ++j)
        if (r0 + j * 512 + 2 * threadIdx.x + 1 < N)
          *reinterpret_cast<double2*>(o + j * 512) = make_double2(b * dd[j].x, b * dd[j].y);
    }
  }
}

template <int PANEL>
__global__ void __launch_bounds__(256) v2(double* out, int64_t ld, const double* d, int64_t N,
                                          const __grid_constant__ Beta be) {
  const int64_t npairs = N / 2;
  for (int k0 = 0; k0 < K; k0 += PANEL) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs;
         p += (int64_t)gridDim.x * blockDim.x) {
      const double2 dd = *reinterpret_cast<const double2*>(d + 2 * p);
      double* o = out + 2 * p + (int64_t)k0 * ld;
#pragma unroll
      for (int k = 0; k < PANEL; ++k)
        *reinterpret_cast<double2*>(o + (int64_t)k * ld)
            = make_double2(be.b[k0 + k] * dd.x, be.b[k0 + k] * dd.y);
    }
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(smem_u32(src))
               : "memory");
}

// per-warp tiles: {32 rows x CW columns}, two slots per warp
template <int CW>
__global__ void __launch_bounds__(512) v3(const __grid_constant__ CUtensorMap tm, const double* d,
                                          int64_t N, const __grid_constant__ Beta be) {
  extern __shared__ __align__(128) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double* slots = sm + (size_t)warp * 2 * 32 * CW;
  const int64_t nblk = (N + 31) / 32;
  int s = 0;
  for (int64_t b = (int64_t)blockIdx.x * nw + warp; b < nblk; b += (int64_t)gridDim.x * nw) {
    const int64_t i = b * 32 + lane;
    const double dv = i < N ? d[i] : 0.0;
    for (int k0 = 0; k0 < K; k0 += CW) {
      double* t = slots + s * 32 * CW;
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
#pragma unroll 8
      for (int k = 0; k < CW; ++k) t[k * 32 + lane] = be.b[k0 + k] * dv;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tm, (int)(b * 32), k0, t);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      s ^= 1;
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// CTA tiles: {256 rows x 32 columns} = 64 KB, two slots; 256 threads, thread = row
__global__ void __launch_bounds__(256) v4(const __grid_constant__ CUtensorMap tm, const double* d,
                                          int64_t N, const __grid_constant__ Beta be) {
  extern __shared__ __align__(128) double sm[];
  constexpr int CW = 32;
  const int64_t nblk = (N + 255) / 256;
  int s = 0;
  for (int64_t b = blockIdx.x; b < nblk; b += gridDim.x) {
    const int64_t i = b * 256 + threadIdx.x;
    const double dv = i < N ? d[i] : 0.0;
    for (int k0 = 0; k0 < K; k0 += CW) {
      double* t = sm + (size_t)s * 256 * CW;
      if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < CW; ++k) t[k * 256 + threadIdx.x] = be.b[k0 + k] * dv;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (threadIdx.x == 0) {
        tma_store_2d(&tm, (int)(b * 256), k0, t);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      s ^= 1;
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
  const int64_t N = 10000000, ld = N;  // (N is a multiple of 16: the layout smc_matrix_create makes)
  double *out, *d;
  CK(cudaMalloc(&out, sizeof(double) * ld * K));
  CK(cudaMalloc(&d, sizeof(double) * N));
  std::vector<double> dh(N);
  for (int64_t i = 0; i < N; ++i) dh[i] = 1.0 + (i % 97) * 0.01;
  CK(cudaMemcpy(d, dh.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
  Beta be;
  for (int k = 0; k < K; ++k) be.b[k] = 0.5 + k;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<double> chk(4);
  auto time = [&](const char* name, auto launch) {
    CK(cudaMemset(out, 0, sizeof(double) * ld * K));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
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
    // spot check: first, an interior and the last element
    const int64_t idx[4] = {0, 5 * ld + 12345, 77 * ld + N - 1, (int64_t)(K - 1) * ld + N - 1};
    const int64_t row[4] = {0, 12345, N - 1, N - 1};
    const int col[4] = {0, 5, 77, K - 1};
    bool ok = true;
    for (int j = 0; j < 4; ++j) {
      double v;
      CK(cudaMemcpy(&v, out + idx[j], 8, cudaMemcpyDeviceToHost));
      ok = ok && v == be.b[col[j]] * dh[row[j]];
    }
    printf("{\"variant\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f, \"ok\": %s}\n", name, best,
           N * K * 8.0 / best / 1e6, ok ? "true" : "false");
    fflush(stdout);
  };
  const int g0 = sms * 16;
  time("0 shipped (2 rows/thread, k inner)", [&] { v0<<<g0, 256>>>(out, ld, d, N, be); });
  time("0 shipped, 8 CTAs/SM", [&] { v0<<<sms * 8, 256>>>(out, ld, d, N, be); });
  time("0 with st.cs", [&] { v0cs<<<g0, 256>>>(out, ld, d, N, be); });
  time("0 with 32-byte stores (4 rows/thread)", [&] { v0x4<<<g0, 256>>>(out, ld, d, N, be); });
  time("1 J=4 (16 KB runs)", [&] { v1<4><<<sms * 8, 256>>>(out, ld, d, N, be); });
  time("1 J=8 (32 KB runs)", [&] { v1<8><<<sms * 8, 256>>>(out, ld, d, N, be); });
  time("1 J=8, 4 CTAs/SM", [&] { v1<8><<<sms * 4, 256>>>(out, ld, d, N, be); });
  time("2 panels of 8 columns", [&] { v2<8><<<sms * 8, 256>>>(out, ld, d, N, be); });
  time("2 panels of 32 columns", [&] { v2<32><<<sms * 8, 256>>>(out, ld, d, N, be); });
  time("2 panels of 2 columns", [&] { v2<2><<<sms * 8, 256>>>(out, ld, d, N, be); });
  {
    CUtensorMap tm = make_map(out, N, K, ld, 32, 32);
    const size_t smem = 12 * 2 * 32 * 32 * 8;
    CK(cudaFuncSetAttribute(v3<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    time("3 TMA per-warp 32x32 tiles, 12 warps", [&] { v3<32><<<sms, 384, smem>>>(tm, d, N, be); });
  }
  {
    CUtensorMap tm = make_map(out, N, K, ld, 256, 32);
    const size_t smem = 2 * 256 * 32 * 8;
    CK(cudaFuncSetAttribute(v4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    time("4 TMA CTA 256x32 tiles", [&] { v4<<<sms, 256, smem>>>(tm, d, N, be); });
  }
  return 0;
}
