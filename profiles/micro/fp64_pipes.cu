// Microbenchmark: do FP64 DMMA (mma.sync.m8n8k4.f64) and FP64 DFMA share one execution
// resource on sm_100a, or can a CTA overlap them?  Three launches with the same per-warp
// work: DFMA warps alone, DMMA warps alone, both together.  If the two are independent
// pipes, "both" takes max(dfma, dmma); if they share the FP64 units, it takes the sum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode bit 0: warps [0, nw/2) run DFMA; bit 1: warps [nw/2, nw) run DMMA
__global__ void __launch_bounds__(512, 1) mix_kernel(int mode, int iters, double* out, double seed) {
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = seed * (i + threadIdx.x);
  const double a = seed + 1.0, b = seed * 0.5;
  if (warp < nw / 2) {
    if (!(mode & 1)) return;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)  // 8 x 16 DFMA = 128 per thread per iteration
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
  } else {
    if (!(mode & 2)) return;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 2; ++r)  // 2 x 8 DMMA = 16 x 256 FMA per warp = 128 FMA per thread
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(acc[2 * i], acc[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

int main() {
  double* out;
  cudaMalloc(&out, 8);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {256, 512}) {
    for (int mode : {1, 2, 3}) {
      mix_kernel<<<sms, threads>>>(mode, 100, out, 1e-9);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      mix_kernel<<<sms, threads>>>(mode, iters, out, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double warps = (threads / 32 / 2) * ((mode & 1) + ((mode >> 1) & 1));
      const double flop = 2.0 * 128 * 32 * iters * warps * sms;
      printf("{\"threads\": %d, \"mode\": \"%s\", \"ms\": %.3f, \"tflops\": %.2f}\n", threads,
             mode == 1 ? "dfma" : mode == 2 ? "dmma" : "both", ms, flop / ms * 1e-9);
    }
  }
  return 0;
}
