// Microbenchmark: what DMMA rate does the inner loop of categorical pass 1 reach in
// isolation (no TMA, no epilogue)?  16 warps per CTA, one CTA per SM, fragments read
// from shared memory in the kernel's pattern (per k4-step: one 16-byte A load, two
// 16-byte B loads, eight DMMAs on eight accumulator pairs).
//   mode 0: DMMA from registers only      mode 1: + LDS fragment loads
//   mode 2: + the per-stage slot release (syncwarp, fence, shared atomic)
//   mode 3: mode 1 with two k4-steps' fragments loaded ahead (16 DMMAs per batch)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_loop dmma_loop.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

constexpr int KS = 16, BOX = 132, C8P = 36, NP = 2, NT = 4;

template <int MODE>
__global__ void __launch_bounds__(512, 1) loop_kernel(int iters, double* out, int warps_active) {
  extern __shared__ double smem[];
  __shared__ int rel_cnt[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = lane >> 2, tig = lane & 3;
  constexpr int xbox = BOX * KS, stage = 2 * xbox + C8P * KS;
  for (int i = tid; i < 4 * stage; i += blockDim.x) smem[i] = 1e-3 * (i % 97);
  if (tid < 4) rel_cnt[tid] = 0;
  __syncthreads();
  if (warp >= warps_active) return;
  double acc[2][NT][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
  const int xoff = (warp >> 3) * xbox + 16 * (warp & 7) + 2 * grp;
  int st = 0;
  for (int it = 0; it < iters; ++it) {
    const double* stg = smem + st * stage;
    const double* xa = stg + xoff + tig * BOX;
    const double* bfrag = stg + 2 * xbox + tig * C8P + 2 * grp;
    if (MODE == 0) {
      const double a0 = 1.0 + it, b0 = 0.5;
#pragma unroll
      for (int h = 0; h < KS / 4; ++h)
#pragma unroll
        for (int pr = 0; pr < NP; ++pr) {
          dmma(acc[0][2 * pr][0], acc[0][2 * pr][1], a0, b0);
          dmma(acc[1][2 * pr][0], acc[1][2 * pr][1], a0, b0);
          dmma(acc[0][2 * pr + 1][0], acc[0][2 * pr + 1][1], a0, b0);
          dmma(acc[1][2 * pr + 1][0], acc[1][2 * pr + 1][1], a0, b0);
        }
    } else if (MODE == 3) {
#pragma unroll
      for (int h2 = 0; h2 < KS / 8; ++h2) {
        double2 af[2], bf[2][NP];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int h = 2 * h2 + u;
          af[u] = *reinterpret_cast<const double2*>(xa + 4 * h * BOX);
#pragma unroll
          for (int pr = 0; pr < NP; ++pr)
            bf[u][pr] = *reinterpret_cast<const double2*>(bfrag + 4 * h * C8P + 16 * pr);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int pr = 0; pr < NP; ++pr) {
            dmma(acc[0][2 * pr][0], acc[0][2 * pr][1], af[u].x, bf[u][pr].x);
            dmma(acc[1][2 * pr][0], acc[1][2 * pr][1], af[u].y, bf[u][pr].x);
            dmma(acc[0][2 * pr + 1][0], acc[0][2 * pr + 1][1], af[u].x, bf[u][pr].y);
            dmma(acc[1][2 * pr + 1][0], acc[1][2 * pr + 1][1], af[u].y, bf[u][pr].y);
          }
      }
    } else {
#pragma unroll
      for (int h = 0; h < KS / 4; ++h) {
        const double2 af = *reinterpret_cast<const double2*>(xa + 4 * h * BOX);
        double2 bf[NP];
#pragma unroll
        for (int pr = 0; pr < NP; ++pr)
          bf[pr] = *reinterpret_cast<const double2*>(bfrag + 4 * h * C8P + 16 * pr);
#pragma unroll
        for (int pr = 0; pr < NP; ++pr) {
          dmma(acc[0][2 * pr][0], acc[0][2 * pr][1], af.x, bf[pr].x);
          dmma(acc[1][2 * pr][0], acc[1][2 * pr][1], af.y, bf[pr].x);
          dmma(acc[0][2 * pr + 1][0], acc[0][2 * pr + 1][1], af.x, bf[pr].y);
          dmma(acc[1][2 * pr + 1][0], acc[1][2 * pr + 1][1], af.y, bf[pr].y);
        }
      }
    }
    if (MODE == 2) {
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(&rel_cnt[st], 1) == warps_active - 1) {
          rel_cnt[st] = 0;
          __threadfence_block();
        }
      }
    }
    if (++st == 4) st = 0;
  }
  double s = 0;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) s += acc[mt][nt][0] + acc[mt][nt][1];
  if (s == 12345.678) out[0] = s;
}

template <int MODE>
void run(int sms, int warps, double* out) {
  const int iters = 20000;
  const size_t smem = 4 * (2 * BOX * KS + C8P * KS) * 8;
  cudaFuncSetAttribute(loop_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  loop_kernel<MODE><<<sms, 512, smem>>>(100, out, warps);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  loop_kernel<MODE><<<sms, 512, smem>>>(iters, out, warps);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 512.0 * 32 * iters * warps * sms;
  printf("{\"mode\": %d, \"warps\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"err\": \"%s\"}\n", MODE, warps,
         ms, flop / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double* out;
  cudaMalloc(&out, 8);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int warps : {4, 8, 16}) {
    run<0>(sms, warps, out);
    run<1>(sms, warps, out);
    run<2>(sms, warps, out);
    run<3>(sms, warps, out);
  }
  return 0;
}
