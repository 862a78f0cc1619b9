// Microbenchmark: how should the softmax epilogue of categorical pass 1 share an SM with
// the DMMA main loop?  Same per-SM work in every mode (256-row blocks, K = 512 in 32
// stages of 16 attributes, 32 classes; per row block 32 exp per row pair ... as in the
// kernel), no TMA (fragments from a static shared-memory ring).
//   mode 0: main loop only (16 warps x 16 rows), no epilogue           -> ceiling
//   mode 1: + the epilogue between row blocks, by the same warps        (round-1 kernel)
//   mode 2: + the epilogue in slices after the first stages of the next block, same warps
//   mode 3: warp-specialised: 8 DMMA warps x 32 rows + 8 epilogue warps, accumulators
//           handed over through shared memory
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_epilogue dmma_epilogue.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double quad_sum(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ double quad_max(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

constexpr int KS = 16, BOX = 132, C8P = 36, NP = 2, NT = 4, STAGES = 32;
constexpr int xbox = BOX * KS, stage_d = 2 * xbox + C8P * KS;

// softmax of one row (8 values per lane of a quad), result stored to out
__device__ __forceinline__ double softmax_row(double* v, double* outp) {
  double m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) m = fmax(m, v[i]);
  m = quad_max(m);
  double se = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = exp(v[i] - m);
    se += v[i];
  }
  se = quad_sum(se);
  const double inv = 1.0 / se;
#pragma unroll
  for (int i = 0; i < 8; ++i) outp[i * 32] = v[i] * -inv;
  return log(inv) - m;
}

__device__ __forceinline__ float softmax_row_f(float* v, double* outp) {
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) m = fmaxf(m, v[i]);
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  float se = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = expf(v[i] - m);
    se += v[i];
  }
  se += __shfl_xor_sync(0xffffffffu, se, 1);
  se += __shfl_xor_sync(0xffffffffu, se, 2);
  const float inv = 1.0f / se;
#pragma unroll
  for (int i = 0; i < 8; ++i) outp[i * 32] = v[i] * -inv;
  return logf(inv) - m;
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) epi_kernel(int blocks, double* out) {
  extern __shared__ double smem[];
  __shared__ int rel_cnt[4];
  __shared__ unsigned long long bars[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, grp = lane >> 2, tig = lane & 3;
  double* stash = smem + 2 * stage_d;  // [16 warps][16][32] (modes 2, 3)
  for (int i = tid; i < 2 * stage_d; i += blockDim.x) smem[i] = 1e-3 * (i % 97);
  if (tid < 4) rel_cnt[tid] = 0;
  if (tid < 32) bars[tid] = 0;
  __syncthreads();
  volatile unsigned long long* full_f = bars;        // [8] blocks written by DMMA warp i
  volatile unsigned long long* free_f = bars + 8;    // [8] blocks read by epilogue warp i
  double lp = 0;
  int st = 0;

  if (MODE == 3) {
    if (warp < 8) {
      // DMMA warp: 32 rows = 4 row tiles (two pairs), 4 class tiles
      double acc[4][NT][2];
      const int xoff = (warp >> 2) * xbox + 32 * (warp & 3) + 2 * grp;
      double* my = stash + warp * 1024 + lane;
      for (int b = 0; b < blocks; ++b) {
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
        for (int ks = 0; ks < STAGES; ++ks) {
          const double* stg = smem + st * stage_d;
          const double* xa = stg + xoff + tig * BOX;
          const double* bfrag = stg + 2 * xbox + tig * C8P + 2 * grp;
#pragma unroll
          for (int h = 0; h < KS / 4; ++h) {
            const double2 a0 = *reinterpret_cast<const double2*>(xa + 4 * h * BOX);
            const double2 a1 = *reinterpret_cast<const double2*>(xa + 4 * h * BOX + 16);
            double2 bf[NP];
#pragma unroll
            for (int pr = 0; pr < NP; ++pr)
              bf[pr] = *reinterpret_cast<const double2*>(bfrag + 4 * h * C8P + 16 * pr);
#pragma unroll
            for (int pr = 0; pr < NP; ++pr) {
              dmma(acc[0][2 * pr][0], acc[0][2 * pr][1], a0.x, bf[pr].x);
              dmma(acc[1][2 * pr][0], acc[1][2 * pr][1], a0.y, bf[pr].x);
              dmma(acc[2][2 * pr][0], acc[2][2 * pr][1], a1.x, bf[pr].x);
              dmma(acc[3][2 * pr][0], acc[3][2 * pr][1], a1.y, bf[pr].x);
              dmma(acc[0][2 * pr + 1][0], acc[0][2 * pr + 1][1], a0.x, bf[pr].y);
              dmma(acc[1][2 * pr + 1][0], acc[1][2 * pr + 1][1], a0.y, bf[pr].y);
              dmma(acc[2][2 * pr + 1][0], acc[2][2 * pr + 1][1], a1.x, bf[pr].y);
              dmma(acc[3][2 * pr + 1][0], acc[3][2 * pr + 1][1], a1.y, bf[pr].y);
            }
          }
          __syncwarp();
          if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&rel_cnt[st], 1) == 7) {
              rel_cnt[st] = 0;
              __threadfence_block();
            }
          }
          if (++st == 2) st = 0;
        }
        // hand the accumulators over: wait until the partner has read block b - 1
        while (free_f[warp] < (unsigned long long)b) __nanosleep(64);
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) my[(8 * mt + 2 * nt + j) * 32] = acc[mt][nt][j];
        __threadfence_block();
        __syncwarp();
        if (lane == 0) full_f[warp] = b + 1;
      }
    } else {
      const int pw = warp - 8;
      double* my = stash + pw * 1024 + lane;
      double* myout = stash + 8 * 1024 + pw * 1024 + lane;
      for (int b = 0; b < blocks; ++b) {
        while (full_f[pw] < (unsigned long long)(b + 1)) __nanosleep(256);
        __threadfence_block();
        double v[4][8];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
#pragma unroll
          for (int i = 0; i < 8; ++i) v[mt][i] = my[(8 * mt + i) * 32];
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          free_f[pw] = b + 1;
        }
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) lp += softmax_row(v[mt], myout + 8 * mt * 32);
      }
    }
  } else {
    double acc[2][NT][2];
    const int xoff = (warp >> 3) * xbox + 16 * (warp & 7) + 2 * grp;
    double* my = stash + warp * 512 + lane;
    double p_m0 = 0, p_m1 = 0, p_s0 = 0, p_s1 = 0;
    auto slice = [&](int s) {
      if (s == 0) {
        double m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          m0 = fmax(m0, my[e * 32]);
          m1 = fmax(m1, my[(8 + e) * 32]);
        }
        p_m0 = quad_max(m0);
        p_m1 = quad_max(m1);
        p_s0 = p_s1 = 0;
      } else if (s <= 8) {
        const int e = s - 1;
        const double e0 = exp(my[e * 32] - p_m0), e1 = exp(my[(8 + e) * 32] - p_m1);
        my[e * 32] = e0;
        my[(8 + e) * 32] = e1;
        p_s0 += e0;
        p_s1 += e1;
      } else if (s == 9) {
        const double i0 = 1.0 / quad_sum(p_s0), i1 = 1.0 / quad_sum(p_s1);
        lp += (log(i0) - p_m0) + (log(i1) - p_m1);
        p_s0 = -i0;
        p_s1 = -i1;
      } else {
        const int e = 2 * (s - 10);
        my[e * 32] *= p_s0;
        my[(e + 1) * 32] *= p_s0;
        my[(8 + e) * 32] *= p_s1;
        my[(9 + e) * 32] *= p_s1;
      }
    };
    bool pending = false;
    for (int b = 0; b < blocks; ++b) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
      for (int ks = 0; ks < STAGES; ++ks) {
        const double* stg = smem + st * stage_d;
        const double* xa = stg + xoff + tig * BOX;
        const double* bfrag = stg + 2 * xbox + tig * C8P + 2 * grp;
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
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          if (atomicAdd(&rel_cnt[st], 1) == 15) {
            rel_cnt[st] = 0;
            __threadfence_block();
          }
        }
        if (++st == 2) st = 0;
        if (MODE == 2 && pending && ks < 14) slice(ks);
      }
      if (MODE == 5) {
        float v0[8], v1[8];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            v0[2 * nt + j] = (float)acc[0][nt][j];
            v1[2 * nt + j] = (float)acc[1][nt][j];
          }
        lp += softmax_row_f(v0, my);
        lp += softmax_row_f(v1, my + 8 * 32);
      }
      if (MODE == 1 || MODE == 4) {
        double v0[8], v1[8];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            v0[2 * nt + j] = acc[0][nt][j];
            v1[2 * nt + j] = acc[1][nt][j];
          }
        lp += softmax_row(v0, my);
        lp += softmax_row(v1, my + 8 * 32);
        if (MODE == 4) {
          lp += softmax_row(v0, my);
          lp += softmax_row(v1, my + 8 * 32);
        }
      }
      if (MODE == 2) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) my[(8 * mt + 2 * nt + j) * 32] = acc[mt][nt][j];
        pending = true;
      }
      if (MODE == 0) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) lp += acc[0][nt][0] + acc[0][nt][1] + acc[1][nt][0] + acc[1][nt][1];
      }
    }
  }
  if (lp == 12345.678) out[0] = lp;
}

template <int MODE>
void run(int sms, double* out) {
  const int blocks = 200;
  const size_t smem = (2 * stage_d + 16 * 1024) * 8;
  cudaFuncSetAttribute(epi_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  epi_kernel<MODE><<<sms, 512, smem>>>(2, out);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  epi_kernel<MODE><<<sms, 512, smem>>>(blocks, out);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 2.0 * 256 * 512 * 32 * blocks * sms;
  printf("{\"mode\": %d, \"ms\": %.3f, \"tflops\": %.2f, \"err\": \"%s\"}\n", MODE, ms,
         flop / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double* out;
  cudaMalloc(&out, 8);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>(sms, out);
  run<1>(sms, out);
  run<2>(sms, out);
  run<3>(sms, out);
  run<4>(sms, out);
  run<5>(sms, out);
  return 0;
}
