// TMA / mbarrier PTX glue and tensor-map encoding shared by the TMA-fed kernels
// (glm_fused.cu, categorical.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

namespace smc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA: 2-D tiled bulk tensor load global -> shared, completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map,
                                            int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::"
      "bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1),
      "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// Bulk copy global -> shared (1-D, 16-byte aligned, a multiple of 16 bytes), completion
// on an mbarrier.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, "
      "[%3];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// TMA: 2-D tiled bulk tensor store shared -> global (bulk async-group
// completion); elements outside the tensor are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1,
                                             const void* src, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint "
      "[%0, {%1, %2}], [%3], %4;" ::"l"(reinterpret_cast<uint64_t>(map)),
      "r"(c0), "r"(c1), "r"(smem_u32(src)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d_plain(const CUtensorMap* map, int c0, int c1,
                                                   const void* src) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
          reinterpret_cast<uint64_t>(map)),
      "r"(c0), "r"(c1), "r"(smem_u32(src))
      : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// Every committed bulk store of this thread has finished READING shared memory.
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// Generic-proxy writes to shared memory become visible to the async proxy (TMA).
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void group_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void st_stream(double* p, double v) {
  asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}


// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t,
                                  void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault,
                                &q)
            == cudaSuccess
        && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}


// 2-D tiled map over a column-major f64 matrix: dim0 = rows (contiguous),
// dim1 = columns (stride ld doubles); box = {box_rows, box_cols}; out-of-range
// elements of a box read as zero.
inline CUresult encode_tmap_f64(CUtensorMap* map, const void* base, int64_t rows,
                                int64_t cols, int64_t ld, int box_rows,
                                int box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return CUDA_ERROR_NOT_SUPPORTED;
  cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
  if (cols == 1) gstr[0] = (cuuint64_t)((rows + 1) & ~1ll) * 8;
  cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
  cuuint32_t estr[2] = {1, 1};
  // (A/B switch, read once: SMC_TMAP_L2PROMO = 0 none, 1 64 B, 2 128 B, 3 256 B -- default)
  static const int promo = [] {
    const char* e = getenv("SMC_TMAP_L2PROMO");
    return e ? atoi(e) : 3;
  }();
  const CUtensorMapL2promotion pr = promo == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                    : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                    : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                 : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<void*>(base), gdim,
             gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace smc
