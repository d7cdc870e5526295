// Probe: can a tcgen05 A-operand descriptor address a SHIFTED window of a TMA-written, 128B-swizzled patch?
// patch: P rows x 64 fp16 (128 B rows, SWIZZLE_128B, written by one TMA 2D box load).
// MMA rows: group g (0..15), row i (0..7) -> patch row r0 + g*S + i.   D[128 x 64] = A_window * B^T (K = 64).
// Variants: base_offset = 0 | (start >> 7) & 7.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -shared ...
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni D;\n\tbra.uni W;\n\tD:\n\t}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

extern "C" __global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map_x,
                                                               const __grid_constant__ CUtensorMap map_w, int P, int r0,
                                                               int S, int use_base_offset, float* out /*[128,64]*/) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* patch = smem;                 // P x 128 B (<= 32 KB)
  uint8_t* wt = smem + 32768;            // 64 x 128 B
  __shared__ uint64_t bar, mbar;
  __shared__ uint32_t tmem_s;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (threadIdx.x == 0) {
    mbar_expect(&bar, (uint32_t)(P * 128 + 64 * 128));
    tma2d(&map_x, &bar, patch, 0, 0);
    tma2d(&map_w, &bar, wt, 0, 0);
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t a_start = smem_u32(patch) + (uint32_t)r0 * 128;
    const uint32_t b_start = smem_u32(wt);
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      uint32_t aa = a_start + k * 32, bb = b_start + k * 32;
      uint64_t adesc = (uint64_t)((aa & 0x3FFFF) >> 4) | ((uint64_t)((S * 128) >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
      if (use_base_offset) adesc |= (uint64_t)((aa >> 7) & 7) << 49;
      uint64_t bdesc = (uint64_t)((bb & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
      uint32_t acc = k ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  __syncthreads();
  mbar_wait(&mbar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = threadIdx.x;  // 4 warps x 32 lanes = 128 TMEM lanes
  for (int c = 0; c < 64; c += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[row * 64 + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int probe_run(const void* x /*[P,64] f16*/, const void* w /*[64,64] f16*/, int P, int r0, int S, int use_bo, float* out) {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return -1;
  EncFn enc = (EncFn)f;
  CUtensorMap mx, mw;
  cuuint64_t dx[2] = {64, (cuuint64_t)P}, dw[2] = {64, 64}, st[1] = {128};
  cuuint32_t bx[2] = {64, (cuuint32_t)P}, bw[2] = {64, 64}, one[2] = {1, 1};
  if (enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)x, dx, st, bx, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return -2;
  if (enc(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)w, dw, st, bw, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return -3;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  probe_kernel<<<1, 128, 32768 + 8192 + 1024>>>(mx, mw, P, r0, S, use_bo, out);
  return (int)cudaDeviceSynchronize();
}
