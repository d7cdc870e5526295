// Probe for round 2 (NOT part of libstep_b200.so; compiled and run by tools/probe/run_2cta_probe.py on a GPU box):
// D[256 x N] = A[256 x K] * B[N x K]^T with ONE tcgen05.mma.cta_group::2 stream issued by the leader CTA of a pair.
//   CTA r (r = %cluster_ctarank) stages rows [128 r, 128 r + 128) of A and rows [N/2 r, N/2 r + N/2) of B in ITS shared
//   memory (SWIZZLE_128B tiles, BK = 64) -> per SM the operand inbound is A/2 + B/2 instead of A/2 + B: the one change
//   that lowers the bytes an SM must pull in per MMA cycle.
//   The accumulator of CTA r (its own TMEM, lanes 0..127, columns 0..N-1) holds rows [128 r, 128 r + 128) of D.
// Synchronisation is deliberately simple (one stage, one k-block at a time); what the probe has to establish is the
// protocol itself: (1) both CTAs' TMA loads complete_tx on the LEADER's mbarrier (.cta_group::2, peer bit cleared),
// (2) smem descriptors are CTA-relative offsets valid in both CTAs, (3) tcgen05.commit.cta_group::2 ... multicast
// releases the stage in both CTAs, (4) idesc M = 256.
// Sources for the instruction forms: cute/arch/copy_sm100_tma.hpp (SM100_TMA_2SM_LOAD_2D), cutlass/arch/barrier.h
// (umma_arrive_multicast_2x1SM), /opt/skills/guides/blackwell_cuda_programming.md section "CTA Pair".
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n\t.reg .pred p;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni D2;\n\tbra.uni W2;\n\tD2:\n\t}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA load: data into THIS CTA's shared memory, transaction bytes onto the LEADER's barrier (peer bit 24 cleared)
__device__ __forceinline__ void tma2sm_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(m), "r"(mbar), "r"(c0), "r"(c1) : "memory");
}

constexpr int BK = 64;

extern "C" __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int N, int K,
              float* out /*[256, N]*/) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // [128 rows][128 B]
  uint8_t* sB = smem + 16384;         // [N/2 rows][128 B]
  __shared__ uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_s;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  uint32_t ncols = 32;
  while (ncols < (uint32_t)N) ncols <<= 1;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {   // both CTAs of the pair execute the paired allocation (cute::TMEM::Allocator2Sm)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();     // the peer's barriers exist before any load signals them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  const uint32_t a_bytes = 128 * BK * 2, b_bytes = (uint32_t)(N / 2) * BK * 2;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // M = 256 across the pair
  uint32_t phase = 0;
  for (int kb = 0; kb < K / BK; ++kb) {
    if (threadIdx.x == 0) {
      // leader expects the bytes of BOTH CTAs on its barrier; each CTA loads its own halves
      if (rank == 0) mbar_expect(&full_bar, 2 * (a_bytes + b_bytes));
      tma2sm_2d(&map_a, &full_bar, sA, kb * BK, (int)rank * 128);
      tma2sm_2d(&map_b, &full_bar, sB, kb * BK, (int)rank * (N / 2));
      if (rank == 0) {
        mbar_wait(&full_bar, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int k = 0; k < BK / 16; ++k) {
          const uint32_t aa = smem_u32(sA) + k * 32, bb = smem_u32(sB) + k * 32;
          const uint64_t adesc = (uint64_t)((aa & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
          const uint64_t bdesc = (uint64_t)((bb & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
          const uint32_t acc = (kb | k) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        }
        // release the stage / publish the accumulator in BOTH CTAs
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&done_bar)), "h"((uint16_t)3) : "memory");
      }
    }
    mbar_wait(&done_bar, phase);   // every thread of both CTAs: the MMAs that read sA / sB have retired
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    phase ^= 1u;
    __syncthreads();
  }
  // epilogue: this CTA's 128 lanes x N columns are rows [128 rank, 128 rank + 128) of D
  const int row = (int)rank * 128 + threadIdx.x;
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(size_t)row * N + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();     // neither CTA frees the paired allocation while the other still reads it
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int probe2_run(const void* a /*[256,K] f16*/, const void* b /*[N,K] f16*/, int N, int K, float* out) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return -1;
  EncFn enc = (EncFn)f;
  CUtensorMap ma, mb;
  cuuint64_t da[2] = {(cuuint64_t)K, 256}, db[2] = {(cuuint64_t)K, (cuuint64_t)N}, st[1] = {(cuuint64_t)K * 2};
  cuuint32_t ba[2] = {BK, 128}, bb[2] = {BK, (cuuint32_t)(N / 2)}, one[2] = {1, 1};
  if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)a, da, st, ba, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return -2;
  if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)b, db, st, bb, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return -3;
  const size_t smem = 16384 + (size_t)(N / 2) * BK * 2 + 1024;
  cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe2_kernel<<<2, 128, smem>>>(ma, mb, N, K, out);
  return (int)cudaDeviceSynchronize();
}
