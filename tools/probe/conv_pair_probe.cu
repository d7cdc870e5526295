// Round-2 experiment (NOT part of libstep_b200.so, NOT yet run on a GPU -- only compile-checked; the protocol it relies
// on is the one tools/probe/umma_2cta_probe.cu validated): a persistent 1x1x1-convolution / GEMM kernel on CTA PAIRS.
//
//   D[M, N] = relu(A[M, K] * B[N, K]^T * scale + shift)        fp16 in, fp32 accumulate, fp16 out
//
// A pair of CTAs (cluster of 2, the two SMs of a TPC) owns a 256 x BN output tile.  CTA r stages rows
// [256 t + 128 r, +128) of A and rows [n0 + r BN/2, + BN/2) of B: per k-block an SM pulls in 16 KB + BN/2 * 128 B
// (32 KB for BN = 256) for 512 MMA cycles = 64 B/clk, against 96 B/clk for the 128 x 256 single-CTA tile of
// conv_umma_persist_kernel -- below the ~70-90 B/clk one SM can take in.
// Structure per CTA: warp 0 producer, warp 1 MMA issuer (leader CTA only), warps 2..5 epilogue (own 128 rows).
//   full[s]   (leader's): both CTAs' TMA loads complete_tx here (.cta_group::2, peer bit cleared)
//   empty[s]  (each CTA): tcgen05.commit.cta_group::2 multicast releases stage s in both CTAs
//   tfull[b]  (each CTA): multicast commit after the last k-block: accumulator set b complete
//   tempty[b] (leader's): 4 epilogue warps of BOTH CTAs arrive (remote arrive from the peer) before set b is reused
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace pairprobe {

constexpr int BK = 64, kStages = 5, kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n\t.reg .pred p;\n\tWP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni DP;\n\tbra.uni WP;\n\tDP:\n\t}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
// wait with cluster-scope acquire: the arrivals come from both CTAs of the pair
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* b, uint32_t ph) {
  asm volatile("{\n\t.reg .pred p;\n\tWC:\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni DC;\n\tbra.uni WC;\n\tDC:\n\t}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* b, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(b)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2sm_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;   // the leader's barrier
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(m), "r"(mbar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

struct Geom {
  int M, N, K, BN, n_tiles, m_tiles256, relu;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_y, Geom g, const float* __restrict__ scale,
                 const float* __restrict__ shift) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint64_t* full_bar = (uint64_t*)raw;            // [kStages]
  uint64_t* empty_bar = full_bar + kStages;       // [kStages]
  uint64_t* tfull_bar = empty_bar + kStages;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;           // [2] (used in the leader)
  uint32_t* tmem_s = (uint32_t*)(tempty_bar + 2);
  float* ss_all = (float*)(raw + 256);            // [4 epilogue warps][scale 256 | shift 256]: private per warp, no races
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 256 + 8192 + 1023) & ~(uintptr_t)1023);
  const int a_bytes = 128 * BK * 2, b_bytes = (g.BN / 2) * BK * 2, stage_bytes = a_bytes + b_bytes;
  uint8_t* slabs = smem + (size_t)kStages * stage_bytes;   // [4 warps][2 buffers][32 rows x 64 B]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total = g.m_tiles256 * g.n_tiles, num_kb = (g.K + BK - 1) / BK;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)g.BN) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 8); }   // 4 warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_s)), "r"(2 * ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_s;

  if (warp == 0) {
    // ===== producer (both CTAs): own 128 rows of A, own half of the B rows =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        const int mt = tile / g.n_tiles, n0 = (tile % g.n_tiles) * g.BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_expect(&full_bar[stage], 2u * (uint32_t)stage_bytes);
          uint8_t* st = smem + (size_t)stage * stage_bytes;
          tma2sm_2d(&map_a, &full_bar[stage], st, kb * BK, mt * 256 + (int)rank * 128);
          tma2sm_2d(&map_b, &full_bar[stage], st + a_bytes, kb * BK, n0 + (int)rank * (g.BN / 2));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader only) =====
    if (rank == 0 && lane == 0) {
      constexpr uint64_t kDescHi = ((uint64_t)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
      const uint32_t idesc = (1u << 4) | ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t st_lo0 = (smem_u32(smem) & 0x3FFFF) >> 4, st_step = (uint32_t)stage_bytes >> 4, b_off = (uint32_t)a_bytes >> 4;
      uint32_t st_lo = st_lo0;
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < total; tile += npairs, ++it) {
        const int buf = it & 1;
        mbar_wait_cluster(&tempty_bar[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);   // both CTAs' epilogues drained this set
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + (uint32_t)buf * ncols;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t acc0 = kb ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ad = kDescHi | (uint64_t)(st_lo + 2 * k), bd = kDescHi | (uint64_t)(st_lo + b_off + 2 * k);
            const uint32_t acc = k ? 1u : acc0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
          }
          commit_pair(&empty_bar[stage]);
          st_lo += st_step;
          if (++stage == kStages) { stage = 0; phase ^= 1; st_lo = st_lo0; }
        }
        commit_pair(&tfull_bar[buf]);
      }
    }
  } else {
    // ===== epilogue (warps 2..5 of both CTAs): rows [256 mt + 128 rank + 32 q, +32) =====
    const int q = warp & 3;
    uint8_t* slab = slabs + (size_t)(warp - 2) * 4096;
    float* s_scale = ss_all + (size_t)(warp - 2) * 512;
    float* s_shift = s_scale + 256;
    int it = 0, pass = 0;
    for (int tile = pair; tile < total; tile += npairs, ++it) {
      const int buf = it & 1;
      const int mt = tile / g.n_tiles, n0 = (tile % g.n_tiles) * g.BN;
      for (int i = lane; i < g.BN; i += 32) {
        s_scale[i] = (scale && n0 + i < g.N) ? scale[n0 + i] : 1.0f;
        s_shift[i] = (shift && n0 + i < g.N) ? shift[n0 + i] : 0.0f;
      }
      __syncwarp();
      mbar_wait(&tfull_bar[buf], ((uint32_t)it >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * ncols;
      const int row0 = mt * 256 + (int)rank * 128 + q * 32;
      for (int cb = 0; cb < g.BN; cb += 32, ++pass) {
        uint8_t* sl = slab + (size_t)(pass & 1) * 2048;
        uint32_t v[32];
        tmem_ld32(taddr + cb, v);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          __align__(16) __half h[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float f = fmaf(__uint_as_float(v[c * 8 + k]), s_scale[cb + c * 8 + k], s_shift[cb + c * 8 + k]);
            if (g.relu) f = fmaxf(f, 0.0f);
            h[k] = __float2half_rn(f);
          }
          *reinterpret_cast<uint4*>(sl + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(h);
        }
        const bool last = cb + 32 >= g.BN;
        if (last) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (last) mbar_arrive_cluster(&tempty_bar[buf], 0);   // hand the set back to the leader's MMA warp
          if (n0 + cb < g.N) tma_store_2d(&map_y, sl, n0 + cb, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * ncols) : "memory");
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace pairprobe

// x [M, K] f16 (row stride x_ld), w [N, K] f16, y [M, N] f16 (row stride y_ld); BN in {64, 128, 256} dividing into N tiles
extern "C" int conv_pair_run(const void* x, int x_ld, const void* w, void* y, int y_ld, int M, int N, int K, int BN, int relu,
                             const float* scale, const float* shift, int reps, float* ms_out) {
  using namespace pairprobe;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) return -1;
  EncFn enc = (EncFn)f;
  CUtensorMap ma, mb, my;
  const cuuint32_t one[2] = {1, 1};
  {
    cuuint64_t d[2] = {(cuuint64_t)K, (cuuint64_t)M}, st[1] = {(cuuint64_t)x_ld * 2};
    cuuint32_t b[2] = {BK, 128};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)x, d, st, b, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return -2;
  }
  {
    cuuint64_t d[2] = {(cuuint64_t)K, (cuuint64_t)N}, st[1] = {(cuuint64_t)K * 2};
    cuuint32_t b[2] = {BK, (cuuint32_t)(BN / 2)};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)w, d, st, b, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return -3;
  }
  {
    cuuint64_t d[2] = {(cuuint64_t)N, (cuuint64_t)M}, st[1] = {(cuuint64_t)y_ld * 2};
    cuuint32_t b[2] = {32, 32};
    if (enc(&my, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, y, d, st, b, one, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return -4;
  }
  Geom g;
  g.M = M; g.N = N; g.K = K; g.BN = BN; g.n_tiles = (N + BN - 1) / BN; g.m_tiles256 = (M + 255) / 256; g.relu = relu;
  const size_t smem = 256 + 8192 + 1024 + (size_t)kStages * (128 * BK * 2 + (BN / 2) * BK * 2) + 4 * 4096;
  if (cudaFuncSetAttribute(conv_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -5;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = (sms / 2) * 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  conv_pair_kernel<<<grid, kThreads, smem>>>(ma, mb, my, g, scale, shift);
  if (cudaDeviceSynchronize() != cudaSuccess) return (int)cudaGetLastError();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) conv_pair_kernel<<<grid, kThreads, smem>>>(ma, mb, my, g, scale, shift);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = reps ? ms / reps : 0.0f;
  return (int)e;
}
