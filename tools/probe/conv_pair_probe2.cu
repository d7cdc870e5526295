// Round-2 experiment harness (NOT part of libstep_b200.so): epilogue variants of the CTA-pair (cta_group::2) persistent
// GEMM of conv_pair_probe.cu, to find out what bounds the store-heavy 1x1 layers (256 -> 1024 + residual + ReLU: 54 us in
// round 1 against an HBM floor of ~25 us).
//
//   D[M, N] = relu(A[M, K] * B[N, K]^T * scale + shift (+ R[M, N]))        fp16 in, fp32 accumulate, fp16 out
//
// Per CTA: warp 0 TMA producer, warp 1 MMA issuer (leader CTA), warps 2..9 epilogue = (TMEM lane quarter) x (column half).
//   EPI 0: TMEM -> regs -> smem slab (2 buffers / warp) -> bulk tensor store, 32-column boxes      (round-1 scheme)
//   EPI 1: same with 4 slab buffers per warp (more bulk stores in flight)
//   EPI 2: TMEM -> regs -> st.global.v8.b32 (32 B per lane, full sectors), no shared memory, no TMA store
//   EPI 3: TMEM -> regs -> 4 x st.global.v4.b32
// RES: the residual row segment is read with ld.global.nc.v8 straight into registers, prefetched one pass ahead.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace pair2 {

constexpr int BK = 64, kMaxStages = 6, kThreads = 320;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n\t.reg .pred p;\n\tWP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni DP;\n\tbra.uni WP;\n\tDP:\n\t}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* b, uint32_t ph) {
  asm volatile("{\n\t.reg .pred p;\n\tWC:\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t@p bra.uni DC;\n\tbra.uni WC;\n\tDC:\n\t}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
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
__device__ __forceinline__ void stg256(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ldg256(const void* p, uint32_t* v) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}

struct Geom {
  int M, N, K, BN, n_tiles, m_tiles256, relu, stages, y_ld, res_ld, knock;
};

template <int EPI, bool RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_pair2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_y, Geom g, const float* __restrict__ scale,
                  const float* __restrict__ shift, const __half* __restrict__ res, __half* __restrict__ y) {
  constexpr int kSlabBufs = (EPI == 1 || EPI == 6) ? 4 : 2;
  constexpr bool kTmaSt = EPI <= 1 || EPI >= 5;
  constexpr bool kPipe = EPI >= 4;
  extern __shared__ __align__(1024) uint8_t raw[];
  uint64_t* full_bar = (uint64_t*)raw;               // [kMaxStages]
  uint64_t* empty_bar = full_bar + kMaxStages;       // [kMaxStages]
  uint64_t* tfull_bar = empty_bar + kMaxStages;      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;              // [2] (used in the leader)
  uint32_t* tmem_s = (uint32_t*)(tempty_bar + 2);
  float* ss_all = (float*)(raw + 256);               // [8 epilogue warps][scale 128 | shift 128]
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 256 + 8192 + 1023) & ~(uintptr_t)1023);
  const int a_bytes = 128 * BK * 2, b_bytes = (g.BN / 2) * BK * 2, stage_bytes = a_bytes + b_bytes;
  uint8_t* slabs = smem + (size_t)g.stages * stage_bytes;   // [8 warps][kSlabBufs][32 rows x 64 B]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int total = g.m_tiles256 * g.n_tiles, num_kb = (g.K + BK - 1) / BK;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)g.BN) ncols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 16); }   // 8 warps x 2 CTAs
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
    if (lane == 0 && !(g.knock & 4)) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = pair; tile < total; tile += npairs) {
        const int mt = tile / g.n_tiles, n0 = (tile % g.n_tiles) * g.BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_expect(&full_bar[stage], 2u * (uint32_t)stage_bytes);
          uint8_t* st = smem + (size_t)stage * stage_bytes;
          tma2sm_2d(&map_a, &full_bar[stage], st, kb * BK, mt * 256 + (int)rank * 128);
          tma2sm_2d(&map_b, &full_bar[stage], st + a_bytes, kb * BK, n0 + (int)rank * (g.BN / 2));
          if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0 && !(g.knock & 4)) {
      constexpr uint64_t kDescHi = ((uint64_t)(1024 >> 4) << 32) | (1ULL << 46) | (2ULL << 61);
      const uint32_t idesc = (1u << 4) | ((uint32_t)(g.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t st_lo0 = (smem_u32(smem) & 0x3FFFF) >> 4, st_step = (uint32_t)stage_bytes >> 4, b_off = (uint32_t)a_bytes >> 4;
      uint32_t st_lo = st_lo0;
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < total; tile += npairs, ++it) {
        const int buf = it & 1;
        mbar_wait_cluster(&tempty_bar[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
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
          if (++stage == g.stages) { stage = 0; phase ^= 1; st_lo = st_lo0; }
        }
        commit_pair(&tfull_bar[buf]);
      }
    }
  } else {
    // ===== epilogue: warp = (lane quarter q, column half sel); rows [256 mt + 128 rank + 32 q, +32), columns [sel BN/2, +BN/2)
    const int q = warp & 3, sel = (warp - 2) >> 2;
    const int ncol = g.BN >> 1, col0 = sel * ncol;
    uint8_t* slab = slabs + (size_t)(warp - 2) * (kSlabBufs * 2048);
    float* s_scale = ss_all + (size_t)(warp - 2) * 256;
    float* s_shift = s_scale + 128;
    int it = 0, pass = 0;
    for (int tile = pair; tile < total; tile += npairs, ++it) {
      const int buf = it & 1;
      const int mt = tile / g.n_tiles, n0 = (tile % g.n_tiles) * g.BN;
      const int nbase = n0 + col0;
      for (int i = lane; i < ncol; i += 32) {
        s_scale[i] = (scale && nbase + i < g.N) ? scale[nbase + i] : 1.0f;
        s_shift[i] = (shift && nbase + i < g.N) ? shift[nbase + i] : 0.0f;
      }
      __syncwarp();
      const int row0 = mt * 256 + (int)rank * 128 + q * 32;
      const int row = row0 + lane;
      const bool row_ok = row < g.M;
      uint32_t rvA[16], rvB[16];
      const __half* rrow = RES ? res + (size_t)(row_ok ? row : 0) * g.res_ld + nbase : nullptr;
      if (RES && row_ok) { ldg256(rrow, rvA); ldg256(rrow + 16, rvA + 8); }
      if (!(g.knock & 4)) mbar_wait(&tfull_bar[buf], ((uint32_t)it >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (g.knock & 2) {   // no epilogue work: hand the accumulator straight back
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tempty_bar[buf], 0);
        continue;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * ncols + (uint32_t)col0;
      // process(): registers of one 32-column pass -> scale/shift (+residual) (+relu) -> fp16 -> global memory
      auto process = [&](int cb, const uint32_t* v, const uint32_t* rcur) {
        uint8_t* sl = slab + (size_t)(pass % kSlabBufs) * 2048;
        uint32_t o[16];
        const bool ss = scale != nullptr || shift != nullptr;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float f0 = __uint_as_float(v[2 * c]), f1 = __uint_as_float(v[2 * c + 1]);
          if (ss) {
            f0 = fmaf(f0, s_scale[cb + 2 * c], s_shift[cb + 2 * c]);
            f1 = fmaf(f1, s_scale[cb + 2 * c + 1], s_shift[cb + 2 * c + 1]);
          }
          if (RES) {
            const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rcur[c]));
            f0 += rf.x; f1 += rf.y;
          }
          if (g.relu) { f0 = fmaxf(f0, 0.0f); f1 = fmaxf(f1, 0.0f); }
          const __half2 h = __floats2half2_rn(f0, f1);
          o[c] = *reinterpret_cast<const uint32_t*>(&h);
        }
        if (kTmaSt) {
          if (lane == 0) {
            if (kSlabBufs == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
          }
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(sl + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (nbase + cb < g.N && !(g.knock & 1)) tma_store_2d(&map_y, sl, nbase + cb, row0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else if (row_ok && nbase + cb < g.N && !(g.knock & 1)) {
          __half* dst = y + (size_t)row * g.y_ld + nbase + cb;
          if (EPI != 3) { stg256(dst, o); stg256(dst + 16, o + 8); }
          else {
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + 8 * c) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
          }
        }
        ++pass;
      };
      auto release = [&]() {   // every tcgen05.ld of this tile has landed in registers: hand the accumulator back
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&tempty_bar[buf], 0);
      };
      if (kPipe) {
        // software pipeline: the TMEM load of pass p+1 is in flight while pass p is converted and stored
        uint32_t vA[32], vB[32];
        tmem_ld32(taddr, vA);
        for (int cb = 0; cb < ncol; cb += 64) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (cb + 32 < ncol) tmem_ld32(taddr + cb + 32, vB); else release();
          if (RES && row_ok && cb + 32 < ncol) { ldg256(rrow + cb + 32, rvB); ldg256(rrow + cb + 48, rvB + 8); }
          process(cb, vA, rvA);
          if (cb + 32 < ncol) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (cb + 64 < ncol) tmem_ld32(taddr + cb + 64, vA); else release();
            if (RES && row_ok && cb + 64 < ncol) { ldg256(rrow + cb + 64, rvA); ldg256(rrow + cb + 80, rvA + 8); }
            process(cb + 32, vB, rvB);
          }
        }
      } else {
        for (int cb = 0; cb < ncol; cb += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + cb, v);
          uint32_t* rcur = ((cb >> 5) & 1) ? rvB : rvA;
          uint32_t* rnext = ((cb >> 5) & 1) ? rvA : rvB;
          if (RES && row_ok && cb + 32 < ncol) { ldg256(rrow + cb + 32, rnext); ldg256(rrow + cb + 48, rnext + 8); }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (cb + 32 >= ncol) release();
          process(cb, v, rcur);
        }
      }
    }
    if (kTmaSt && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

template <int EPI, bool RES>
static int run(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& my, Geom g, const float* scale, const float* shift,
               const __half* res, __half* y, int reps, float* ms_out) {
  const int slab_bufs = (EPI == 1 || EPI == 6) ? 4 : 2;
  size_t fixed = 256 + 8192 + 1024 + 8 * (size_t)slab_bufs * 2048;
  const size_t stage_bytes = 128 * BK * 2 + (g.BN / 2) * BK * 2;
  int st = (int)((227 * 1024 - fixed) / stage_bytes);
  if (st > kMaxStages) st = kMaxStages;
  if (g.stages > 0 && g.stages < st) st = g.stages;
  g.stages = st;
  const size_t smem = fixed + (size_t)st * stage_bytes;
  if (cudaFuncSetAttribute(conv_pair2_kernel<EPI, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -5;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int total = g.m_tiles256 * g.n_tiles;
  int pairs = sms / 2;
  if (pairs > total) pairs = total;
  const int grid = pairs * 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  conv_pair2_kernel<EPI, RES><<<grid, kThreads, smem>>>(ma, mb, my, g, scale, shift, res, y);
  if (cudaDeviceSynchronize() != cudaSuccess) return (int)cudaGetLastError();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) conv_pair2_kernel<EPI, RES><<<grid, kThreads, smem>>>(ma, mb, my, g, scale, shift, res, y);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = reps ? ms / reps : 0.0f;
  return (int)e;
}

}  // namespace pair2

// x [M, K] f16 (row stride x_ld), w [N, K] f16, y [M, N] f16 (row stride y_ld), res [M, N] f16 (row stride res_ld) or null
extern "C" int conv_pair2_run(const void* x, int x_ld, const void* w, void* y, int y_ld, const void* res, int res_ld, int M, int N,
                              int K, int BN, int relu, const float* scale, const float* shift, int epi, int reps, float* ms_out, int knock,
                              int stages_cap) {
  using namespace pair2;
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
  g.stages = stages_cap; g.y_ld = y_ld; g.res_ld = res_ld; g.knock = knock;
  const __half* r = (const __half*)res;
  __half* yy = (__half*)y;
#define RUN(E) (r ? run<E, true>(ma, mb, my, g, scale, shift, r, yy, reps, ms_out) : run<E, false>(ma, mb, my, g, scale, shift, r, yy, reps, ms_out))
  switch (epi) {
    case 0: return RUN(0);
    case 1: return RUN(1);
    case 2: return RUN(2);
    case 3: return RUN(3);
    case 4: return RUN(4);
    case 5: return RUN(5);
    case 6: return RUN(6);
  }
  return -9;
}
