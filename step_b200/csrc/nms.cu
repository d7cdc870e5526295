// nms.cu -- bit-exact greedy NMS on sm_100a, fully device resident.
//
// Replaces _C.nms (external/maskrcnn_benchmark/csrc/nms.h:34-51).  The semantics every reference
// driver exercises are the CPU ones (test.py:192 calls nms on CPU tensors): legacy "+1" areas,
// suppress when IoU >= thr (cpu/nms_cpu.cpp:84); ge=0 gives the CUDA variant (cuda/nms.cu:84, '>').
// The IoU predicate is evaluated with explicitly rounded fp32 operations (__fadd_rn/__fmul_rn/
// __fdiv_rn: no FMA contraction, IEEE division) in the operand order of nms_cpu.cpp:74-83 so the
// keep set is bit-identical to the reference's.  Unlike cuda/nms.cu:118-147 there is no blocking
// D2H copy and no host scan: ordering, the 64x64 bitmask tiles, the greedy resolve and the
// compaction all run on the caller's stream.
#include "common.cuh"

namespace step {

__device__ __forceinline__ float box_area(float4 b) {
  // nms_cpu.cpp:46  (x2 - x1 + 1) * (y2 - y1 + 1)
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

__device__ __forceinline__ bool suppresses(float4 a, float area_a, float4 b, float area_b, float thr, int ge) {
  // nms_cpu.cpp:74-85
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(0.0f, __fadd_rn(__fsub_rn(xx2, xx1), 1.0f));
  float h = fmaxf(0.0f, __fadd_rn(__fsub_rn(yy2, yy1), 1.0f));
  float inter = __fmul_rn(w, h);
  float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return ge ? (ovr >= thr) : (ovr > thr);
}

// ---- 1. ordering: rank sort (score desc, index asc).  O(n^2) compares, embarrassingly parallel;
//         10k boxes = 1e8 compares ~ 10 us.  Deterministic and stable by construction, unlike the
//         reference's torch.sort (unstable for n > 16).
__global__ void __launch_bounds__(256) nms_rank_kernel(const float* __restrict__ boxes,
                                                       const float* __restrict__ scores, int n,
                                                       float4* __restrict__ sorted_boxes,
                                                       int* __restrict__ order) {
  __shared__ float tile[1024];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float si = i < n ? scores[i] : 0.0f;
  int rank = 0;
  for (int base = 0; base < n; base += 1024) {
    int cnt = min(1024, n - base);
    for (int t = threadIdx.x; t < cnt; t += blockDim.x) tile[t] = scores[base + t];
    __syncthreads();
    if (i < n) {
      // elements before i in index order win ties; split the loop so the compare is branch-free
      int lim = min(max(i - base, 0), cnt);
      for (int t = 0; t < lim; ++t) rank += (tile[t] >= si);
      for (int t = lim; t < cnt; ++t) rank += (tile[t] > si);
    }
    __syncthreads();
  }
  if (i < n) {
    sorted_boxes[rank] = reinterpret_cast<const float4*>(boxes)[i];
    order[rank] = i;
  }
}

// ---- 2. upper-triangular 64x64 bitmask tiles on the sorted boxes
constexpr int kTile = 64;
__global__ void __launch_bounds__(kTile) nms_mask_kernel(const float4* __restrict__ sb, int n, float thr,
                                                         int ge, int col_blocks,
                                                         unsigned long long* __restrict__ mask) {
  const int row_blk = blockIdx.y, col_blk = blockIdx.x;
  if (col_blk < row_blk) return;  // never read by the resolve
  __shared__ float4 cb[kTile];
  __shared__ float ca[kTile];
  const int col_size = min(n - col_blk * kTile, kTile);
  if (threadIdx.x < col_size) {
    float4 b = sb[col_blk * kTile + threadIdx.x];
    cb[threadIdx.x] = b;
    ca[threadIdx.x] = box_area(b);
  }
  __syncthreads();
  const int r = row_blk * kTile + threadIdx.x;
  if (r < n) {
    float4 a = sb[r];
    float aa = box_area(a);
    unsigned long long bits = 0;
    int start = (row_blk == col_blk) ? threadIdx.x + 1 : 0;
    for (int j = start; j < col_size; ++j)
      if (suppresses(a, aa, cb[j], ca[j], thr, ge)) bits |= 1ULL << j;
    mask[(size_t)r * col_blocks + col_blk] = bits;
  }
}

// ---- 3. greedy resolve on device: one CTA walks the 64-box blocks in score order.
__global__ void __launch_bounds__(1024) nms_resolve_kernel(const unsigned long long* __restrict__ mask,
                                                           const int* __restrict__ order, int n,
                                                           int col_blocks, uint8_t* __restrict__ keep_flag) {
  extern __shared__ unsigned long long remv[];  // [col_blocks]
  __shared__ unsigned long long diag[kTile];
  __shared__ unsigned long long kept_s;
  for (int j = threadIdx.x; j < col_blocks; j += blockDim.x) remv[j] = 0;
  __syncthreads();
  for (int blk = 0; blk < col_blocks; ++blk) {
    const int cnt = min(kTile, n - blk * kTile);
    if (threadIdx.x < cnt) diag[threadIdx.x] = mask[(size_t)(blk * kTile + threadIdx.x) * col_blocks + blk];
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long r = remv[blk], kept = 0;
      for (int b = 0; b < cnt; ++b)
        if (!((r >> b) & 1ULL)) { kept |= 1ULL << b; r |= diag[b]; }
      kept_s = kept;
    }
    __syncthreads();
    const unsigned long long kept = kept_s;
    if (threadIdx.x < cnt) keep_flag[order[blk * kTile + threadIdx.x]] = (uint8_t)((kept >> threadIdx.x) & 1ULL);
    for (int j = blk + 1 + threadIdx.x; j < col_blocks; j += blockDim.x) {
      unsigned long long acc = remv[j], k = kept;
      while (k) {
        int b = __ffsll((long long)k) - 1;
        k &= k - 1;
        acc |= mask[(size_t)(blk * kTile + b) * col_blocks + j];
      }
      remv[j] = acc;
    }
    __syncthreads();
  }
}

// ---- 4. compaction: kept original indices ascending (nms_cpu.cpp:88 nonzero(suppressed == 0))
__global__ void __launch_bounds__(1024) nms_compact_kernel(const uint8_t* __restrict__ keep_flag, int n,
                                                           int64_t* __restrict__ keep_out,
                                                           int* __restrict__ n_keep) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int start = 0; start < n; start += blockDim.x) {
    int i = start + threadIdx.x;
    int f = (i < n) ? keep_flag[i] : 0;
    unsigned bal = __ballot_sync(0xffffffffu, f);
    int within = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < wid; ++w) off += warp_tot[w];
    if (f) keep_out[off + within] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
      base_s += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_keep = base_s;
}

// ---- segmented small-problem kernel: one CTA per (clip, class) segment, everything in smem.
constexpr int kSegMax = 1024;
__global__ void __launch_bounds__(256) nms_segmented_kernel(const float* __restrict__ boxes,
                                                            const float* __restrict__ scores,
                                                            const int* __restrict__ seg_offsets, float thr,
                                                            int ge, float min_score,
                                                            uint8_t* __restrict__ keep_mask) {
  __shared__ float4 sb[kSegMax];
  __shared__ float sa[kSegMax];
  __shared__ float ss[kSegMax];
  __shared__ short sorig[kSegMax];
  __shared__ uint8_t sup[kSegMax];
  __shared__ int m_s;
  const int beg = seg_offsets[blockIdx.x], n = seg_offsets[blockIdx.x + 1] - beg;
  if (n <= 0) return;
  if (threadIdx.x == 0) m_s = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ss[i] = scores[beg + i];
  __syncthreads();
  // rank among rows passing the confidence threshold (test.py:183); others are dropped
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float si = ss[i];
    if (si >= min_score) {
      int rank = 0;
      for (int t = 0; t < n; ++t) {
        float st = ss[t];
        rank += (st >= min_score) && ((st > si) || (st == si && t < i));
      }
      float4 b = reinterpret_cast<const float4*>(boxes)[beg + i];
      sb[rank] = b;
      sa[rank] = box_area(b);
      sorig[rank] = (short)i;
      atomicAdd(&m_s, 1);
    } else {
      keep_mask[beg + i] = 0;
    }
  }
  __syncthreads();
  const int m = m_s;
  for (int i = threadIdx.x; i < m; i += blockDim.x) sup[i] = 0;
  __syncthreads();
  for (int i = 0; i < m; ++i) {
    if (!sup[i]) {  // uniform across the CTA (smem, synced)
      float4 a = sb[i];
      float aa = sa[i];
      for (int j = i + 1 + threadIdx.x; j < m; j += blockDim.x)
        if (!sup[j] && suppresses(a, aa, sb[j], sa[j], thr, ge)) sup[j] = 1;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < m; i += blockDim.x) keep_mask[beg + sorig[i]] = sup[i] ? 0 : 1;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace step

using namespace step;

extern "C" size_t step_nms_workspace_bytes(int n) {
  if (n <= 0) return 256;
  size_t cb = (size_t)ceil_div(n, kTile);
  return align256((size_t)n * sizeof(float4)) + align256((size_t)n * sizeof(int)) +
         align256((size_t)n * cb * sizeof(unsigned long long)) + align256((size_t)n) + 256;
}

extern "C" int step_nms_f32(const float* boxes, const float* scores, int n, float thr, int ge,
                            int64_t* keep_out, int* n_keep, void* workspace, size_t ws_bytes,
                            step_stream_t stream) {
  STEP_CHECK_ARG(n >= 0 && n_keep != nullptr, "step_nms_f32: bad n / n_keep");
  cudaStream_t s = cu(stream);
  if (n == 0) {  // nms_cpu.cpp:37-39: empty in, empty out
    cudaError_t e = cudaMemsetAsync(n_keep, 0, sizeof(int), s);
    return e == cudaSuccess ? 0 : fail((int)e, "step_nms_f32: memset: %s", cudaGetErrorString(e));
  }
  STEP_CHECK_ARG(boxes && scores && keep_out && workspace, "step_nms_f32: null pointer");
  STEP_CHECK_ARG(((uintptr_t)boxes & 15) == 0, "step_nms_f32: boxes must be 16-byte aligned");
  const int cb = ceil_div(n, kTile);
  STEP_CHECK_ARG((size_t)cb * 8 <= 200 * 1024, "step_nms_f32: n too large (max %d)", 200 * 1024 / 8 * kTile);
  if (ws_bytes < step_nms_workspace_bytes(n))
    return fail(STEP_E_WORKSPACE, "step_nms_f32: workspace %zu < %zu", ws_bytes, step_nms_workspace_bytes(n));
  char* w = (char*)workspace;
  float4* sorted = (float4*)w;  w += align256((size_t)n * sizeof(float4));
  int* order = (int*)w;         w += align256((size_t)n * sizeof(int));
  unsigned long long* mask = (unsigned long long*)w;  w += align256((size_t)n * cb * sizeof(unsigned long long));
  uint8_t* flag = (uint8_t*)w;

  nms_rank_kernel<<<ceil_div(n, 256), 256, 0, s>>>(boxes, scores, n, sorted, order);
  STEP_LAUNCH_CHECK("nms_rank_kernel");
  nms_mask_kernel<<<dim3(cb, cb), kTile, 0, s>>>(sorted, n, thr, ge, cb, mask);
  STEP_LAUNCH_CHECK("nms_mask_kernel");
  size_t smem = (size_t)cb * sizeof(unsigned long long);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nms_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "step_nms_f32: smem attr: %s", cudaGetErrorString(e));
  }
  nms_resolve_kernel<<<1, 1024, smem, s>>>(mask, order, n, cb, flag);
  STEP_LAUNCH_CHECK("nms_resolve_kernel");
  nms_compact_kernel<<<1, 1024, 0, s>>>(flag, n, keep_out, n_keep);
  STEP_LAUNCH_CHECK("nms_compact_kernel");
  return 0;
}

extern "C" int step_nms_segmented_f32(const float* boxes, const float* scores, const int* seg_offsets,
                                      int n_seg, float thr, int ge, float min_score, uint8_t* keep_mask,
                                      step_stream_t stream) {
  STEP_CHECK_ARG(n_seg >= 0, "step_nms_segmented_f32: n_seg < 0");
  if (n_seg == 0) return 0;
  STEP_CHECK_ARG(boxes && scores && seg_offsets && keep_mask, "step_nms_segmented_f32: null pointer");
  STEP_CHECK_ARG(((uintptr_t)boxes & 15) == 0, "step_nms_segmented_f32: boxes must be 16-byte aligned");
  nms_segmented_kernel<<<n_seg, 256, 0, cu(stream)>>>(boxes, scores, seg_offsets, thr, ge, min_score, keep_mask);
  STEP_LAUNCH_CHECK("nms_segmented_kernel");
  return 0;
}
