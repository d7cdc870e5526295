// train.cu -- the first pieces of the training step (train.py:286-348 of the reference):
//   * head_losses_kernel       TwoBranchNet's three losses (models/two_branch.py:276-333) and, in the same pass, the
//                              gradient of the training objective  mean(loss_cls) + w_loc * loss_loc + w_nb * loss_nb
//                              (train.py:323-347) with respect to the head outputs;
//   * roi_align_bwd_nhwc       channels-last ROIAlign backward WITHOUT float atomics: every feature pixel gathers its
//                              contributions in a fixed order (ROI index, bin, sample), so the result is bit-for-bit
//                              repeatable -- the reference's RoIAlignBackwardFeature (cuda/ROIAlign_cuda.cu:201-278) scatters
//                              with atomicAdd and is not;
//   * linear_bwd_*             backward of the small-N linears of the head (global_cls, local_reg, neighbor_reg*):
//                              dx = dy W, dW = dy^T x, db = sum dy, fixed summation order;
//   * conv1x1_wgrad_*          weight gradient of a 1x1(x1) convolution, dW[Cout, Cin] = dz^T x with the reduction over
//                              the pixels on the tensor cores (fp16 operands, fp32 accumulate), split over pixel chunks with
//                              a fixed-order second pass.
// All fp32 arithmetic of the losses is explicitly rounded in the reference's operand order; exp / log are the only
// operations that are not bit-exact (tolerance stated in tests/test_gpu_train.py).
#include <math.h>

#include <mma.h>

#include "common.cuh"

namespace step {

// ---- losses -------------------------------------------------------------------------------------------------------
struct CS4 { float x, y, w, h; };
__device__ __forceinline__ CS4 centre_size(const float* b) {
  // tube_utils.py:127-141
  CS4 c;
  c.w = __fadd_rn(__fsub_rn(b[2], b[0]), 1.0f);
  c.h = __fadd_rn(__fsub_rn(b[3], b[1]), 1.0f);
  c.x = __fadd_rn(b[0], __fmul_rn(0.5f, c.w));
  c.y = __fadd_rn(b[1], __fmul_rn(0.5f, c.h));
  return c;
}
// tube_utils.py:143-163 encode_coef(gt, anchor) -> (dx, dy, dw, dh)
__device__ __forceinline__ void encode4(const float* gt, const float* anchor, float* out) {
  const CS4 g = centre_size(gt), a = centre_size(anchor);
  out[0] = __fdiv_rn(__fsub_rn(g.x, a.x), a.w);
  out[1] = __fdiv_rn(__fsub_rn(g.y, a.y), a.h);
  out[2] = logf(__fdiv_rn(g.w, a.w));
  out[3] = logf(__fdiv_rn(g.h, a.h));
}
// F.smooth_l1_loss(beta = 1): 0.5 d^2 if |d| < 1 else |d| - 0.5; derivative d | sign(d)
__device__ __forceinline__ float smooth_l1(float d, float* grad) {
  const float ad = fabsf(d);
  if (ad < 1.0f) { *grad = d; return __fmul_rn(__fmul_rn(0.5f, d), d); }
  *grad = d > 0.0f ? 1.0f : -1.0f;
  return __fsub_rn(ad, 0.5f);
}

struct LossGeom {
  int N, cls, T_len, Tc;     // tubes, classes, frames of local_loc, frames of first/last_loc (= T)
  int centre, first_idx, last_idx, half_T;   // chunk_idx[chunks/2], chunk_idx[0], chunk_idx[-1] (two_branch.py:226-228)
  int s0, e0;                // first / last chunk start inside local_loc (two_branch.py:265-266)
  int tgt_ld;                // 6 + cls
  float w_loc, w_nb;         // lambda_reg, lambda_neighbor (train.py:335-336)
};

// One CTA.  Sums are taken in tube order by one thread after a block-wide staging pass: N is a few hundred at most and
// this keeps the reductions bit-for-bit repeatable.
__global__ void __launch_bounds__(256) head_losses_kernel(LossGeom g, const float* __restrict__ logits,
                                                          const float* __restrict__ local_loc, const float* __restrict__ first_loc,
                                                          const float* __restrict__ last_loc, const float* __restrict__ tubes,
                                                          const float* __restrict__ targets, float* __restrict__ loss_cls,
                                                          float* __restrict__ loss_loc, float* __restrict__ loss_nb,
                                                          int* __restrict__ flags, float* __restrict__ dlogits,
                                                          float* __restrict__ dloc, float* __restrict__ dfirst,
                                                          float* __restrict__ dlast, float* __restrict__ scratch) {
  __shared__ float s_sum[3];   // sum of cls mask, loc mask x4, neighbour mask x4
  __shared__ float s_loss[2];
  const int N = g.N;
  // targets[n][j] with j = 0 first, 1 centre, 2 last (two_branch.py:283-285: [:, 0], [:, 1], [:, -1])
  auto tgt = [&](int n, int j) { return targets + ((size_t)n * 3 + j) * g.tgt_ld; };
  if (threadIdx.x == 0) {
    float mc = 0.0f, ml = 0.0f, mn = 0.0f;
    for (int n = 0; n < N; ++n) mc = __fadd_rn(mc, tgt(n, 1)[4]);
    for (int n = 0; n < N; ++n) for (int k = 0; k < 4; ++k) ml = __fadd_rn(ml, tgt(n, 1)[5]);
    for (int n = 0; n < N; ++n) for (int k = 0; k < 4; ++k) mn = __fadd_rn(mn, tgt(n, 0)[5]);
    for (int n = 0; n < N; ++n) for (int k = 0; k < 4; ++k) mn = __fadd_rn(mn, tgt(n, 2)[5]);
    s_sum[0] = mc; s_sum[1] = ml; s_sum[2] = mn;
    flags[0] = mc != 0.0f; flags[1] = ml != 0.0f; flags[2] = mn != 0.0f;
  }
  __syncthreads();
  const bool has_cls = s_sum[0] != 0.0f, has_loc = s_sum[1] != 0.0f, has_nb = s_sum[2] != 0.0f;
  // ---- classification: BCE with logits on the centre chunk, background samples masked (two_branch.py:291-297)
  const float inv_ncls = 1.0f / (float)((long long)N * g.cls);
  for (int i = threadIdx.x; i < N * g.cls; i += blockDim.x) {
    const int n = i / g.cls, c = i - n * g.cls;
    const float x = logits[i];
    float l = 0.0f, gx = 0.0f;
    if (has_cls) {
      const float t = __fmul_rn(tgt(n, 1)[6 + c], tgt(n, 1)[4]);
      // ATen: (1 - t) * x - log_sigmoid(x),  log_sigmoid(x) = min(x, 0) - log1p(exp(-|x|))
      const float ls = __fsub_rn(fminf(x, 0.0f), log1pf(expf(-fabsf(x))));
      l = __fsub_rn(__fmul_rn(__fsub_rn(1.0f, t), x), ls);
      const float sg = 1.0f / (1.0f + expf(-x));
      gx = __fmul_rn(__fsub_rn(sg, t), inv_ncls);              // d mean(loss_cls) / d logit
    }
    loss_cls[i] = l;
    if (dlogits) dlogits[i] = gx;
  }
  // ---- regression: per-tube smooth-L1 terms staged in `scratch` [N][3][4] (loss) and gradients written in place
  if (dloc) for (int i = threadIdx.x; i < N * g.T_len * 4; i += blockDim.x) dloc[i] = 0.0f;
  if (dfirst) for (int i = threadIdx.x; i < N * g.Tc * 4; i += blockDim.x) { dfirst[i] = 0.0f; dlast[i] = 0.0f; }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float* tb = tubes + (size_t)n * g.T_len * 5;
    float enc[4], gr;
    // centre (two_branch.py:301-311)
    encode4(tgt(n, 1), tb + (size_t)g.centre * 5 + 1, enc);
    const float mloc = tgt(n, 1)[5];
    for (int k = 0; k < 4; ++k) {
      const float d = __fsub_rn(local_loc[((size_t)n * g.T_len + g.centre) * 4 + k], enc[k]);
      const float l = smooth_l1(d, &gr);
      scratch[((size_t)n * 3 + 0) * 4 + k] = __fmul_rn(l, mloc);
      if (dloc && has_loc) dloc[((size_t)n * g.T_len + g.centre) * 4 + k] = __fdiv_rn(__fmul_rn(__fmul_rn(gr, mloc), g.w_loc), s_sum[1]);
    }
    // neighbours (two_branch.py:315-333): first then last
    for (int j = 0; j < 2; ++j) {
      const int tj = j == 0 ? 0 : 2, idx = j == 0 ? g.first_idx : g.last_idx;
      const float* pred = (j == 0 ? first_loc : last_loc) + ((size_t)n * g.Tc + g.half_T) * 4;
      float* dpred = (j == 0 ? dfirst : dlast);
      encode4(tgt(n, tj), tb + (size_t)idx * 5 + 1, enc);
      const float m = tgt(n, tj)[5];
      for (int k = 0; k < 4; ++k) {
        const float d = __fsub_rn(pred[k], enc[k]);
        const float l = smooth_l1(d, &gr);
        scratch[((size_t)n * 3 + 1 + j) * 4 + k] = __fmul_rn(l, m);
        if (dpred && has_nb) dpred[((size_t)n * g.Tc + g.half_T) * 4 + k] = __fdiv_rn(__fmul_rn(__fmul_rn(gr, m), g.w_nb), s_sum[2]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sl = 0.0f, sn = 0.0f;
    for (int n = 0; n < N; ++n) for (int k = 0; k < 4; ++k) sl = __fadd_rn(sl, scratch[((size_t)n * 3 + 0) * 4 + k]);
    for (int n = 0; n < N; ++n) for (int k = 0; k < 4; ++k) sn = __fadd_rn(sn, scratch[((size_t)n * 3 + 1) * 4 + k]);   // cat([first, last])
    for (int n = 0; n < N; ++n) for (int k = 0; k < 4; ++k) sn = __fadd_rn(sn, scratch[((size_t)n * 3 + 2) * 4 + k]);
    s_loss[0] = has_loc ? __fdiv_rn(sl, s_sum[1]) : 0.0f;
    s_loss[1] = has_nb ? __fdiv_rn(sn, s_sum[2]) : 0.0f;
    loss_loc[0] = s_loss[0];
    loss_nb[0] = s_loss[1];
  }
  __syncthreads();
  // first_loc / last_loc are slices of local_loc plus the neighbour regressors (two_branch.py:265-270): their gradient
  // also flows into local_loc at the slice positions
  if (dloc && dfirst) {
    for (int i = threadIdx.x; i < N * g.Tc * 4; i += blockDim.x) {
      const int n = i / (g.Tc * 4), r = i - n * g.Tc * 4;
      const float a = dfirst[i], b = dlast[i];
      // element (n, r) belongs to this thread alone: the [s0, s0 + Tc) and [e0, e0 + Tc) ranges are either identical
      // (one chunk) or disjoint, so the two updates never race and their order is fixed
      if (a != 0.0f) dloc[((size_t)n * g.T_len + g.s0) * 4 + r] += a;
      if (b != 0.0f) dloc[((size_t)n * g.T_len + g.e0) * 4 + r] += b;
    }
  }
}

// ---- ROIAlign backward, channels-last, deterministic ----------------------------------------------------------------
// grad_in[k][h][w][c] = sum over ROIs r on frame k (ascending r), bins (ph, pw) and samples (iy, ix) in loop order of
//   w_tap(r, bin, sample, this pixel) * grad_out[r][bin][c] / count(r)        (ROIAlign_cuda.cu:201-278)
// One CTA per (frame, pixel row); the per-ROI sample table (bin, 4 pixel ids, 4 weights) is built in shared memory by the
// CTA and scanned by every (pixel, channel vector) thread.  No atomics: a thread owns its output element.
struct RoiBwdSample { int pix[4]; float w[4]; int bin; };
constexpr int kBwdMaxSamples = 49 * 16;   // 7x7 bins x sampling grid <= 4x4; larger grids are processed in slices

__device__ __forceinline__ void bilinear_taps(int H, int W, float y, float x, int* pix, float* wt) {
  // ROIAlign_cuda.cu:149-199 bilinear_interpolate_gradient
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) { pix[0] = pix[1] = pix[2] = pix[3] = -1; wt[0] = wt[1] = wt[2] = wt[3] = 0.0f; return; }
  if (y <= 0.0f) y = 0.0f;
  if (x <= 0.0f) x = 0.0f;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= H - 1) { y_high = y_low = H - 1; y = (float)y_low; } else { y_high = y_low + 1; }
  if (x_low >= W - 1) { x_high = x_low = W - 1; x = (float)x_low; } else { x_high = x_low + 1; }
  const float ly = __fsub_rn(y, (float)y_low), lx = __fsub_rn(x, (float)x_low);
  const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
  wt[0] = __fmul_rn(hy, hx); wt[1] = __fmul_rn(hy, lx); wt[2] = __fmul_rn(ly, hx); wt[3] = __fmul_rn(ly, lx);
  pix[0] = y_low * W + x_low; pix[1] = y_low * W + x_high; pix[2] = y_high * W + x_low; pix[3] = y_high * W + x_high;
}

template <typename T>
__global__ void __launch_bounds__(256) roi_align_bwd_nhwc_kernel(const T* __restrict__ grad_out, int out_ld,
                                                                 const float* __restrict__ rois, int R, float scale, int H,
                                                                 int W, int C, int ph, int pw, int sampling_ratio,
                                                                 float* __restrict__ grad_in, int in_ld) {
  __shared__ RoiBwdSample tab[kBwdMaxSamples];
  __shared__ int n_s;
  const int frame = blockIdx.x;
  constexpr int VN = 4;
  const int nvec = C / VN, npix = H * W;
  // this CTA owns grad_in[frame]: zero it, then accumulate ROI by ROI (ascending index: fixed order)
  for (int i = threadIdx.x; i < npix * nvec; i += blockDim.x) {
    const int p = i / nvec, cv = i - p * nvec;
    *reinterpret_cast<float4*>(grad_in + ((size_t)frame * npix + p) * in_ld + cv * VN) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int r = 0; r < R; ++r) {
    const float* roi = rois + 5 * (size_t)r;
    if ((int)roi[0] != frame) continue;                       // uniform across the CTA
    // ROIAlign_cuda.cu:214-233
    const float sw = __fmul_rn(roi[1], scale), sh = __fmul_rn(roi[2], scale), ew = __fmul_rn(roi[3], scale), eh = __fmul_rn(roi[4], scale);
    const float roi_w = fmaxf(__fsub_rn(ew, sw), 1.0f), roi_h = fmaxf(__fsub_rn(eh, sh), 1.0f);
    const float bin_h = __fdiv_rn(roi_h, (float)ph), bin_w = __fdiv_rn(roi_w, (float)pw);
    const int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(roi_h, (float)ph));
    const int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(roi_w, (float)pw));
    const float count = (float)(gh * gw);
    const int per_bin = gh * gw, total = ph * pw * per_bin;
    for (int base = 0; base < total; base += kBwdMaxSamples) {
      const int cnt = min(kBwdMaxSamples, total - base);
      __syncthreads();                                        // the previous table has been consumed
      for (int s = threadIdx.x; s < cnt; s += blockDim.x) {
        const int gidx = base + s;
        const int bin = gidx / per_bin, rem = gidx - bin * per_bin;
        const int iy = rem / gw, ix = rem - iy * gw;
        const int p = bin / pw, q = bin - p * pw;
        // ROIAlign_cuda.cu:236-240: start + ph * bin + (iy + .5) * bin / grid
        const float y = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)p, bin_h)), __fdiv_rn(__fmul_rn(__fadd_rn((float)iy, 0.5f), bin_h), (float)gh));
        const float x = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)q, bin_w)), __fdiv_rn(__fmul_rn(__fadd_rn((float)ix, 0.5f), bin_w), (float)gw));
        RoiBwdSample e;
        bilinear_taps(H, W, y, x, e.pix, e.w);
        e.bin = bin;
        tab[s] = e;
      }
      if (threadIdx.x == 0) n_s = cnt;
      __syncthreads();
      const T* go = grad_out + (size_t)r * ph * pw * out_ld;
      for (int i = threadIdx.x; i < npix * nvec; i += blockDim.x) {
        const int p = i / nvec, cv = i - p * nvec;
        float acc[VN] = {0.f, 0.f, 0.f, 0.f};
        bool hit = false;
        for (int s = 0; s < cnt; ++s) {
          const RoiBwdSample& e = tab[s];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (e.pix[k] == p) {
              float gv[VN];
              if constexpr (sizeof(T) == 4) {
                const float4 v = *reinterpret_cast<const float4*>(go + (size_t)e.bin * out_ld + cv * VN);
                gv[0] = v.x; gv[1] = v.y; gv[2] = v.z; gv[3] = v.w;
              } else {
                const uint2 raw = *reinterpret_cast<const uint2*>(go + (size_t)e.bin * out_ld + cv * VN);
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
                gv[0] = a.x; gv[1] = a.y; gv[2] = b.x; gv[3] = b.y;
              }
              // ROIAlign_cuda.cu:263-271: g = top_diff * w / count
#pragma unroll
              for (int c = 0; c < VN; ++c) acc[c] = __fadd_rn(acc[c], __fdiv_rn(__fmul_rn(gv[c], e.w[k]), count));
              hit = true;
            }
          }
        }
        if (hit) {
          float4* dst = reinterpret_cast<float4*>(grad_in + ((size_t)frame * npix + p) * in_ld + cv * VN);
          float4 cur = *dst;
          cur.x = __fadd_rn(cur.x, acc[0]); cur.y = __fadd_rn(cur.y, acc[1]); cur.z = __fadd_rn(cur.z, acc[2]); cur.w = __fadd_rn(cur.w, acc[3]);
          *dst = cur;
        }
      }
    }
  }
}

// ---- small-N linear backward -----------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) linear_bwd_dw_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ dy,
                                                            int M, int K, int Nn, float* __restrict__ dw,
                                                            float* __restrict__ db) {
  // dw[n][k] = sum_m dy[m][n] * x[m][k]   (m ascending: fixed order);  db[n] = sum_m dy[m][n]
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)Nn * K) {
    const int n = (int)(i / K), k = (int)(i - (long long)n * K);
    float acc = 0.0f;
    for (int m = 0; m < M; ++m) acc = fmaf(dy[(size_t)m * Nn + n], to_f32<T>(x[(size_t)m * x_ld + k]), acc);
    dw[i] = acc;
  }
  if (db && i < Nn) {
    float acc = 0.0f;
    for (int m = 0; m < M; ++m) acc += dy[(size_t)m * Nn + (int)i];
    db[i] = acc;
  }
}

__global__ void __launch_bounds__(256) linear_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ w, int M,
                                                            int K, int Nn, float* __restrict__ dx, int accumulate) {
  // dx[m][k] (+)= sum_n dy[m][n] * w[n][k]
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * K) return;
  const int m = (int)(i / K), k = (int)(i - (long long)m * K);
  float acc = accumulate ? dx[i] : 0.0f;
  for (int n = 0; n < Nn; ++n) acc = fmaf(dy[(size_t)m * Nn + n], w[(size_t)n * K + k], acc);
  dx[i] = acc;
}


// ---- elementwise pieces of the conv backward ---------------------------------------------------------------------------
// y = relu(scale * conv + shift (+ residual))  (Unit3Dpy, i3dpt.py:103-111; Bottleneck, two_branch.py:60-111):
//   dz = dy * [y > 0] * scale      gradient w.r.t. the raw convolution output (operand of dgrad / wgrad)
//   dres += dy * [y > 0]           gradient flowing into the residual input, accumulated in place
// dy / y / dres are channel slices of wider fp16 buffers (ld, coff); dz is dense [M, C] at dz_ld / dz_coff.
__global__ void __launch_bounds__(256) act_bwd_kernel(const __half* __restrict__ dy, int dy_ld, const __half* __restrict__ y,
                                                      int y_ld, const float* __restrict__ scale, int relu, long long M, int C,
                                                      __half* __restrict__ dz, int dz_ld, __half* __restrict__ dres,
                                                      int dres_ld) {
  const int cv = C >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * cv) return;
  const long long m = i / cv;
  const int c = (int)(i - m * cv) * 8;
  float g[8], v[8];
  load16(dy + (size_t)m * dy_ld + c, g);
  if (relu) {
    load16(y + (size_t)m * y_ld + c, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = v[k] > 0.0f ? g[k] : 0.0f;
  }
  if (dres) {
    float r[8];
    load16(dres + (size_t)m * dres_ld + c, r);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] += g[k];
    store16(dres + (size_t)m * dres_ld + c, r);
  }
  if (scale) {
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] *= scale[c + k];
  }
  store16(dz + (size_t)m * dz_ld + c, g);
}

// column sums of a [M, C] fp16 matrix in fp32 (bias gradients), fixed order: thread per column block, rows ascending in
// chunks that a second pass adds in chunk order
__global__ void __launch_bounds__(256) colsum_partial_kernel(const __half* __restrict__ x, int ld, long long M, int C, int rows_per,
                                                             float* __restrict__ partial) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long m0 = (long long)blockIdx.y * rows_per, m1 = m0 + rows_per < M ? m0 + rows_per : M;
  float acc = 0.0f;
  for (long long m = m0; m < m1; ++m) acc += __half2float(x[(size_t)m * ld + c]);
  partial[(size_t)blockIdx.y * C + c] = acc;
}
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ partial, int chunks, int C, float scale,
                                                            float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.0f;
  for (int k = 0; k < chunks; ++k) acc += partial[(size_t)k * C + c];
  out[c] = acc * scale;
}

// temporal-mean backward (two_branch.py:249: the class scores are averaged over T'):  dx[a][b][p][c] += g[a][p*C + c] / B
__global__ void __launch_bounds__(256) mean_mid_bwd_kernel(const float* __restrict__ g, int A, int B, int P, int C, float gscale,
                                                           __half* __restrict__ dx, int ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)A * B * P * C) return;
  const int c = (int)(i % C);
  long long r = i / C;
  const int p = (int)(r % P); r /= P;
  const int b = (int)(r % B);
  const int a = (int)(r / B);
  __half* d = dx + ((((size_t)a * B + b) * P + p) * ld + c);
  *d = __float2half_rn(__half2float(*d) + g[(size_t)a * P * C + (size_t)p * C + c] * gscale / (float)B);
}

// fp32 [M, C] (scaled) accumulated into an fp16 channel slice
__global__ void __launch_bounds__(256) f32_accum_f16_kernel(const float* __restrict__ src, long long M, int C, float gscale,
                                                            __half* __restrict__ dst, int ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const long long m = i / C;
  const int c = (int)(i - m * C);
  __half* d = dst + (size_t)m * ld + c;
  *d = __float2half_rn(__half2float(*d) + src[i] * gscale);
}

// ---- max-pool backward (MaxPool3dTFPadding, i3dpt.py:114-126: zero ConstantPad3d, then MaxPool3d) ---------------------------
// Pass 1 (per output): which tap of the window holds the maximum -- first maximum in (kt, kh, kw) scan order with '>' like
// ATen's max_pool3d_with_indices; a padded position holds the value 0 and takes part (its gradient is dropped).
// Pass 2 (per input): gather dy from the windows whose recorded tap points at this input.  No atomics.
struct PoolGeom { int N, T, H, W, C, KT, KH, KW, ST, SH, SW, PT, PH, PW, OT, OH, OW, QT, QH, QW; };   // Q*: high-side zero padding

__global__ void __launch_bounds__(256) maxpool_argmax_kernel(PoolGeom g, const __half* __restrict__ x, int x_ld,
                                                             uint8_t* __restrict__ arg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * g.OT * g.OH * g.OW * g.C;
  if (i >= total) return;
  const int c = (int)(i % g.C);
  long long r = i / g.C;
  const int ow = (int)(r % g.OW); r /= g.OW;
  const int oh = (int)(r % g.OH); r /= g.OH;
  const int ot = (int)(r % g.OT);
  const int n = (int)(r / g.OT);
  float best = 0.0f;
  int bi = -1, tap = 0;
  for (int kt = 0; kt < g.KT; ++kt)
    for (int kh = 0; kh < g.KH; ++kh)
      for (int kw = 0; kw < g.KW; ++kw, ++tap) {
        const int t = ot * g.ST + kt - g.PT, h = oh * g.SH + kh - g.PH, w = ow * g.SW + kw - g.PW;
        // positions beyond the padded extent do not exist (ceil_mode overhang); inside the pad the value is 0
        const bool in_t = t >= 0 && t < g.T, in_h = h >= 0 && h < g.H, in_w = w >= 0 && w < g.W;
        const bool pad_t = t < 0 || (t >= g.T && t < g.T + g.QT), pad_h = h < 0 || (h >= g.H && h < g.H + g.QH),
                   pad_w = w < 0 || (w >= g.W && w < g.W + g.QW);
        if (!((in_t || pad_t) && (in_h || pad_h) && (in_w || pad_w))) continue;
        const bool real = in_t && in_h && in_w;
        const float v = real ? __half2float(x[((((size_t)n * g.T + t) * g.H + h) * g.W + w) * x_ld + c]) : 0.0f;
        if (bi < 0 || v > best) { best = v; bi = real ? tap : 254; }
      }
  arg[i] = (uint8_t)(bi < 0 ? 255 : bi);
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(PoolGeom g, const __half* __restrict__ dy, int dy_ld,
                                                          const uint8_t* __restrict__ arg, __half* __restrict__ dx, int dx_ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)g.N * g.T * g.H * g.W * g.C;
  if (i >= total) return;
  const int c = (int)(i % g.C);
  long long r = i / g.C;
  const int w = (int)(r % g.W); r /= g.W;
  const int h = (int)(r % g.H); r /= g.H;
  const int t = (int)(r % g.T);
  const int n = (int)(r / g.T);
  float acc = 0.0f;
  // windows (ot, oh, ow) with ot*ST + kt - PT == t  ->  kt = t + PT - ot*ST in [0, KT)
  for (int kt = 0; kt < g.KT; ++kt) {
    const int tn = t + g.PT - kt;
    if (tn < 0 || tn % g.ST) continue;
    const int ot = tn / g.ST;
    if (ot >= g.OT) continue;
    for (int kh = 0; kh < g.KH; ++kh) {
      const int hn = h + g.PH - kh;
      if (hn < 0 || hn % g.SH) continue;
      const int oh = hn / g.SH;
      if (oh >= g.OH) continue;
      for (int kw = 0; kw < g.KW; ++kw) {
        const int wn = w + g.PW - kw;
        if (wn < 0 || wn % g.SW) continue;
        const int ow = wn / g.SW;
        if (ow >= g.OW) continue;
        const size_t o = (((size_t)n * g.OT + ot) * g.OH + oh) * g.OW + ow;
        if (arg[o * g.C + c] == (uint8_t)((kt * g.KH + kh) * g.KW + kw)) acc += __half2float(dy[o * dy_ld + c]);
      }
    }
  }
  __half* d = dx + ((((size_t)n * g.T + t) * g.H + h) * g.W + w) * dx_ld + c;
  *d = __float2half_rn(__half2float(*d) + acc);
}

// ---- 1x1 convolution weight gradient on the tensor cores ------------------------------------------------------------
// dW[co][ci] = sum_m dz[m][co] * x[m][ci].  A CTA owns a 64 (co) x 64 (ci) tile of dW and one chunk of kWgChunk pixels;
// both operands are staged [32 pixels][64 channels] in shared memory and fed to wmma (fp16 x fp16 -> fp32) as A^T
// (column-major: element (co, m) at m * ld + co) and B (row-major: element (m, ci) at m * ld + ci).  The per-chunk
// partial tiles are summed by a second kernel in chunk order (deterministic), which also applies `scale`.
constexpr int kWgTile = 64, kWgPix = 32, kWgChunk = 2048;

// Filters larger than 1x1x1 (stride 1): tap (kt, kh, kw) pairs output pixel (n, t, h, w) with input pixel
// (t + kt - PT, h + kh - PH, w + kw - PW), zero outside the map (the TF-"SAME" padding, i3dpt.py:14-31); blockIdx.y also
// enumerates the taps and the staging of x applies the shift.
struct WgGeom { int T, H, W, KT, KH, KW, PT, PH, PW, taps; };

__global__ void __launch_bounds__(128) conv1x1_wgrad_partial_kernel(const __half* __restrict__ dz, int dz_ld,
                                                                    const __half* __restrict__ x, int x_ld, int M, int Cout,
                                                                    int Cin, WgGeom wg, float* __restrict__ partial) {
  using namespace nvcuda;
  __shared__ __align__(32) __half sA[kWgPix][kWgTile + 8];
  __shared__ __align__(32) __half sB[kWgPix][kWgTile + 8];
  const int tiles_ci = gridDim.y / wg.taps;
  const int tap = blockIdx.y / tiles_ci;
  const int co0 = blockIdx.x * kWgTile, ci0 = (blockIdx.y - tap * tiles_ci) * kWgTile, chunk = blockIdx.z;
  const int dkw = tap % wg.KW - wg.PW, dkh = (tap / wg.KW) % wg.KH - wg.PH, dkt = tap / (wg.KW * wg.KH) - wg.PT;
  const int m_beg = chunk * kWgChunk, m_end = min(M, m_beg + kWgChunk);
  const int warp = threadIdx.x >> 5;                 // 4 warps: warp w owns rows (co) [16 w, 16 w + 16) x all 64 ci
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) wmma::fill_fragment(acc[j], 0.0f);
  for (int m0 = m_beg; m0 < m_end; m0 += kWgPix) {
    // stage 32 pixels x 64 channels of each operand (16-byte vectors; rows past M and channels past C read as zero)
    for (int i = threadIdx.x; i < kWgPix * (kWgTile / 8); i += blockDim.x) {
      const int r = i / (kWgTile / 8), v = i - r * (kWgTile / 8);
      const int m = m0 + r;
      uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
      if (m < m_end) {
        if (co0 + v * 8 < Cout) a = *reinterpret_cast<const uint4*>(dz + (size_t)m * dz_ld + co0 + v * 8);
        long long src = m;
        bool ok = true;
        if (wg.taps > 1) {
          const int w = m % wg.W, h = (m / wg.W) % wg.H, t = (m / (wg.W * wg.H)) % wg.T;
          const int ws = w + dkw, hs = h + dkh, ts = t + dkt;
          ok = ws >= 0 && ws < wg.W && hs >= 0 && hs < wg.H && ts >= 0 && ts < wg.T;
          src = (long long)m + ((long long)dkt * wg.H + dkh) * wg.W + dkw;
        }
        if (ok && ci0 + v * 8 < Cin) b = *reinterpret_cast<const uint4*>(x + (size_t)src * x_ld + ci0 + v * 8);
      }
      *reinterpret_cast<uint4*>(&sA[r][v * 8]) = a;
      *reinterpret_cast<uint4*>(&sB[r][v * 8]) = b;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kWgPix; kk += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::col_major> fa;   // A(co, m) = sA[m][co]
      wmma::load_matrix_sync(fa, &sA[kk][warp * 16], kWgTile + 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> fb;  // B(m, ci) = sB[m][ci]
        wmma::load_matrix_sync(fb, &sB[kk][j * 16], kWgTile + 8);
        wmma::mma_sync(acc[j], fa, fb, acc[j]);
      }
    }
    __syncthreads();
  }
  float* out = partial + ((size_t)chunk * gridDim.x * gridDim.y + (size_t)blockIdx.x * gridDim.y + blockIdx.y) * (kWgTile * kWgTile);   // [chunk][co tile][tap][ci tile]
#pragma unroll
  for (int j = 0; j < 4; ++j) wmma::store_matrix_sync(out + (warp * 16) * kWgTile + j * 16, acc[j], kWgTile, wmma::mem_row_major);
}

// partial tiles are laid out [chunk][co tile][tap][ci tile][64 x 64]; dw is [Cout][taps][dw_ld >= Cin]
__global__ void __launch_bounds__(256) conv1x1_wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int tiles_co,
                                                                   int tiles_ci, int taps, int Cout, int Cin, float scale,
                                                                   float* __restrict__ dw, int dw_ld, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)Cout * taps * Cin) return;
  const int ci = (int)(i % Cin);
  const int tap = (int)((i / Cin) % taps);
  const int co = (int)(i / ((long long)Cin * taps));
  const int tco = co / kWgTile, tci = ci / kWgTile;
  const size_t off = (((size_t)tco * taps + tap) * tiles_ci + tci) * (kWgTile * kWgTile) + (size_t)(co - tco * kWgTile) * kWgTile + (ci - tci * kWgTile);
  const size_t stride = (size_t)tiles_co * taps * tiles_ci * (kWgTile * kWgTile);
  float acc = 0.0f;
  for (int c = 0; c < chunks; ++c) acc += partial[off + (size_t)c * stride];   // chunk order: deterministic
  float* d = dw + ((size_t)co * taps + tap) * dw_ld + ci;
  *d = accumulate ? *d + acc * scale : acc * scale;
}

}  // namespace step

using namespace step;

extern "C" int step_head_losses_f32(const float* logits, const float* local_loc, const float* first_loc, const float* last_loc,
                                    const float* tubes, const float* targets, int N, int cls, int T_len, int T, int Tc, float w_loc,
                                    float w_nb, float* loss_cls, float* loss_loc, float* loss_nb, int* flags, float* dlogits,
                                    float* dloc, float* dfirst, float* dlast, float* scratch, step_stream_t stream) {
  STEP_CHECK_ARG(N > 0 && cls > 0 && T > 0 && T_len >= T && T_len % T == 0, "head_losses: bad shape N=%d cls=%d T'=%d T=%d", N, cls, T_len, T);
  STEP_CHECK_ARG(logits && local_loc && first_loc && last_loc && tubes && targets && loss_cls && loss_loc && loss_nb && flags && scratch,
                 "head_losses: null pointer");
  STEP_CHECK_ARG((dfirst == nullptr) == (dlast == nullptr), "head_losses: dfirst / dlast go together");
  LossGeom g;
  const int chunks = T_len / T;
  g.N = N; g.cls = cls; g.T_len = T_len; g.Tc = Tc; g.half_T = T / 2;   // Tc = frames of first_loc / last_loc (python slice clipping)
  g.first_idx = T / 2; g.last_idx = (chunks - 1) * T + T / 2; g.centre = (chunks / 2) * T + T / 2;   // two_branch.py:226-228
  g.s0 = g.first_idx - g.half_T; g.e0 = g.last_idx - g.half_T;
  g.tgt_ld = 6 + cls; g.w_loc = w_loc; g.w_nb = w_nb;
  STEP_CHECK_ARG(Tc > g.half_T && g.s0 >= 0 && g.s0 + Tc <= T_len && g.e0 + Tc <= T_len, "head_losses: chunk slices out of range");
  head_losses_kernel<<<1, 256, 0, cu(stream)>>>(g, logits, local_loc, first_loc, last_loc, tubes, targets, loss_cls, loss_loc, loss_nb,
                                                flags, dlogits, dloc, dfirst, dlast, scratch);
  STEP_LAUNCH_CHECK("head_losses_kernel");
  return 0;
}

extern "C" int step_roi_align_bwd_nhwc(const void* grad_out, int dtype, int out_ld, const float* rois, int R, float scale, int ph,
                                       int pw, int K, int H, int W, int C, int sampling_ratio, float* grad_in, int in_ld,
                                       step_stream_t stream) {
  STEP_CHECK_ARG(K > 0 && H > 0 && W > 0 && C > 0 && R >= 0 && ph > 0 && pw > 0, "roi_align_bwd_nhwc: bad shape");
  STEP_CHECK_ARG(dtype == STEP_F32 || dtype == STEP_F16, "roi_align_bwd_nhwc: bad dtype");
  STEP_CHECK_ARG(C % 4 == 0 && in_ld % 4 == 0 && out_ld % 4 == 0 && in_ld >= C && out_ld >= C, "roi_align_bwd_nhwc: C / ld must be multiples of 4");
  STEP_CHECK_ARG(grad_in && (R == 0 || (grad_out && rois)), "roi_align_bwd_nhwc: null pointer");
  STEP_CHECK_ARG((((uintptr_t)grad_in | (uintptr_t)grad_out) & 15) == 0, "roi_align_bwd_nhwc: pointers must be 16-byte aligned");
  if (dtype == STEP_F32)
    roi_align_bwd_nhwc_kernel<float><<<K, 256, 0, cu(stream)>>>((const float*)grad_out, out_ld, rois, R, scale, H, W, C, ph, pw,
                                                                 sampling_ratio, grad_in, in_ld);
  else
    roi_align_bwd_nhwc_kernel<__half><<<K, 256, 0, cu(stream)>>>((const __half*)grad_out, out_ld, rois, R, scale, H, W, C, ph, pw,
                                                                  sampling_ratio, grad_in, in_ld);
  STEP_LAUNCH_CHECK("roi_align_bwd_nhwc_kernel");
  return 0;
}

extern "C" int step_linear_small_n_bwd(const void* x, int dtype, int M, int K, int x_ld, const float* w, const float* dy, int Nn,
                                       float* dx, int dx_accumulate, float* dw, float* db, step_stream_t stream) {
  STEP_CHECK_ARG(M > 0 && K > 0 && Nn > 0 && dy, "linear_small_n_bwd: bad arguments");
  STEP_CHECK_ARG(dtype == STEP_F32 || dtype == STEP_F16, "linear_small_n_bwd: bad dtype");
  if (dw) {
    STEP_CHECK_ARG(x != nullptr, "linear_small_n_bwd: dw needs x");
    const long long tot = (long long)Nn * K;
    const int grid = ceil_div(tot, 256);
    if (dtype == STEP_F32) linear_bwd_dw_kernel<float><<<grid, 256, 0, cu(stream)>>>((const float*)x, x_ld, dy, M, K, Nn, dw, db);
    else linear_bwd_dw_kernel<__half><<<grid, 256, 0, cu(stream)>>>((const __half*)x, x_ld, dy, M, K, Nn, dw, db);
    STEP_LAUNCH_CHECK("linear_bwd_dw_kernel");
  }
  if (dx) {
    STEP_CHECK_ARG(w != nullptr, "linear_small_n_bwd: dx needs w");
    linear_bwd_dx_kernel<<<ceil_div((long long)M * K, 256), 256, 0, cu(stream)>>>(dy, w, M, K, Nn, dx, dx_accumulate);
    STEP_LAUNCH_CHECK("linear_bwd_dx_kernel");
  }
  return 0;
}

extern "C" size_t step_conv_wgrad_workspace_bytes(int M, int Cout, int Cin, int taps) {
  const size_t chunks = (size_t)ceil_div(M, kWgChunk), tco = (size_t)ceil_div(Cout, kWgTile), tci = (size_t)ceil_div(Cin, kWgTile);
  return chunks * tco * tci * (size_t)(taps > 0 ? taps : 1) * kWgTile * kWgTile * sizeof(float);
}
extern "C" size_t step_conv1x1_wgrad_workspace_bytes(int M, int Cout, int Cin) { return step_conv_wgrad_workspace_bytes(M, Cout, Cin, 1); }

extern "C" int step_conv1x1_wgrad_f16(const void* dz, int dz_ld, const void* x, int x_ld, int M, int Cout, int Cin, float scale,
                                      float* dw, int dw_ld, int accumulate, void* workspace, size_t ws_bytes,
                                      step_stream_t stream) {
  STEP_CHECK_ARG(M > 0 && Cout > 0 && Cin > 0 && dz && x && dw && workspace, "conv1x1_wgrad: bad arguments");
  STEP_CHECK_ARG(Cout % 8 == 0 && Cin % 8 == 0 && dz_ld % 8 == 0 && x_ld % 8 == 0 && dz_ld >= Cout && x_ld >= Cin && dw_ld >= Cin,
                 "conv1x1_wgrad: channel counts / strides must be multiples of 8");
  STEP_CHECK_ARG((((uintptr_t)dz | (uintptr_t)x) & 15) == 0, "conv1x1_wgrad: pointers must be 16-byte aligned");
  if (ws_bytes < step_conv1x1_wgrad_workspace_bytes(M, Cout, Cin))
    return fail(STEP_E_WORKSPACE, "conv1x1_wgrad: workspace %zu < %zu", ws_bytes, step_conv1x1_wgrad_workspace_bytes(M, Cout, Cin));
  const int chunks = ceil_div(M, kWgChunk), tco = ceil_div(Cout, kWgTile), tci = ceil_div(Cin, kWgTile);
  WgGeom wg = {1, 1, M, 1, 1, 1, 0, 0, 0, 1};
  conv1x1_wgrad_partial_kernel<<<dim3(tco, tci, chunks), 128, 0, cu(stream)>>>((const __half*)dz, dz_ld, (const __half*)x, x_ld, M, Cout,
                                                                               Cin, wg, (float*)workspace);
  STEP_LAUNCH_CHECK("conv1x1_wgrad_partial_kernel");
  conv1x1_wgrad_reduce_kernel<<<ceil_div((long long)Cout * Cin, 256), 256, 0, cu(stream)>>>((const float*)workspace, chunks, tco, tci, 1,
                                                                                            Cout, Cin, scale, dw, dw_ld, accumulate);
  STEP_LAUNCH_CHECK("conv1x1_wgrad_reduce_kernel");
  return 0;
}

extern "C" int step_conv_wgrad_f16(const void* dz, int dz_ld, const void* x, int x_ld, int N, int T, int H, int W, int Cout, int Cin,
                                   int KT, int KH, int KW, int PT, int PH, int PW, float scale, float* dw, int dw_ld, int accumulate,
                                   void* workspace, size_t ws_bytes, step_stream_t stream) {
  const long long M = (long long)N * T * H * W;
  const int taps = KT * KH * KW;
  STEP_CHECK_ARG(M > 0 && M < (1LL << 31) && Cout > 0 && Cin > 0 && taps > 0 && dz && x && dw && workspace, "conv_wgrad: bad arguments");
  STEP_CHECK_ARG(Cout % 8 == 0 && Cin % 8 == 0 && dz_ld % 8 == 0 && x_ld % 8 == 0 && dz_ld >= Cout && x_ld >= Cin && dw_ld >= Cin,
                 "conv_wgrad: channel counts / strides must be multiples of 8");
  STEP_CHECK_ARG((((uintptr_t)dz | (uintptr_t)x) & 15) == 0, "conv_wgrad: pointers must be 16-byte aligned");
  STEP_CHECK_ARG(PT >= 0 && PT < KT && PH >= 0 && PH < KH && PW >= 0 && PW < KW, "conv_wgrad: bad padding");
  if (ws_bytes < step_conv_wgrad_workspace_bytes((int)M, Cout, Cin, taps))
    return fail(STEP_E_WORKSPACE, "conv_wgrad: workspace %zu < %zu", ws_bytes, step_conv_wgrad_workspace_bytes((int)M, Cout, Cin, taps));
  const int chunks = ceil_div(M, kWgChunk), tco = ceil_div(Cout, kWgTile), tci = ceil_div(Cin, kWgTile);
  STEP_CHECK_ARG((long long)tci * taps <= 65535 && chunks <= 65535, "conv_wgrad: grid too large");
  WgGeom wg = {T, H, W, KT, KH, KW, PT, PH, PW, taps};
  conv1x1_wgrad_partial_kernel<<<dim3(tco, tci * taps, chunks), 128, 0, cu(stream)>>>((const __half*)dz, dz_ld, (const __half*)x, x_ld, (int)M,
                                                                                      Cout, Cin, wg, (float*)workspace);
  STEP_LAUNCH_CHECK("conv_wgrad_partial_kernel");
  conv1x1_wgrad_reduce_kernel<<<ceil_div((long long)Cout * taps * Cin, 256), 256, 0, cu(stream)>>>((const float*)workspace, chunks, tco, tci,
                                                                                                   taps, Cout, Cin, scale, dw, dw_ld, accumulate);
  STEP_LAUNCH_CHECK("conv_wgrad_reduce_kernel");
  return 0;
}

extern "C" int step_act_bwd_f16(const void* dy, int dy_ld, const void* y, int y_ld, const float* scale, int relu, long long M, int C,
                                void* dz, int dz_ld, void* dres, int dres_ld, step_stream_t stream) {
  STEP_CHECK_ARG(M > 0 && C > 0 && C % 8 == 0 && dy_ld % 8 == 0 && dz_ld % 8 == 0 && dy && dz && (!relu || (y && y_ld % 8 == 0)) &&
                 (!dres || dres_ld % 8 == 0), "act_bwd: bad arguments");
  STEP_CHECK_ARG((((uintptr_t)dy | (uintptr_t)y | (uintptr_t)dz | (uintptr_t)dres) & 15) == 0, "act_bwd: pointers must be 16-byte aligned");
  act_bwd_kernel<<<ceil_div(M * (C / 8), 256), 256, 0, cu(stream)>>>((const __half*)dy, dy_ld, (const __half*)y, y_ld, scale, relu, M, C,
                                                                    (__half*)dz, dz_ld, (__half*)dres, dres_ld);
  STEP_LAUNCH_CHECK("act_bwd_kernel");
  return 0;
}

extern "C" int step_colsum_f16(const void* x, int ld, long long M, int C, float scale, float* out, float* workspace /* >= 64*C floats */,
                               step_stream_t stream) {
  STEP_CHECK_ARG(x && out && workspace && M > 0 && C > 0, "colsum: bad arguments");
  const int chunks = (int)(M < 64 ? M : 64);
  const int rows_per = ceil_div(M, chunks);
  colsum_partial_kernel<<<dim3(ceil_div(C, 256), chunks), 256, 0, cu(stream)>>>((const __half*)x, ld, M, C, rows_per, workspace);
  STEP_LAUNCH_CHECK("colsum_partial_kernel");
  colsum_reduce_kernel<<<ceil_div(C, 256), 256, 0, cu(stream)>>>(workspace, chunks, C, scale, out);
  STEP_LAUNCH_CHECK("colsum_reduce_kernel");
  return 0;
}

extern "C" int step_mean_mid_bwd(const float* g, int A, int B, int P, int C, float gscale, void* dx, int ld, step_stream_t stream) {
  STEP_CHECK_ARG(g && dx && A > 0 && B > 0 && P > 0 && C > 0 && ld >= C, "mean_mid_bwd: bad arguments");
  mean_mid_bwd_kernel<<<ceil_div((long long)A * B * P * C, 256), 256, 0, cu(stream)>>>(g, A, B, P, C, gscale, (__half*)dx, ld);
  STEP_LAUNCH_CHECK("mean_mid_bwd_kernel");
  return 0;
}

extern "C" int step_f32_accum_f16(const float* src, long long M, int C, float gscale, void* dst, int ld, step_stream_t stream) {
  STEP_CHECK_ARG(src && dst && M > 0 && C > 0 && ld >= C, "f32_accum_f16: bad arguments");
  f32_accum_f16_kernel<<<ceil_div(M * C, 256), 256, 0, cu(stream)>>>(src, M, C, gscale, (__half*)dst, ld);
  STEP_LAUNCH_CHECK("f32_accum_f16_kernel");
  return 0;
}

extern "C" int step_maxpool3d_bwd_f16(const void* x, int x_ld, const void* dy, int dy_ld, int N, int T, int H, int W, int C, int KT,
                                      int KH, int KW, int ST, int SH, int SW, int PT, int PH, int PW, int pad_hi_t, int pad_hi_h,
                                      int pad_hi_w, int OT, int OH, int OW, void* dx /* accumulated in place */, int dx_ld, uint8_t* argmax_ws /* N*OT*OH*OW*C bytes */,
                                      step_stream_t stream) {
  STEP_CHECK_ARG(x && dy && dx && argmax_ws && N > 0 && T > 0 && H > 0 && W > 0 && C > 0 && KT * KH * KW < 254, "maxpool3d_bwd: bad arguments");
  PoolGeom g = {N, T, H, W, C, KT, KH, KW, ST, SH, SW, PT, PH, PW, OT, OH, OW, pad_hi_t, pad_hi_h, pad_hi_w};
  const long long n_out = (long long)N * OT * OH * OW * C, n_in = (long long)N * T * H * W * C;
  maxpool_argmax_kernel<<<ceil_div(n_out, 256), 256, 0, cu(stream)>>>(g, (const __half*)x, x_ld, argmax_ws);
  STEP_LAUNCH_CHECK("maxpool_argmax_kernel");
  maxpool_bwd_kernel<<<ceil_div(n_in, 256), 256, 0, cu(stream)>>>(g, (const __half*)dy, dy_ld, argmax_ws, (__half*)dx, dx_ld);
  STEP_LAUNCH_CHECK("maxpool_bwd_kernel");
  return 0;
}
