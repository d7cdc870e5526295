// pool_layout.cu -- HBM-bound channels-last kernels: TF-SAME max pooling, layout conversion at the
// module boundary, small reductions and the small-N linear layers of the head.
//
// MaxPool3dTFPadding (models/i3dpt.py:114-126) = ConstantPad3d(0) + MaxPool3d(ceil_mode=True):
// the padded cells hold 0 (not -inf), and windows may hang over the *padded* extent (ceil mode),
// where they see nothing.  Everything is 16-byte vectorised along C.
#include "common.cuh"

namespace step {

template <typename T>
__global__ void __launch_bounds__(256) maxpool3d_kernel(const T* __restrict__ x, int N, int T_, int H, int W, int C,
                                                        int in_ld, int KT, int KH, int KW, int ST, int SH, int SW,
                                                        int PT, int PH, int PW, int pad_hi_t, int pad_hi_h,
                                                        int pad_hi_w, int OT, int OH, int OW, T* __restrict__ y,
                                                        int out_ld) {
  constexpr int VN = Vec16<T>::N;
  const int nvec = C / VN;
  const long long total = (long long)N * OT * OH * OW * nvec;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(idx % nvec);
    long long pix = idx / nvec;
    int ow = (int)(pix % OW); long long r = pix / OW;
    int oh = (int)(r % OH); r /= OH;
    int ot = (int)(r % OT);
    int n = (int)(r / OT);
    float m[VN];
#pragma unroll
    for (int k = 0; k < VN; ++k) m[k] = -3.402823466e+38f;
    bool touches_pad = false, any = false;
    for (int kt = 0; kt < KT; ++kt) {
      int t = ot * ST + kt - PT;          // coordinate in the un-padded tensor
      if (t >= T_ + pad_hi_t) continue;   // beyond the padded extent (ceil_mode overhang)
      bool tp = (t < 0) || (t >= T_);
      for (int kh = 0; kh < KH; ++kh) {
        int h = oh * SH + kh - PH;
        if (h >= H + pad_hi_h) continue;
        bool hp = (h < 0) || (h >= H);
        for (int kw = 0; kw < KW; ++kw) {
          int w = ow * SW + kw - PW;
          if (w >= W + pad_hi_w) continue;
          if (tp || hp || w < 0 || w >= W) { touches_pad = true; continue; }
          float v[VN];
          load16(x + ((((size_t)n * T_ + t) * H + h) * W + w) * in_ld + cv * VN, v);
#pragma unroll
          for (int k = 0; k < VN; ++k) m[k] = fmaxf(m[k], v[k]);
          any = true;
        }
      }
    }
    if (touches_pad || !any) {
#pragma unroll
      for (int k = 0; k < VN; ++k) m[k] = fmaxf(m[k], 0.0f);
    }
    store16(y + (size_t)pix * out_ld + cv * VN, m);
  }
}

// x [A,B,P,C] (C contiguous, pixel stride ld) -> y [A, P*C]  (mean over B, fp32 accumulate in index order)
template <typename TI, typename TO>
__global__ void mean_mid_kernel(const TI* __restrict__ x, int A, int B, int P, int C, int ld, TO* __restrict__ y) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)A * P * C) return;
  int c = (int)(idx % C);
  int p = (int)((idx / C) % P);
  int a = (int)(idx / ((long long)C * P));
  float s = 0.0f;
  for (int b = 0; b < B; ++b) s += to_f32<TI>(x[(((size_t)a * B + b) * P + p) * ld + c]);
  y[idx] = from_f32<TO>(s / (float)B);
}

// clip [N,T,Cc,H,W] fp32 -> [N,T,H,W,ld]
template <typename T>
__global__ void clip_to_ndhwc_kernel(const float* __restrict__ clip, int N, int T_, int Cc, int H, int W,
                                     T* __restrict__ out, int ld) {
  long long total = (long long)N * T_ * H * W;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    long long hw = pix % ((long long)H * W);
    long long nt = pix / ((long long)H * W);
    T* o = out + (size_t)pix * ld;
    for (int c = 0; c < ld; ++c)
      o[c] = from_f32<T>(c < Cc ? clip[((size_t)nt * Cc + c) * H * W + hw] : 0.0f);
  }
}

// clip [N,T,Cc,H,W] fp32 -> s2d [N,T/2,H/2,W/2,ld] f16, channel ((rt*2+rh)*2+rw)*Cc + c
__global__ void clip_to_s2d_kernel(const float* __restrict__ clip, int N, int T_, int Cc, int H, int W,
                                   __half* __restrict__ out, int ld) {
  const int T2 = T_ / 2, H2 = H / 2, W2 = W / 2;
  long long total = (long long)N * T2 * H2 * W2;
  for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < total;
       pix += (long long)gridDim.x * blockDim.x) {
    int w2 = (int)(pix % W2); long long r = pix / W2;
    int h2 = (int)(r % H2); r /= H2;
    int t2 = (int)(r % T2);
    int n = (int)(r / T2);
    __half* o = out + (size_t)pix * ld;
    int ch = 0;
    for (int rt = 0; rt < 2; ++rt)
      for (int rh = 0; rh < 2; ++rh)
        for (int rw = 0; rw < 2; ++rw)
          for (int c = 0; c < Cc; ++c, ++ch)
            o[ch] = __float2half_rn(
                clip[((((size_t)n * T_ + 2 * t2 + rt) * Cc + c) * H + 2 * h2 + rh) * W + 2 * w2 + rw]);
    for (; ch < ld; ++ch) o[ch] = __float2half_rn(0.0f);
  }
}

// [N*S, C] (ld) -> [N, C, S] fp32 via a 32x32 smem transpose
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, int S, int C, int ld, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int s = s0 + i, c = c0 + threadIdx.x;
    if (s < S && c < C) tile[i][threadIdx.x] = to_f32<T>(in[((size_t)n * S + s) * ld + c]);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, s = s0 + threadIdx.x;
    if (s < S && c < C) out[((size_t)n * C + c) * S + s] = tile[threadIdx.x][i];
  }
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int S, int C, T* __restrict__ out, int ld) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, s = s0 + threadIdx.x;
    if (s < S && c < C) tile[i][threadIdx.x] = in[((size_t)n * C + c) * S + s];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int s = s0 + i, c = c0 + threadIdx.x;
    if (s < S && c < C) out[((size_t)n * S + s) * ld + c] = from_f32<T>(tile[threadIdx.x][i]);
  }
}

// Small-N linear: one CTA per row-pair block; each warp owns output columns, lanes stride K.
// y[m,n] = act(sum_k x[m,k] w[n,k] + b[n]).  K is long (12544), N tiny (4 / 60): the x row is read
// once per CTA from HBM, w stays in L2.
constexpr int kLinRows = 4;
template <typename T>
__global__ void __launch_bounds__(256) linear_small_n_kernel(const T* __restrict__ x, int M, int K, int x_ld,
                                                             const T* __restrict__ w, const float* __restrict__ bias,
                                                             int N, float* __restrict__ y, int y_ld, int act,
                                                             int accumulate, const int32_t* __restrict__ row_map) {
  constexpr int VN = Vec16<T>::N;
  const int m0 = blockIdx.x * kLinRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int kvec = K / VN;
  for (int n = warp; n < N; n += nwarps) {
    float acc[kLinRows];
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) acc[r] = 0.0f;
    const T* wr = w + (size_t)n * K;
    for (int kv = lane; kv < kvec; kv += 32) {
      float wv[VN];
      load16(wr + kv * VN, wv);
#pragma unroll
      for (int r = 0; r < kLinRows; ++r) {
        if (m0 + r < M) {
          float xv[VN];
          const int xr = row_map ? row_map[m0 + r] : m0 + r;
          load16(x + (size_t)xr * x_ld + kv * VN, xv);
#pragma unroll
          for (int k = 0; k < VN; ++k) acc[r] = fmaf(xv[k], wv[k], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kLinRows; ++r) {
      float v = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && m0 + r < M) {
        v += bias ? bias[n] : 0.0f;
        float* dst = y + (size_t)(m0 + r) * y_ld + n;
        if (accumulate) v += *dst;
        if (act == 1) v = 1.0f / (1.0f + expf(-v));
        *dst = v;
      }
    }
  }
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  long long cap = (long long)kNumSMs * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace step

using namespace step;

extern "C" int step_maxpool3d_fwd(const void* x, int dtype, int N, int T, int H, int W, int C, int in_ld, int KT,
                                  int KH, int KW, int ST, int SH, int SW, int PT, int PH, int PW, int pad_hi_t,
                                  int pad_hi_h, int pad_hi_w, int OT, int OH, int OW, void* y, int out_ld,
                                  step_stream_t stream) {
  const int vn = dtype == STEP_F16 ? 8 : 4;
  STEP_CHECK_ARG(dtype == STEP_F16 || dtype == STEP_F32, "maxpool3d: bad dtype");
  STEP_CHECK_ARG(x && y && N > 0 && T > 0 && H > 0 && W > 0, "maxpool3d: bad shape/pointer");
  STEP_CHECK_ARG(C % vn == 0 && in_ld % vn == 0 && out_ld % vn == 0, "maxpool3d: C/ld must be multiples of %d", vn);
  STEP_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "maxpool3d: pointers must be 16-byte aligned");
  long long total = (long long)N * OT * OH * OW * (C / vn);
  if (dtype == STEP_F16)
    maxpool3d_kernel<__half><<<grid_for(total, 256), 256, 0, cu(stream)>>>(
        (const __half*)x, N, T, H, W, C, in_ld, KT, KH, KW, ST, SH, SW, PT, PH, PW, pad_hi_t, pad_hi_h, pad_hi_w, OT,
        OH, OW, (__half*)y, out_ld);
  else
    maxpool3d_kernel<float><<<grid_for(total, 256), 256, 0, cu(stream)>>>(
        (const float*)x, N, T, H, W, C, in_ld, KT, KH, KW, ST, SH, SW, PT, PH, PW, pad_hi_t, pad_hi_h, pad_hi_w, OT,
        OH, OW, (float*)y, out_ld);
  STEP_LAUNCH_CHECK("maxpool3d_kernel");
  return 0;
}

extern "C" int step_mean_mid(const void* x, int dtype, int A, int B, int P, int C, int ld, void* y, int out_dtype,
                             step_stream_t stream) {
  STEP_CHECK_ARG(x && y && A > 0 && B > 0 && P > 0 && C > 0 && ld >= C, "mean_mid: bad args");
  long long total = (long long)A * P * C;
  int g = ceil_div(total, 256);
  if (dtype == STEP_F16 && out_dtype == STEP_F16)
    mean_mid_kernel<__half, __half><<<g, 256, 0, cu(stream)>>>((const __half*)x, A, B, P, C, ld, (__half*)y);
  else if (dtype == STEP_F16 && out_dtype == STEP_F32)
    mean_mid_kernel<__half, float><<<g, 256, 0, cu(stream)>>>((const __half*)x, A, B, P, C, ld, (float*)y);
  else if (dtype == STEP_F32 && out_dtype == STEP_F32)
    mean_mid_kernel<float, float><<<g, 256, 0, cu(stream)>>>((const float*)x, A, B, P, C, ld, (float*)y);
  else
    return fail(STEP_E_UNSUPPORTED, "mean_mid: dtype combination %d -> %d", dtype, out_dtype);
  STEP_LAUNCH_CHECK("mean_mid_kernel");
  return 0;
}

extern "C" int step_clip_to_ndhwc(const float* clip, int N, int T, int Cc, int H, int W, void* out, int dtype, int ld,
                                  step_stream_t stream) {
  STEP_CHECK_ARG(clip && out && N > 0 && T > 0 && Cc > 0 && H > 0 && W > 0 && ld >= Cc, "clip_to_ndhwc: bad args");
  long long total = (long long)N * T * H * W;
  if (dtype == STEP_F16)
    clip_to_ndhwc_kernel<__half><<<grid_for(total, 256), 256, 0, cu(stream)>>>(clip, N, T, Cc, H, W, (__half*)out, ld);
  else if (dtype == STEP_F32)
    clip_to_ndhwc_kernel<float><<<grid_for(total, 256), 256, 0, cu(stream)>>>(clip, N, T, Cc, H, W, (float*)out, ld);
  else
    return fail(STEP_E_ARG, "clip_to_ndhwc: bad dtype");
  STEP_LAUNCH_CHECK("clip_to_ndhwc_kernel");
  return 0;
}

extern "C" int step_clip_to_s2d_f16(const float* clip, int N, int T, int Cc, int H, int W, void* out, int ld,
                                    step_stream_t stream) {
  STEP_CHECK_ARG(clip && out && N > 0 && T > 0 && Cc > 0 && H > 0 && W > 0, "clip_to_s2d: bad args");
  STEP_CHECK_ARG(T % 2 == 0 && H % 2 == 0 && W % 2 == 0 && ld >= 8 * Cc, "clip_to_s2d: T,H,W must be even, ld >= 8*Cc");
  long long total = (long long)N * (T / 2) * (H / 2) * (W / 2);
  clip_to_s2d_kernel<<<grid_for(total, 256), 256, 0, cu(stream)>>>(clip, N, T, Cc, H, W, (__half*)out, ld);
  STEP_LAUNCH_CHECK("clip_to_s2d_kernel");
  return 0;
}

extern "C" int step_nhwc_to_nchw_f32(const void* in, int dtype, int N, int S, int C, int ld, float* out,
                                     step_stream_t stream) {
  STEP_CHECK_ARG(in && out && N > 0 && S > 0 && C > 0 && ld >= C && N <= 65535, "nhwc_to_nchw: bad args");
  dim3 grid(ceil_div(S, 32), ceil_div(C, 32), N), block(32, 8);
  if (dtype == STEP_F16)
    nhwc_to_nchw_kernel<__half><<<grid, block, 0, cu(stream)>>>((const __half*)in, S, C, ld, out);
  else if (dtype == STEP_F32)
    nhwc_to_nchw_kernel<float><<<grid, block, 0, cu(stream)>>>((const float*)in, S, C, ld, out);
  else
    return fail(STEP_E_ARG, "nhwc_to_nchw: bad dtype");
  STEP_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return 0;
}

extern "C" int step_nchw_to_nhwc(const float* in, int N, int S, int C, void* out, int dtype, int ld,
                                 step_stream_t stream) {
  STEP_CHECK_ARG(in && out && N > 0 && S > 0 && C > 0 && ld >= C && N <= 65535, "nchw_to_nhwc: bad args");
  dim3 grid(ceil_div(S, 32), ceil_div(C, 32), N), block(32, 8);
  if (dtype == STEP_F16)
    nchw_to_nhwc_kernel<__half><<<grid, block, 0, cu(stream)>>>(in, S, C, (__half*)out, ld);
  else if (dtype == STEP_F32)
    nchw_to_nhwc_kernel<float><<<grid, block, 0, cu(stream)>>>(in, S, C, (float*)out, ld);
  else
    return fail(STEP_E_ARG, "nchw_to_nhwc: bad dtype");
  STEP_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}

extern "C" int step_linear_small_n(const void* x, int dtype, int M, int K, int x_ld, const void* w, const float* bias,
                                   int N, float* y, int y_ld, int act, int accumulate, const int32_t* row_map,
                                   step_stream_t stream) {
  const int vn = dtype == STEP_F16 ? 8 : 4;
  STEP_CHECK_ARG(dtype == STEP_F16 || dtype == STEP_F32, "linear_small_n: bad dtype");
  STEP_CHECK_ARG(x && w && y && M >= 0 && K > 0 && N > 0 && N <= 64 && y_ld >= N, "linear_small_n: bad args (N <= 64)");
  STEP_CHECK_ARG(K % vn == 0 && x_ld % vn == 0, "linear_small_n: K and x_ld must be multiples of %d", vn);
  STEP_CHECK_ARG((((uintptr_t)x | (uintptr_t)w) & 15) == 0, "linear_small_n: pointers must be 16-byte aligned");
  if (M == 0) return 0;
  int grid = ceil_div(M, kLinRows);
  if (dtype == STEP_F16)
    linear_small_n_kernel<__half><<<grid, 256, 0, cu(stream)>>>((const __half*)x, M, K, x_ld, (const __half*)w, bias, N,
                                                                y, y_ld, act, accumulate, row_map);
  else
    linear_small_n_kernel<float><<<grid, 256, 0, cu(stream)>>>((const float*)x, M, K, x_ld, (const float*)w, bias, N, y,
                                                               y_ld, act, accumulate, row_map);
  STEP_LAUNCH_CHECK("linear_small_n_kernel");
  return 0;
}
