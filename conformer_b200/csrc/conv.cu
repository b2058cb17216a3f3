// Convolution kernels on CUDA cores:
//   * subsample_conv1: Conv2d(1,C,3,stride 2)+ReLU of the front-end (convolution.py:37,44-45),
//     written channels-last so that the second conv becomes a K-contiguous implicit GEMM.
//   * dwconv_glu_bn_swish: GLU -> depthwise Conv1d(k, pad (k-1)/2) -> BatchNorm(eval, folded) -> Swish
//     of the ConvolutionModule (convolution.py:25-28), channels-last, time halo staged in shared memory.
#include "ops.h"

namespace cfm {
namespace {

constexpr int C1_TT = 8;   // conv1: output time steps per block

template <typename T>
__global__ void __launch_bounds__(256) subsample_conv1_kernel(const float* __restrict__ x,
                                                              const float* __restrict__ w,
                                                              const float* __restrict__ bias, T* __restrict__ out,
                                                              int n_mel, int Tin, int T1, int T1p, int F1, int C) {
  extern __shared__ float xs[];                       // [n_mel][2*C1_TT+1]
  constexpr int XW = 2 * C1_TT + 1;
  const int b = blockIdx.y, t0 = blockIdx.x * C1_TT;
  for (int i = threadIdx.x; i < n_mel * XW; i += blockDim.x) {
    const int m = i / XW, j = i - m * XW;
    const int t = 2 * t0 + j;
    xs[i] = (t < Tin) ? x[(static_cast<int64_t>(b) * n_mel + m) * Tin + t] : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float wr[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) wr[i] = w[c * 9 + i];          // [kh][kw]
    const float bc = bias[c];
    for (int tt = 0; tt < C1_TT; ++tt) {
      const int t1 = t0 + tt;
      if (t1 >= T1p) break;
      T* orow = out + ((static_cast<int64_t>(b) * T1p + t1) * F1) * C + c;
      if (t1 >= T1) {                                           // padding row (T1p = T1 rounded up to even)
        for (int f = 0; f < F1; ++f) orow[static_cast<int64_t>(f) * C] = from_f32<T>(0.f);
        continue;
      }
      for (int f = 0; f < F1; ++f) {
        float acc = bc;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) acc = fmaf(wr[kh * 3 + kw], xs[(2 * f + kh) * XW + 2 * tt + kw], acc);
        orow[static_cast<int64_t>(f) * C] = from_f32<T>(fmaxf(acc, 0.f));
      }
    }
  }
}

// bf16 variant: one thread = 8 channels of one (t, f) output (16-byte stores; every broadcast shared-memory
// read of an input pixel feeds 8 FMAs), the frequency axis is split over the remaining thread groups.
__global__ void __launch_bounds__(256) subsample_conv1_bf16_kernel(const float* __restrict__ x,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ bias,
                                                                   bf16* __restrict__ out, int n_mel, int Tin, int T1,
                                                                   int T1p, int F1, int C) {
  extern __shared__ float xs[];                       // [n_mel][2*C1_TT+1]
  constexpr int XW = 2 * C1_TT + 1;
  const int b = blockIdx.y, t0 = blockIdx.x * C1_TT;
  for (int i = threadIdx.x; i < n_mel * XW; i += blockDim.x) {
    const int m = i / XW, j = i - m * XW;
    const int t = 2 * t0 + j;
    xs[i] = (t < Tin) ? x[(static_cast<int64_t>(b) * n_mel + m) * Tin + t] : 0.f;
  }
  __syncthreads();
  const int groups = C >> 3;                          // 8-channel groups
  const int ngrp = blockDim.x / groups;               // thread groups along the frequency axis (>= 1)
  const int cg = threadIdx.x % groups, fg = threadIdx.x / groups;
  if (fg >= ngrp) return;
  float wr[8][9], br[8];
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    br[ch] = bias[8 * cg + ch];
#pragma unroll
    for (int i = 0; i < 9; ++i) wr[ch][i] = w[(8 * cg + ch) * 9 + i];
  }
  for (int tt = 0; tt < C1_TT; ++tt) {
    const int t1 = t0 + tt;
    if (t1 >= T1p) break;
    bf16* orow = out + ((static_cast<int64_t>(b) * T1p + t1) * F1) * C + 8 * cg;
    for (int f = fg; f < F1; f += ngrp) {
      float a[8];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) a[ch] = br[ch];
      if (t1 < T1) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float xv = xs[(2 * f + kh) * XW + 2 * tt + kw];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) a[ch] = fmaf(wr[ch][kh * 3 + kw], xv, a[ch]);
          }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) a[ch] = fmaxf(a[ch], 0.f);
      } else {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) a[ch] = 0.f;     // padding row (T1p = T1 rounded up to even)
      }
      uint4 o;
      o.x = pack_bf16x2(a[0], a[1]); o.y = pack_bf16x2(a[2], a[3]);
      o.z = pack_bf16x2(a[4], a[5]); o.w = pack_bf16x2(a[6], a[7]);
      *reinterpret_cast<uint4*>(orow + static_cast<int64_t>(f) * C) = o;
    }
  }
}

// ---- depthwise conv -----------------------------------------------------------------------------
constexpr int DW_CH = 64;      // channels per block (one per thread column)
constexpr int DW_TT = 64;      // output time steps per block
constexpr int DW_RT = 8;       // outputs per thread per pass (register window)

template <typename T, int KS>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ u, const float* __restrict__ w,
                                                     const float* __restrict__ bias, T* __restrict__ y, int Tn,
                                                     int d, int ksize_rt) {
  const int ksize = KS > 0 ? KS : ksize_rt;
  const int half = (ksize - 1) / 2;
  extern __shared__ float sm[];
  float* gs = sm;                                              // [(DW_TT + ksize - 1)][DW_CH]
  float* ws = sm + (DW_TT + ksize - 1) * DW_CH;               // [ksize][DW_CH]
  const int b = blockIdx.z, c0 = blockIdx.y * DW_CH, t0 = blockIdx.x * DW_TT;
  const int cl = threadIdx.x % DW_CH, tg = threadIdx.x / DW_CH;   // 4 time groups
  const int c = c0 + cl;
  const bool c_ok = c < d;
  // stage GLU(u) for t in [t0-half, t0+DW_TT+half); zero outside [0,Tn) (Conv1d zero padding)
  const int rows = DW_TT + ksize - 1;
  for (int r = tg; r < rows; r += 4) {
    const int t = t0 - half + r;
    float g = 0.f;
    if (c_ok && t >= 0 && t < Tn) {
      const T* ur = u + (static_cast<int64_t>(b) * Tn + t) * 2 * d;
      g = to_f32(ur[c]) * sigmoid_f(to_f32(ur[d + c]));         // activation.py:16-17
    }
    gs[r * DW_CH + cl] = g;
  }
  for (int j = tg; j < ksize; j += 4) ws[j * DW_CH + cl] = c_ok ? w[static_cast<int64_t>(j) * d + c] : 0.f;
  __syncthreads();
  if (!c_ok) return;
  const float bc = bias[c];
  for (int chunk = tg; chunk < DW_TT / DW_RT; chunk += 4) {
    const int tl = chunk * DW_RT;
    if (t0 + tl >= Tn) break;
    float acc[DW_RT];
#pragma unroll
    for (int i = 0; i < DW_RT; ++i) acc[i] = bc;
    if (KS > 0) {
      float win[DW_RT + (KS > 0 ? KS : 1) - 1];
#pragma unroll
      for (int i = 0; i < DW_RT + KS - 1; ++i) win[i] = gs[(tl + i) * DW_CH + cl];
#pragma unroll
      for (int j = 0; j < KS; ++j) {
        const float wj = ws[j * DW_CH + cl];
#pragma unroll
        for (int i = 0; i < DW_RT; ++i) acc[i] = fmaf(wj, win[i + j], acc[i]);
      }
    } else {
      for (int j = 0; j < ksize; ++j) {
        const float wj = ws[j * DW_CH + cl];
#pragma unroll
        for (int i = 0; i < DW_RT; ++i) acc[i] = fmaf(wj, gs[(tl + i + j) * DW_CH + cl], acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < DW_RT; ++i) {
      const int t = t0 + tl + i;
      if (t < Tn) y[(static_cast<int64_t>(b) * Tn + t) * d + c] = from_f32<T>(swish_f(acc[i]));
    }
  }
}


// ---- depthwise conv, bf16 fast path -------------------------------------------------------------
// Input already gated (GLU fused into the pointwise_conv_1 GEMM epilogue).  One CTA = 64 time steps x 64
// channels of one utterance: the (64+k-1) x 64 halo tile is fetched with 16-byte loads (8 channels per
// thread), widened to fp32 in shared memory, and every thread then produces 8 consecutive time steps of a
// CHANNEL PAIR from a register window, so the 31-tap MAC runs as packed fma.rn.f32x2 (two channels per
// instruction) with one conflict-free 64-bit shared load per window element.
constexpr int DV_TT = 64, DV_CH = 64, DV_RT = 8;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ra, rb, rc, rd;
  ra = (static_cast<uint64_t>(__float_as_uint(a.y)) << 32) | __float_as_uint(a.x);
  rb = (static_cast<uint64_t>(__float_as_uint(b.y)) << 32) | __float_as_uint(b.x);
  rc = (static_cast<uint64_t>(__float_as_uint(c.y)) << 32) | __float_as_uint(c.x);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return make_float2(__uint_as_float(static_cast<uint32_t>(rd)), __uint_as_float(static_cast<uint32_t>(rd >> 32)));
}

template <int KS>
__global__ void __launch_bounds__(256) dwconv_bf16_kernel(const bf16* __restrict__ g, const float* __restrict__ w,
                                                          const float* __restrict__ bias, bf16* __restrict__ y, int Tn,
                                                          int d) {
  constexpr int HALF = (KS - 1) / 2, ROWS = DV_TT + KS - 1;
  pdl_wait();
  __shared__ __align__(16) float gs[ROWS * DV_CH];
  __shared__ __align__(16) float ws[KS * DV_CH];
  const int b = blockIdx.z, c0 = blockIdx.y * DV_CH, t0 = blockIdx.x * DV_TT;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < ROWS * (DV_CH / 8); idx += 256) {
    const int r = idx >> 3, v = idx & 7;
    const int t = t0 - HALF + r;
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < Tn) raw = *reinterpret_cast<const uint4*>(g + (static_cast<int64_t>(b) * Tn + t) * d + c0 + v * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
    float4 lo, hi;
    lo.x = __low2float(h[0]); lo.y = __high2float(h[0]); lo.z = __low2float(h[1]); lo.w = __high2float(h[1]);
    hi.x = __low2float(h[2]); hi.y = __high2float(h[2]); hi.z = __low2float(h[3]); hi.w = __high2float(h[3]);
    float4* dst = reinterpret_cast<float4*>(&gs[r * DV_CH + v * 8]);
    dst[0] = lo;
    dst[1] = hi;
  }
  for (int idx = tid; idx < KS * DV_CH; idx += 256) {
    const int j = idx / DV_CH, c = idx - j * DV_CH;
    ws[idx] = w[static_cast<int64_t>(j) * d + c0 + c];
  }
  __syncthreads();
  const int cp = tid & 31, tg = tid >> 5;                      // channel pair, time group (8 outputs)
  const int tl = tg * DV_RT;
  if (t0 + tl >= Tn) return;
  const float2 bc = *reinterpret_cast<const float2*>(bias + c0 + 2 * cp);
  float2 acc[DV_RT];
#pragma unroll
  for (int i = 0; i < DV_RT; ++i) acc[i] = bc;
  float2 win[DV_RT + KS - 1];
#pragma unroll
  for (int i = 0; i < DV_RT + KS - 1; ++i) win[i] = *reinterpret_cast<const float2*>(&gs[(tl + i) * DV_CH + 2 * cp]);
#pragma unroll
  for (int j = 0; j < KS; ++j) {
    const float2 wj = *reinterpret_cast<const float2*>(&ws[j * DV_CH + 2 * cp]);
#pragma unroll
    for (int i = 0; i < DV_RT; ++i) acc[i] = ffma2(wj, win[i + j], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < DV_RT; ++i) {
    const int t = t0 + tl + i;
    if (t < Tn)
      *reinterpret_cast<uint32_t*>(y + (static_cast<int64_t>(b) * Tn + t) * d + c0 + 2 * cp) =
          pack_bf16x2(swish_tanh(acc[i].x), swish_tanh(acc[i].y));
  }
}

}  // namespace

template <typename T>
int subsample_conv1(const float* x, const float* w, const float* bias, T* out, int B, int n_mel, int Tin,
                    int T1, int T1p, int F1, int C, cudaStream_t stream) {
  CFM_REQUIRE(B > 0 && T1 > 0 && F1 > 0, "input too short for the subsampling convs");
  dim3 grid(static_cast<unsigned>(ceil_div(T1p, C1_TT)), static_cast<unsigned>(B));
  const size_t smem = static_cast<size_t>(n_mel) * (2 * C1_TT + 1) * sizeof(float);
  if (sizeof(T) == 2 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0) {
    subsample_conv1_bf16_kernel<<<grid, 256, smem, stream>>>(x, w, bias, reinterpret_cast<bf16*>(out), n_mel, Tin, T1,
                                                             T1p, F1, C);
  } else {
    subsample_conv1_kernel<T><<<grid, 256, smem, stream>>>(x, w, bias, out, n_mel, Tin, T1, T1p, F1, C);
  }
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
template int subsample_conv1<float>(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, cudaStream_t);
template int subsample_conv1<bf16>(const float*, const float*, const float*, bf16*, int, int, int, int, int, int, int, cudaStream_t);

template <typename T>
int dwconv_glu_bn_swish(const T* u, const float* w, const float* bias, T* y, int B, int Tn, int d, int ksize,
                        cudaStream_t stream) {
  CFM_REQUIRE(ksize % 2 == 1 && ksize >= 1 && ksize <= 255, "depthwise kernel size must be odd");
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, DW_TT)), static_cast<unsigned>(ceil_div(d, DW_CH)),
            static_cast<unsigned>(B));
  const size_t smem = static_cast<size_t>(DW_TT + 2 * ksize - 1) * DW_CH * sizeof(float);
  CFM_REQUIRE(smem <= 48 * 1024, "depthwise kernel too wide for the staging buffer");
  if (ksize == 31)
    dwconv_kernel<T, 31><<<grid, 256, smem, stream>>>(u, w, bias, y, Tn, d, ksize);
  else if (ksize == 7)
    dwconv_kernel<T, 7><<<grid, 256, smem, stream>>>(u, w, bias, y, Tn, d, ksize);
  else
    dwconv_kernel<T, 0><<<grid, 256, smem, stream>>>(u, w, bias, y, Tn, d, ksize);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
template int dwconv_glu_bn_swish<float>(const float*, const float*, const float*, float*, int, int, int, int, cudaStream_t);
template int dwconv_glu_bn_swish<bf16>(const bf16*, const float*, const float*, bf16*, int, int, int, int, cudaStream_t);

}  // namespace cfm

namespace cfm {
int dwconv_bn_swish_bf16(const bf16* g, const float* w, const float* bias, bf16* y, int B, int Tn, int d, int ksize,
                         cudaStream_t stream) {
  CFM_REQUIRE(d % DV_CH == 0 && ksize == 31, "fast depthwise path needs channels % 64 == 0 and kernel size 31");
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, DV_TT)), static_cast<unsigned>(d / DV_CH), static_cast<unsigned>(B));
  CFM_CUDA_OK(launch_pdl(dwconv_bf16_kernel<31>, grid, dim3(256), 0, stream, 1, g, w, bias, y, Tn, d));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
}  // namespace cfm
