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

}  // namespace

template <typename T>
int subsample_conv1(const float* x, const float* w, const float* bias, T* out, int B, int n_mel, int Tin,
                    int T1, int T1p, int F1, int C, cudaStream_t stream) {
  CFM_REQUIRE(B > 0 && T1 > 0 && F1 > 0, "input too short for the subsampling convs");
  dim3 grid(static_cast<unsigned>(ceil_div(T1p, C1_TT)), static_cast<unsigned>(B));
  const size_t smem = static_cast<size_t>(n_mel) * (2 * C1_TT + 1) * sizeof(float);
  subsample_conv1_kernel<T><<<grid, 256, smem, stream>>>(x, w, bias, out, n_mel, Tin, T1, T1p, F1, C);
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
