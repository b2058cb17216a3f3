// Element-wise, normalisation and reduction kernels of the TRAINING step (BASELINE config 4; train.py:232-243 runs
// forward -> CTC loss -> backward -> optimizer step over the modules of model/utils/*.py).  Everything dense goes through
// the GEMM kernels (gemm_tc.cu / gemm_f32.cu); this file holds what surrounds them: dropout (nn.Dropout at ffn.py:19,21,
// attention.py:12, convolution.py:19, encoder.py:14), LayerNorm backward, Swish / GLU backward, depthwise conv forward
// and backward, training-mode BatchNorm1d (batch statistics over every frame of the batch, padded ones included, running
// statistics updated with momentum 0.1, convolution.py:16,27), bias gradients, transposes for the weight-gradient GEMMs,
// and the front-end's ReLU / im2col / col2im steps.
//
// Dropout masks are never stored: keep(seed, site, element) is a stateless hash, recomputed in the backward pass.
// Activations are stored as TA (bf16 in CFM_COMPUTE_BF16, fp32 in CFM_COMPUTE_F32); the residual stream, every
// gradient of it, LayerNorm / BatchNorm arithmetic and all parameter gradients are fp32.
#include "train.h"

namespace cfm {
namespace {

inline unsigned nblk(int64_t n, int t) { return static_cast<unsigned>((n + t - 1) / t); }

// ------------------------------------------------------------------------------------------- element-wise
// 4 consecutive elements per thread (n % 4 == 0 and 16-byte aligned buffers: every caller passes [rows, multiple of 8])
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
  v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void st4(bf16* p, const float (&v)[4]) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

template <typename TOut>
__global__ void cast_scale_drop_kernel(const float* __restrict__ in, TOut* __restrict__ out, int64_t n, float scale, Drop dr) {
  pdl_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float v[4];
  ld4(in + i, v);
  const DropK dk = dr.load();
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] *= scale * dk.factor(i + e);
  st4(out + i, v);
}
// bf16 activations: the result is rounded to bf16, so the one-MUFU forms (tanh.approx, error ~5e-4) are exact enough;
// fp32 activations (the gradient-parity mode) keep expf and a true division.
template <typename TA> __device__ __forceinline__ float swish_act(float x) {
  if constexpr (sizeof(TA) == 2) return swish_tanh(x);
  else return swish_f(x);
}
template <typename TA> __device__ __forceinline__ float sigmoid_act(float x) {
  if constexpr (sizeof(TA) == 2) return sigmoid_tanh(x);
  else return sigmoid_f(x);
}
template <typename TA>
__global__ void swish_drop_fwd_kernel(const TA* __restrict__ hp, TA* __restrict__ out, int64_t n, Drop dr) {
  pdl_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float v[4];
  ld4(hp + i, v);
  const DropK dk = dr.load();
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = swish_act<TA>(v[e]) * dk.factor(i + e);
  st4(out + i, v);
}
// bf16 activations, 8 per thread (16-byte accesses; n % 8 == 0)
__global__ void swish_drop_fwd8_kernel(const bf16* __restrict__ hp, bf16* __restrict__ out, int64_t n, Drop dr) {
  pdl_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  float v[8];
  ld8(hp + i, v);
  const DropK dk = dr.load();
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = swish_tanh(v[e]) * dk.factor(i + e);
  *reinterpret_cast<uint4*>(out + i) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
template <typename TA>
__global__ void swish_drop_bwd_kernel(const float* __restrict__ dh, const TA* __restrict__ hp, TA* __restrict__ out, int64_t n, Drop dr) {
  pdl_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float x[4], g[4];
  ld4(hp + i, x);
  ld4(dh + i, g);
  const DropK dk = dr.load();
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float s = sigmoid_act<TA>(x[e]);
    g[e] *= dk.factor(i + e) * (s + x[e] * s * (1.f - s));
  }
  st4(out + i, g);
}
__global__ void residual_drop_add_kernel(const float* __restrict__ x, const float* __restrict__ t, float* __restrict__ out, int64_t n,
                                         float scale, Drop dr) {
  pdl_wait();
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float a[4], b[4];
  ld4(x + i, a);
  ld4(t + i, b);
  const DropK dk = dr.load();
#pragma unroll
  for (int e = 0; e < 4; ++e) a[e] += scale * dk.factor(i + e) * b[e];
  st4(out + i, a);
}
__global__ void drop_inplace_kernel(float* __restrict__ x, int64_t n, Drop dr) {
  pdl_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= dr.load().factor(i);
}
template <typename TA>
__global__ void glu_fwd_kernel(const TA* __restrict__ u, TA* __restrict__ g, int64_t rows, int d) {
  pdl_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  const int64_t r = i / d;
  const int c = static_cast<int>(i - r * d);
  g[i] = from_f32<TA>(to_f32(u[r * 2 * d + c]) * sigmoid_f(to_f32(u[r * 2 * d + d + c])));     // activation.py:16-17
}
template <typename TA>
__global__ void glu_bwd_kernel(const float* __restrict__ dg, const TA* __restrict__ u, TA* __restrict__ du, int64_t rows, int d) {
  pdl_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  const int64_t r = i / d;
  const int c = static_cast<int>(i - r * d);
  const float a = to_f32(u[r * 2 * d + c]), s = sigmoid_f(to_f32(u[r * 2 * d + d + c])), g = dg[i];
  du[r * 2 * d + c] = from_f32<TA>(g * s);
  du[r * 2 * d + d + c] = from_f32<TA>(g * a * s * (1.f - s));
}
template <typename TA>
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const TA* __restrict__ y, TA* __restrict__ out, int64_t n) {
  pdl_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = from_f32<TA>(to_f32(y[i]) > 0.f ? dy[i] : 0.f);
}

// [M,N] (row stride ld) -> [N, Mp] with the columns M..Mp-1 zeroed: operands of the weight-gradient GEMMs, whose
// contraction runs over the M rows
template <typename T>
__global__ void transpose_kernel(const T* __restrict__ in, int ld, T* __restrict__ out, int M, int N, int Mp) {
  pdl_wait();
  __shared__ T tile[32][33];
  const int m0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int m = m0 + r, n = n0 + tx;
    tile[r][tx] = (m < M && n < N) ? in[static_cast<int64_t>(m) * ld + n] : from_f32<T>(0.f);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, m = m0 + tx;
    if (n < N && m < Mp) out[static_cast<int64_t>(n) * Mp + m] = tile[tx][r];
  }
}

// column sums: out[n] += sum_m in[m, n]   (bias gradients)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ in, int ld, float* __restrict__ out, int M, int N, int rows_per_block) {
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc = 0.f;
#pragma unroll 8
  for (int m = m0; m < m1; ++m) acc += to_f32(in[static_cast<int64_t>(m) * ld + n]);
  atomicAdd(out + n, acc);
}

// ------------------------------------------------------------------------------------------- LayerNorm backward
// dx (+)= (rstd / d) * (d * gy - sum(gy) - xhat * sum(gy * xhat)),  gy = dy * gamma;  dgamma += dy * xhat, dbeta += dy.
// One warp per row (row in registers); the parameter gradients are reduced per block in shared memory, then atomically.
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ gamma, float* dx, int accumulate,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int d,
                                                     int rows_per_warp) {
  pdl_wait();
  extern __shared__ float red[];                                // [2][d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float gsum[NV], bsum[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) gsum[i] = bsum[i] = 0.f;
  const int64_t r0 = (static_cast<int64_t>(blockIdx.x) * nwarps + warp) * rows_per_warp;
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const int64_t row = r0 + rr;
    if (row >= rows) break;
    float xv[NV], gv[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      xv[i] = c < d ? x[row * d + c] : 0.f;
      s += xv[i];
    }
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      xv[i] = c < d ? xv[i] - mean : 0.f;
      q += xv[i] * xv[i];
    }
    const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      xv[i] *= rstd;                                            // xhat
      const float dyv = c < d ? dy[row * d + c] : 0.f;
      gv[i] = c < d ? dyv * gamma[c] : 0.f;
      s1 += gv[i];
      s2 += gv[i] * xv[i];
      gsum[i] += dyv * xv[i];
      bsum[i] += dyv;
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < d) {
        const float v = rstd * (gv[i] - s1 - xv[i] * s2);
        dx[row * d + c] = accumulate ? dx[row * d + c] + v : v;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < d) {
      atomicAdd(&red[c], gsum[i]);
      atomicAdd(&red[d + c], bsum[i]);
    }
  }
  __syncthreads();
  if ((d & 3) == 0 && ((reinterpret_cast<uintptr_t>(dgamma) | reinterpret_cast<uintptr_t>(dbeta)) & 15) == 0) {
    for (int i = threadIdx.x * 4; i < 2 * d; i += blockDim.x * 4) {       // 16-byte vector reductions: 4x fewer L2 atomics
      float* dst = i < d ? dgamma + i : dbeta + (i - d);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(red[i]), "f"(red[i + 1]), "f"(red[i + 2]),
                   "f"(red[i + 3])
                   : "memory");
    }
  } else {
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      atomicAdd(dgamma + i, red[i]);
      atomicAdd(dbeta + i, red[d + i]);
    }
  }
}

// d = 256 * G: lane owns 8 contiguous columns of every 256-column group (two 16-byte accesses), nothing is predicated.
template <int G>
__global__ void __launch_bounds__(256) ln_bwd_vec_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                         const float* __restrict__ gamma, float* dx, int accumulate,
                                                         float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows,
                                                         int rows_per_warp) {
  pdl_wait();
  constexpr int d = 256 * G, NV = 8 * G;
  __shared__ float red[2 * d];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float gam[NV], gsum[NV], bsum[NV];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(gamma + g * 256 + lane * 8)), b = __ldg(reinterpret_cast<const float4*>(gamma + g * 256 + lane * 8) + 1);
    gam[8 * g + 0] = a.x; gam[8 * g + 1] = a.y; gam[8 * g + 2] = a.z; gam[8 * g + 3] = a.w;
    gam[8 * g + 4] = b.x; gam[8 * g + 5] = b.y; gam[8 * g + 6] = b.z; gam[8 * g + 7] = b.w;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) gsum[i] = bsum[i] = 0.f;
  const int64_t r0 = (static_cast<int64_t>(blockIdx.x) * nwarps + warp) * rows_per_warp;
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const int64_t row = r0 + rr;
    if (row >= rows) break;
    float xv[NV], dv[NV];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float4* xp = reinterpret_cast<const float4*>(x + row * d + g * 256 + lane * 8);
      const float4* dp = reinterpret_cast<const float4*>(dy + row * d + g * 256 + lane * 8);
      const float4 a = xp[0], b = xp[1], c = dp[0], e = dp[1];
      xv[8 * g + 0] = a.x; xv[8 * g + 1] = a.y; xv[8 * g + 2] = a.z; xv[8 * g + 3] = a.w;
      xv[8 * g + 4] = b.x; xv[8 * g + 5] = b.y; xv[8 * g + 6] = b.z; xv[8 * g + 7] = b.w;
      dv[8 * g + 0] = c.x; dv[8 * g + 1] = c.y; dv[8 * g + 2] = c.z; dv[8 * g + 3] = c.w;
      dv[8 * g + 4] = e.x; dv[8 * g + 5] = e.y; dv[8 * g + 6] = e.z; dv[8 * g + 7] = e.w;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += xv[i];
    const float mean = warp_sum(s) * (1.0f / d);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xv[i] -= mean;
      q += xv[i] * xv[i];
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / d) + 1e-5f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xv[i] *= rstd;                                            // xhat
      const float gy = dv[i] * gam[i];
      s1 += gy;
      s2 += gy * xv[i];
      gsum[i] += dv[i] * xv[i];
      bsum[i] += dv[i];
      dv[i] = gy;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {                          // both sums in one butterfly
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= (1.0f / d);
    s2 *= (1.0f / d);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float4* op = reinterpret_cast<float4*>(dx + row * d + g * 256 + lane * 8);
      float4 a, b;
      a.x = rstd * (dv[8 * g + 0] - s1 - xv[8 * g + 0] * s2); a.y = rstd * (dv[8 * g + 1] - s1 - xv[8 * g + 1] * s2);
      a.z = rstd * (dv[8 * g + 2] - s1 - xv[8 * g + 2] * s2); a.w = rstd * (dv[8 * g + 3] - s1 - xv[8 * g + 3] * s2);
      b.x = rstd * (dv[8 * g + 4] - s1 - xv[8 * g + 4] * s2); b.y = rstd * (dv[8 * g + 5] - s1 - xv[8 * g + 5] * s2);
      b.z = rstd * (dv[8 * g + 6] - s1 - xv[8 * g + 6] * s2); b.w = rstd * (dv[8 * g + 7] - s1 - xv[8 * g + 7] * s2);
      if (accumulate) {
        const float4 pa = op[0], pb = op[1];
        a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
        b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
      }
      op[0] = a;
      op[1] = b;
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&red[g * 256 + lane * 8 + e], gsum[8 * g + e]);
      atomicAdd(&red[d + g * 256 + lane * 8 + e], bsum[8 * g + e]);
    }
  __syncthreads();
  for (int i = threadIdx.x * 4; i < 2 * d; i += blockDim.x * 4) {
    float* dst = i < d ? dgamma + i : dbeta + (i - d);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(red[i]), "f"(red[i + 1]), "f"(red[i + 2]), "f"(red[i + 3])
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------- depthwise conv
// z[b,t,c] = bias[c] + sum_j w[c,j] g[b,t+j-half,c]   (Conv1d groups = channels, zero padding at the tensor edges,
// convolution.py:15,26); weights in the module's own layout [c][k].  FLIP = true computes the input gradient
// dg[b,t,c] = sum_j w[c,j] dz[b,t-j+half,c] with the same code (taps reversed, no bias).
// One block = 32 channels x 64 frames of one utterance, input rows staged in shared memory; thread = (channel, 8 frames).
template <typename TIn, bool FLIP>
__global__ void __launch_bounds__(256) dwconv_tile_kernel(const TIn* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                                          float* __restrict__ out, int Tn, int d, int ks) {
  pdl_wait();
  constexpr int TT = 64, KMAX = 64;
  __shared__ float in_s[TT + KMAX][32];
  __shared__ float w_s[KMAX][33];
  const int c0 = blockIdx.x * 32, t0 = blockIdx.y * TT, b = blockIdx.z;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int half = (ks - 1) / 2, c = c0 + lane;
  for (int r = wp; r < TT + ks - 1; r += 8) {
    const int t = t0 + r - half;
    in_s[r][lane] = (t >= 0 && t < Tn && c < d) ? to_f32(in[(static_cast<int64_t>(b) * Tn + t) * d + c]) : 0.f;
  }
  for (int i = threadIdx.x; i < 32 * ks; i += 256) {
    const int cc = i / ks, j = i - cc * ks;
    w_s[FLIP ? ks - 1 - j : j][cc] = (c0 + cc < d) ? w[(c0 + cc) * ks + j] : 0.f;
  }
  __syncthreads();
  if (c >= d) return;
  const float bc = FLIP ? 0.f : bias[c];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = wp * 8 + i, t = t0 + r;
    if (t >= Tn) break;
    float acc = bc;
    for (int j = 0; j < ks; ++j) acc = fmaf(w_s[j][lane], in_s[r + j][lane], acc);
    out[(static_cast<int64_t>(b) * Tn + t) * d + c] = acc;
  }
}
// dw[c,j] += sum_{b,t} dz[b,t,c] g[b,t+j-half,c];  db[c] += sum dz.  One block = 32 channels x 64 frames of one utterance:
// the dz rows and the g rows they touch (64 + ks - 1) are staged in shared memory, thread = (channel, taps j0, j0+8, ...)
template <typename TA>
__global__ void __launch_bounds__(256) dwconv_bwd_weight_kernel(const float* __restrict__ dz, const TA* __restrict__ g, float* __restrict__ dw,
                                                                float* __restrict__ db, int Tn, int d, int ks) {
  pdl_wait();
  constexpr int TT = 64, KMAX = 64;
  __shared__ float dz_s[TT][32];
  __shared__ float g_s[TT + KMAX][32];
  const int c0 = blockIdx.x * 32, t0 = blockIdx.y * TT, b = blockIdx.z;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int half = (ks - 1) / 2, c = c0 + lane;
  for (int r = w; r < TT; r += 8) {
    const int t = t0 + r;
    dz_s[r][lane] = (t < Tn && c < d) ? dz[(static_cast<int64_t>(b) * Tn + t) * d + c] : 0.f;
  }
  for (int r = w; r < TT + ks - 1; r += 8) {
    const int t = t0 + r - half;
    g_s[r][lane] = (t >= 0 && t < Tn && c < d) ? to_f32(g[(static_cast<int64_t>(b) * Tn + t) * d + c]) : 0.f;
  }
  __syncthreads();
  if (c >= d) return;
  for (int j = w; j < ks; j += 8) {
    float acc = 0.f;
#pragma unroll 8
    for (int r = 0; r < TT; ++r) acc = fmaf(dz_s[r][lane], g_s[r + j][lane], acc);
    atomicAdd(dw + c * ks + j, acc);
  }
  if (w == 0) {
    float bs = 0.f;
    for (int r = 0; r < TT; ++r) bs += dz_s[r][lane];
    atomicAdd(db + c, bs);
  }
}

// ------------------------------------------------------------------------------------------- BatchNorm1d, training mode
// per-channel sums over all M = B*T rows: acc[c] += sum z, acc[d + c] += sum z^2 (fp32 partials, one atomic per block)
__global__ void bn_sums_kernel(const float* __restrict__ z, float* __restrict__ acc, int M, int d, int rows_per_block) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float s = 0.f, q = 0.f;
#pragma unroll 8
  for (int m = m0; m < m1; ++m) {
    const float v = z[static_cast<int64_t>(m) * d + c];
    s += v;
    q = fmaf(v, v, q);
  }
  atomicAdd(acc + c, s);
  atomicAdd(acc + d + c, q);
}
// mean / rstd of the batch (biased variance normalises, nn.BatchNorm1d), running statistics updated in place with the
// unbiased variance and momentum 0.1
__global__ void bn_finish_kernel(const float* __restrict__ acc, float* __restrict__ stats, float* running_mean, float* running_var,
                                 int M, int d, float momentum, float eps) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const float mean = acc[c] / M;
  const float var = fmaxf(acc[d + c] / M - mean * mean, 0.f);
  stats[c] = mean;
  stats[d + c] = rsqrtf(var + eps);
  if (running_mean != nullptr) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (M > 1 ? static_cast<float>(M) / (M - 1) : 1.f);
  }
}
template <typename TA>
__global__ void bn_swish_fwd_kernel(const float* __restrict__ z, const float* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, TA* __restrict__ s, int64_t M, int d) {
  pdl_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M * d) return;
  const int c = static_cast<int>(i % d);
  s[i] = from_f32<TA>(swish_f((z[i] - stats[c]) * stats[d + c] * gamma[c] + beta[c]));
}
// first backward pass: dzn = ds * swish'(zn) written over ds; acc[c] += sum dzn, acc[d+c] += sum dzn * zhat
__global__ void bn_swish_bwd_reduce_kernel(float* ds, const float* __restrict__ z, const float* __restrict__ stats,
                                           const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ acc, int M,
                                           int d, int rows_per_block) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  const float mean = stats[c], rstd = stats[d + c], ga = gamma[c], be = beta[c];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
  for (int m = m0; m < m1; ++m) {
    const int64_t i = static_cast<int64_t>(m) * d + c;
    const float zh = (z[i] - mean) * rstd, zn = zh * ga + be, sg = sigmoid_f(zn);
    const float dzn = ds[i] * (sg + zn * sg * (1.f - sg));
    ds[i] = dzn;
    s1 += dzn;
    s2 = fmaf(dzn, zh, s2);
  }
  atomicAdd(acc + c, s1);
  atomicAdd(acc + d + c, s2);
}
// second pass: dz = gamma * rstd * (dzn - mean(dzn) - zhat * mean(dzn * zhat)), in place; block row 0 adds dgamma, dbeta
__global__ void bn_bwd_apply_kernel(float* dzn, const float* __restrict__ z, const float* __restrict__ stats,
                                    const float* __restrict__ gamma, const float* __restrict__ acc, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int M, int d, int rows_per_block) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  const float mean = stats[c], rstd = stats[d + c], k = gamma[c] * rstd, a1 = acc[c] / M, a2 = acc[d + c] / M;
#pragma unroll 8
  for (int m = m0; m < m1; ++m) {
    const int64_t i = static_cast<int64_t>(m) * d + c;
    dzn[i] = k * (dzn[i] - a1 - (z[i] - mean) * rstd * a2);
  }
  if (blockIdx.y == 0) {
    atomicAdd(dgamma + c, acc[d + c]);
    atomicAdd(dbeta + c, acc[c]);
  }
}

// ------------------------------------------------------------------------------------------- front-end backward
// colT[(kw*3+kh)*C + c][m] = out1[b][2 t2 + kw][2 f2 + kh][c], m = (b*T2 + t2)*F2 + f2: the transposed im2col operand of
// conv_2's weight gradient (K order of pack_conv2_weight); columns m >= M2 are zero
template <typename TA>
__global__ void im2col_t_kernel(const TA* __restrict__ out1, TA* __restrict__ colT, ConvGeom g, int64_t M2, int64_t M2p) {
  pdl_wait();
  __shared__ TA tile[32][33];
  const int tap = blockIdx.z, kw = tap / 3, kh = tap - kw * 3;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {                            // rows m, columns c (coalesced along c)
    const int64_t m = m0 + r;
    TA v = from_f32<TA>(0.f);
    if (m < M2 && c0 + tx < g.C) {
      const int f2 = static_cast<int>(m % g.F2), t2 = static_cast<int>((m / g.F2) % g.T2), b = static_cast<int>(m / (static_cast<int64_t>(g.F2) * g.T2));
      v = out1[((static_cast<int64_t>(b) * g.T1p + 2 * t2 + kw) * g.F1 + 2 * f2 + kh) * g.C + c0 + tx];
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r;
    const int64_t m = m0 + tx;
    if (c < g.C && m < M2p) colT[(static_cast<int64_t>(tap) * g.C + c) * M2p + m] = tile[tx][r];
  }
}
// col[m][(kw*3+kh)*C + c] = out1[b][2 t2 + kw][2 f2 + kh][c], m = (b*T2 + t2)*F2 + f2: the im2col operand of conv_2's weight
// gradient in the row-major form the TN GEMM reads in place (8 channels = 16 bytes per thread, fully coalesced)
__global__ void im2col_rows_kernel(const bf16* __restrict__ out1, bf16* __restrict__ col, ConvGeom g, int64_t M2) {
  pdl_wait();
  const int C8 = g.C / 8;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M2 * 9 * C8) return;
  const int c8 = static_cast<int>(i % C8), tap = static_cast<int>((i / C8) % 9);
  const int64_t m = i / (9 * static_cast<int64_t>(C8));
  const int kw = tap / 3, kh = tap - kw * 3;
  const int f2 = static_cast<int>(m % g.F2), t2 = static_cast<int>((m / g.F2) % g.T2), b = static_cast<int>(m / (static_cast<int64_t>(g.F2) * g.T2));
  const uint4 v = *reinterpret_cast<const uint4*>(out1 + ((static_cast<int64_t>(b) * g.T1p + 2 * t2 + kw) * g.F1 + 2 * f2 + kh) * g.C + c8 * 8);
  *reinterpret_cast<uint4*>(col + (m * 9 + tap) * g.C + c8 * 8) = v;
}

// dOut1[b,t1,f1,c] = relu'(out1) * sum over the taps that read this input of dcol[m(t2,f2)][(kw*3+kh)*C + c]
template <typename TA>
__global__ void col2im_relu_kernel(const TA* __restrict__ dcol, const TA* __restrict__ out1, float* __restrict__ dout1, ConvGeom g) {
  pdl_wait();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(g.B) * g.T1p * g.F1 * g.C;
  if (i >= total) return;
  const int c = static_cast<int>(i % g.C), f1 = static_cast<int>((i / g.C) % g.F1), t1 = static_cast<int>((i / (static_cast<int64_t>(g.C) * g.F1)) % g.T1p);
  const int b = static_cast<int>(i / (static_cast<int64_t>(g.C) * g.F1 * g.T1p));
  float acc = 0.f;
  if (to_f32(out1[i]) > 0.f) {
    for (int kw = 0; kw < 3; ++kw) {
      const int tt = t1 - kw;
      if (tt < 0 || (tt & 1) || (tt >> 1) >= g.T2) continue;
      for (int kh = 0; kh < 3; ++kh) {
        const int ff = f1 - kh;
        if (ff < 0 || (ff & 1) || (ff >> 1) >= g.F2) continue;
        const int64_t m = (static_cast<int64_t>(b) * g.T2 + (tt >> 1)) * g.F2 + (ff >> 1);
        acc += to_f32(dcol[m * 9 * g.C + (kw * 3 + kh) * g.C + c]);
      }
    }
  }
  dout1[i] = acc;
}
// Fused front-end tail of the backward: dOut1 = relu'(out1) * col2im(dcol) is never written - each thread gathers it for
// 8 channels of one (b, t1, f1) position (<= 4 taps of dcol, 16-byte loads) and immediately accumulates conv_1's
// parameter gradients  dw[c][kh*3+kw] += dOut1 * mel[b][2 f1 + kh][2 t1 + kw],  db[c] += dOut1  in registers; a block
// walks positions grid-stride, reduces in shared memory and adds once to global memory.
template <typename TA>
__global__ void __launch_bounds__(256) conv1_wgrad_fused_kernel(const TA* __restrict__ dcol, const TA* __restrict__ out1,
                                                               const float* __restrict__ mel, float* __restrict__ dw,
                                                               float* __restrict__ db, ConvGeom g, int n_mel, int Tin, int T1, int py) {
  pdl_wait();
  extern __shared__ float red[];                                // [C][10]
  const int cg = g.C >> 3;                                      // groups of 8 channels
  const int c8 = threadIdx.x % cg, pl = threadIdx.x / cg;
  for (int i = threadIdx.x; i < g.C * 10; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float acc[9][8], accb[8];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) accb[e] = 0.f;
  const int64_t npos = static_cast<int64_t>(g.B) * T1 * g.F1;
  if (pl < py) {
    for (int64_t pos = static_cast<int64_t>(blockIdx.x) * py + pl; pos < npos; pos += static_cast<int64_t>(gridDim.x) * py) {
      const int f1 = static_cast<int>(pos % g.F1), t1 = static_cast<int>((pos / g.F1) % T1), b = static_cast<int>(pos / (static_cast<int64_t>(g.F1) * T1));
      float o[8], gsum[8];
      ld8(out1 + ((static_cast<int64_t>(b) * g.T1p + t1) * g.F1 + f1) * g.C + c8 * 8, o);
#pragma unroll
      for (int e = 0; e < 8; ++e) gsum[e] = 0.f;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int tt = t1 - kw;
        if (tt < 0 || (tt & 1) || (tt >> 1) >= g.T2) continue;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const int ff = f1 - kh;
          if (ff < 0 || (ff & 1) || (ff >> 1) >= g.F2) continue;
          const int64_t m = (static_cast<int64_t>(b) * g.T2 + (tt >> 1)) * g.F2 + (ff >> 1);
          float v[8];
          ld8(dcol + (m * 9 + (kw * 3 + kh)) * g.C + c8 * 8, v);
#pragma unroll
          for (int e = 0; e < 8; ++e) gsum[e] += v[e];
        }
      }
      bool any = false;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        gsum[e] = o[e] > 0.f ? gsum[e] : 0.f;
        any |= gsum[e] != 0.f;
      }
      if (!any) continue;
      const float* mrow = mel + (static_cast<int64_t>(b) * n_mel + 2 * f1) * Tin + 2 * t1;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float mv = __ldg(mrow + static_cast<int64_t>(kh) * Tin + kw);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[kh * 3 + kw][e] = fmaf(gsum[e], mv, acc[kh * 3 + kw][e]);
        }
#pragma unroll
      for (int e = 0; e < 8; ++e) accb[e] += gsum[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float* r = red + (c8 * 8 + e) * 10;
#pragma unroll
      for (int k = 0; k < 9; ++k) atomicAdd(r + k, acc[k][e]);
      atomicAdd(r + 9, accb[e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < g.C * 10; i += blockDim.x) {
    const int c = i / 10, k = i - c * 10;
    if (k < 9) atomicAdd(dw + c * 9 + k, red[i]);
    else atomicAdd(db + c, red[i]);
  }
}

// conv_1 parameter gradients: dw[c][kh*3+kw] += sum_{b,t1,f1} dout1[b,t1,f1,c] mel[b][2 f1 + kh][2 t1 + kw]; db[c] += sum dout1.
// block = (32 channels) x (a slab of t1 rows of one utterance)
__global__ void conv1_wgrad_kernel(const float* __restrict__ dout1, const float* __restrict__ mel, float* __restrict__ dw, float* __restrict__ db,
                                   int n_mel, int Tin, int T1, int T1p, int F1, int C, int t_per_block) {
  pdl_wait();
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), b = blockIdx.z;
  const int sub = threadIdx.x >> 5, nsub = blockDim.x >> 5;
  const int t0 = blockIdx.y * t_per_block, t1e = min(T1, t0 + t_per_block);
  float acc[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) acc[k] = 0.f;
  if (c < C) {
    for (int t1 = t0 + sub; t1 < t1e; t1 += nsub)
      for (int f1 = 0; f1 < F1; ++f1) {
        const float gsc = dout1[((static_cast<int64_t>(b) * T1p + t1) * F1 + f1) * C + c];
        if (gsc == 0.f) continue;
        acc[9] += gsc;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
            acc[kh * 3 + kw] = fmaf(gsc, mel[(static_cast<int64_t>(b) * n_mel + 2 * f1 + kh) * Tin + 2 * t1 + kw], acc[kh * 3 + kw]);
      }
  }
  __shared__ float red[8][32][10];
#pragma unroll
  for (int k = 0; k < 10; ++k) red[sub][threadIdx.x & 31][k] = acc[k];
  __syncthreads();
  if (sub == 0 && c < C) {
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      float t = 0.f;
      for (int s = 0; s < nsub; ++s) t += red[s][threadIdx.x & 31][k];
      if (k < 9) atomicAdd(dw + c * 9 + k, t);
      else atomicAdd(db + c, t);
    }
  }
}

}  // namespace

#define LAUNCH_OK()                    \
  do {                                 \
    CFM_CUDA_OK(cudaGetLastError());   \
    return 0;                          \
  } while (0)

template <typename TOut>
int cast_scale_drop(const float* in, TOut* out, int64_t n, float scale, Drop dr, cudaStream_t st) {
  CFM_REQUIRE(n % 4 == 0, "element count must be a multiple of 4");
  CFM_CUDA_OK(launch_pdl(cast_scale_drop_kernel<TOut>, dim3(nblk(n, 1024)), dim3(256), 0, st, 1, in, out, n, scale, dr));
  LAUNCH_OK();
}
template <typename TA>
int swish_drop_fwd(const TA* hp, TA* out, int64_t n, Drop dr, cudaStream_t st) {
  if constexpr (sizeof(TA) == 2) {
    if (n % 8 == 0 && ((reinterpret_cast<uintptr_t>(hp) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
      CFM_CUDA_OK(launch_pdl(swish_drop_fwd8_kernel, dim3(nblk(n, 2048)), dim3(256), 0, st, 1, reinterpret_cast<const bf16*>(hp), reinterpret_cast<bf16*>(out), n, dr));
      LAUNCH_OK();
    }
  }
  CFM_CUDA_OK(launch_pdl(swish_drop_fwd_kernel<TA>, dim3(nblk(n, 1024)), dim3(256), 0, st, 1, hp, out, n, dr));
  LAUNCH_OK();
}
template <typename TA>
int swish_drop_bwd(const float* dh, const TA* hp, TA* out, int64_t n, Drop dr, cudaStream_t st) {
  CFM_CUDA_OK(launch_pdl(swish_drop_bwd_kernel<TA>, dim3(nblk(n, 1024)), dim3(256), 0, st, 1, dh, hp, out, n, dr));
  LAUNCH_OK();
}
int residual_drop_add(const float* x, const float* t, float* out, int64_t n, float scale, Drop dr, cudaStream_t st) {
  CFM_CUDA_OK(launch_pdl(residual_drop_add_kernel, dim3(nblk(n, 1024)), dim3(256), 0, st, 1, x, t, out, n, scale, dr));
  LAUNCH_OK();
}
int drop_inplace(float* x, int64_t n, Drop dr, cudaStream_t st) {
  if (dr.p <= 0.f) return 0;
  CFM_CUDA_OK(launch_pdl(drop_inplace_kernel, dim3(nblk(n, 256)), dim3(256), 0, st, 1, x, n, dr));
  LAUNCH_OK();
}
template <typename TA>
int glu_fwd(const TA* u, TA* g, int64_t rows, int d, cudaStream_t st) {
  CFM_CUDA_OK(launch_pdl(glu_fwd_kernel<TA>, dim3(nblk(rows * d, 256)), dim3(256), 0, st, 1, u, g, rows, d));
  LAUNCH_OK();
}
template <typename TA>
int glu_bwd(const float* dg, const TA* u, TA* du, int64_t rows, int d, cudaStream_t st) {
  CFM_CUDA_OK(launch_pdl(glu_bwd_kernel<TA>, dim3(nblk(rows * d, 256)), dim3(256), 0, st, 1, dg, u, du, rows, d));
  LAUNCH_OK();
}
template <typename TA>
int relu_bwd(const float* dy, const TA* y, TA* out, int64_t n, cudaStream_t st) {
  CFM_CUDA_OK(launch_pdl(relu_bwd_kernel<TA>, dim3(nblk(n, 256)), dim3(256), 0, st, 1, dy, y, out, n));
  LAUNCH_OK();
}
template <typename T>
int transpose_pad(const T* in, int ld, T* out, int M, int N, int Mp, cudaStream_t st) {
  dim3 grid(nblk(Mp, 32), nblk(N, 32));
  CFM_CUDA_OK(launch_pdl(transpose_kernel<T>, dim3(grid), dim3(256), 0, st, 1, in, ld, out, M, N, Mp));
  LAUNCH_OK();
}
template <typename T>
int colsum_add(const T* in, int ld, float* out, int M, int N, cudaStream_t st) {
  const int rpb = 32;
  dim3 grid(nblk(N, 128), nblk(M, rpb));
  CFM_CUDA_OK(launch_pdl(colsum_kernel<T>, dim3(grid), dim3(128), 0, st, 1, in, ld, out, M, N, rpb));
  LAUNCH_OK();
}
int ln_bwd(const float* dy, const float* x, const float* gamma, float* dx, bool accumulate, float* dgamma, float* dbeta, int64_t rows,
           int d, cudaStream_t st) {
  CFM_REQUIRE(d <= 1024, "LayerNorm backward supports d_model <= 1024");
  // about two blocks per SM: every block ends with 2*d atomics on the same addresses, which serialise in the L2
  const int warps = 8;
  int rpw = static_cast<int>((rows + 296 * warps - 1) / (296 * warps));
  if (rpw < 2) rpw = 2;
  const unsigned grid = nblk(rows, rpw * warps);
  const bool aligned = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) |
                         reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dgamma) | reinterpret_cast<uintptr_t>(dbeta)) & 15) == 0;
  if (aligned && d == 256) {
    CFM_CUDA_OK(launch_pdl(ln_bwd_vec_kernel<1>, dim3(grid), dim3(warps * 32), 0, st, 1, dy, x, gamma, dx, accumulate ? 1 : 0, dgamma, dbeta, rows, rpw));
    LAUNCH_OK();
  }
  if (aligned && d == 512) {
    CFM_CUDA_OK(launch_pdl(ln_bwd_vec_kernel<2>, dim3(grid), dim3(warps * 32), 0, st, 1, dy, x, gamma, dx, accumulate ? 1 : 0, dgamma, dbeta, rows, rpw));
    LAUNCH_OK();
  }
  const size_t smem = static_cast<size_t>(2) * d * 4;
  const int nv = (d + 31) / 32;
  if (nv <= 8) CFM_CUDA_OK(launch_pdl(ln_bwd_kernel<8>, dim3(grid), dim3(warps * 32), smem, st, 1, dy, x, gamma, dx, accumulate ? 1 : 0, dgamma, dbeta, rows, d, rpw));
  else if (nv <= 16) CFM_CUDA_OK(launch_pdl(ln_bwd_kernel<16>, dim3(grid), dim3(warps * 32), smem, st, 1, dy, x, gamma, dx, accumulate ? 1 : 0, dgamma, dbeta, rows, d, rpw));
  else CFM_CUDA_OK(launch_pdl(ln_bwd_kernel<32>, dim3(grid), dim3(warps * 32), smem, st, 1, dy, x, gamma, dx, accumulate ? 1 : 0, dgamma, dbeta, rows, d, rpw));
  LAUNCH_OK();
}
template <typename TA>
int dwconv_fwd(const TA* g, const float* w, const float* bias, float* z, int B, int Tn, int d, int ks, cudaStream_t st) {
  CFM_REQUIRE(ks <= 64, "depthwise kernel size up to 64");
  dim3 grid(nblk(d, 32), nblk(Tn, 64), B);
  CFM_CUDA_OK(launch_pdl(dwconv_tile_kernel<TA, false>, dim3(grid), dim3(256), 0, st, 1, g, w, bias, z, Tn, d, ks));
  LAUNCH_OK();
}
int dwconv_bwd_input(const float* dz, const float* w, float* dg, int B, int Tn, int d, int ks, cudaStream_t st) {
  CFM_REQUIRE(ks <= 64, "depthwise kernel size up to 64");
  dim3 grid(nblk(d, 32), nblk(Tn, 64), B);
  CFM_CUDA_OK(launch_pdl(dwconv_tile_kernel<float, true>, dim3(grid), dim3(256), 0, st, 1, dz, w, nullptr, dg, Tn, d, ks));
  LAUNCH_OK();
}
template <typename TA>
int dwconv_bwd_weight(const float* dz, const TA* g, float* dw, float* db, int B, int Tn, int d, int ks, cudaStream_t st) {
  CFM_REQUIRE(ks <= 64, "depthwise kernel size up to 64");
  dim3 grid(nblk(d, 32), nblk(Tn, 64), B);
  CFM_CUDA_OK(launch_pdl(dwconv_bwd_weight_kernel<TA>, dim3(grid), dim3(256), 0, st, 1, dz, g, dw, db, Tn, d, ks));
  LAUNCH_OK();
}
int bn_train_stats(const float* z, float* acc2d, float* stats2d, float* running_mean, float* running_var, int M, int d, float momentum,
                   cudaStream_t st) {
  CFM_CUDA_OK(cudaMemsetAsync(acc2d, 0, static_cast<size_t>(2) * d * 4, st));
  const int rpb = 32;
  dim3 grid(nblk(d, 128), nblk(M, rpb));
  CFM_CUDA_OK(launch_pdl(bn_sums_kernel, dim3(grid), dim3(128), 0, st, 1, z, acc2d, M, d, rpb));
  CFM_CUDA_OK(launch_pdl(bn_finish_kernel, dim3(nblk(d, 128)), dim3(128), 0, st, 1, acc2d, stats2d, running_mean, running_var, M, d, momentum, 1e-5f));
  LAUNCH_OK();
}
template <typename TA>
int bn_swish_fwd(const float* z, const float* stats, const float* gamma, const float* beta, TA* s, int64_t M, int d, cudaStream_t st) {
  CFM_CUDA_OK(launch_pdl(bn_swish_fwd_kernel<TA>, dim3(nblk(M * d, 256)), dim3(256), 0, st, 1, z, stats, gamma, beta, s, M, d));
  LAUNCH_OK();
}
int bn_swish_bwd(float* ds_to_dz, const float* z, const float* stats, const float* gamma, const float* beta, float* acc2d, float* dgamma,
                 float* dbeta, int M, int d, cudaStream_t st) {
  CFM_CUDA_OK(cudaMemsetAsync(acc2d, 0, static_cast<size_t>(2) * d * 4, st));
  const int rpb = 32;                                           // ~750 blocks at C4: enough loads in flight to reach the HBM rate
  dim3 grid(nblk(d, 128), nblk(M, rpb));
  CFM_CUDA_OK(launch_pdl(bn_swish_bwd_reduce_kernel, dim3(grid), dim3(128), 0, st, 1, ds_to_dz, z, stats, gamma, beta, acc2d, M, d, rpb));
  CFM_CUDA_OK(launch_pdl(bn_bwd_apply_kernel, dim3(grid), dim3(128), 0, st, 1, ds_to_dz, z, stats, gamma, acc2d, dgamma, dbeta, M, d, rpb));
  LAUNCH_OK();
}
template <typename TA>
int im2col_t(const TA* out1, TA* colT, const ConvGeom& g, int64_t M2, int64_t M2p, cudaStream_t st) {
  dim3 grid(nblk(M2p, 32), nblk(g.C, 32), 9);
  CFM_CUDA_OK(launch_pdl(im2col_t_kernel<TA>, dim3(grid), dim3(256), 0, st, 1, out1, colT, g, M2, M2p));
  LAUNCH_OK();
}
int im2col_rows(const bf16* out1, bf16* col, const ConvGeom& g, int64_t M2, cudaStream_t st) {
  CFM_REQUIRE(g.C % 8 == 0, "im2col needs channels to be a multiple of 8");
  const int64_t n = M2 * 9 * (g.C / 8);
  CFM_CUDA_OK(launch_pdl(im2col_rows_kernel, dim3(nblk(n, 256)), dim3(256), 0, st, 1, out1, col, g, M2));
  LAUNCH_OK();
}
template <typename TA>
int col2im_relu(const TA* dcol, const TA* out1, float* dout1, const ConvGeom& g, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(g.B) * g.T1p * g.F1 * g.C;
  CFM_CUDA_OK(launch_pdl(col2im_relu_kernel<TA>, dim3(nblk(total, 256)), dim3(256), 0, st, 1, dcol, out1, dout1, g));
  LAUNCH_OK();
}
int conv1_wgrad(const float* dout1, const float* mel, float* dw, float* db, int B, int n_mel, int Tin, int T1, int T1p, int F1, int C,
                cudaStream_t st) {
  const int tpb = 16;
  dim3 grid(nblk(C, 32), nblk(T1, tpb), B);
  CFM_CUDA_OK(launch_pdl(conv1_wgrad_kernel, dim3(grid), dim3(256), 0, st, 1, dout1, mel, dw, db, n_mel, Tin, T1, T1p, F1, C, tpb));
  LAUNCH_OK();
}

template <typename TA>
int conv1_wgrad_fused(const TA* dcol, const TA* out1, const float* mel, float* dw, float* db, const ConvGeom& g, int n_mel, int Tin, int T1,
                      int num_sms, cudaStream_t st) {
  CFM_REQUIRE(g.C % 8 == 0 && g.C / 8 <= 256, "fused conv_1 weight gradient needs a channel count that is a multiple of 8 (<= 2048)");
  const int cg = g.C / 8, py = 256 / cg;
  CFM_CUDA_OK(launch_pdl(conv1_wgrad_fused_kernel<TA>, dim3(2 * num_sms), dim3(cg * py), static_cast<size_t>(g.C) * 10 * 4, st, 1, dcol, out1, mel, dw, db, g, n_mel, Tin, T1, py));
  LAUNCH_OK();
}

#define INST_TA(T)                                                                                                   \
  template int swish_drop_fwd<T>(const T*, T*, int64_t, Drop, cudaStream_t);                                          \
  template int swish_drop_bwd<T>(const float*, const T*, T*, int64_t, Drop, cudaStream_t);                            \
  template int glu_fwd<T>(const T*, T*, int64_t, int, cudaStream_t);                                                  \
  template int glu_bwd<T>(const float*, const T*, T*, int64_t, int, cudaStream_t);                                    \
  template int relu_bwd<T>(const float*, const T*, T*, int64_t, cudaStream_t);                                        \
  template int transpose_pad<T>(const T*, int, T*, int, int, int, cudaStream_t);                                      \
  template int colsum_add<T>(const T*, int, float*, int, int, cudaStream_t);                                          \
  template int dwconv_fwd<T>(const T*, const float*, const float*, float*, int, int, int, int, cudaStream_t);         \
  template int dwconv_bwd_weight<T>(const float*, const T*, float*, float*, int, int, int, int, cudaStream_t);        \
  template int bn_swish_fwd<T>(const float*, const float*, const float*, const float*, T*, int64_t, int, cudaStream_t); \
  template int im2col_t<T>(const T*, T*, const ConvGeom&, int64_t, int64_t, cudaStream_t);                            \
  template int col2im_relu<T>(const T*, const T*, float*, const ConvGeom&, cudaStream_t);                             \
  template int conv1_wgrad_fused<T>(const T*, const T*, const float*, float*, float*, const ConvGeom&, int, int, int, int, cudaStream_t); \
  template int cast_scale_drop<T>(const float*, T*, int64_t, float, Drop, cudaStream_t);
INST_TA(float)
INST_TA(bf16)

}  // namespace cfm
