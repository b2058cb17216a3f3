// LayerNorm, positional table, length/mask arithmetic, casts and weight packing kernels.
#include "ops.h"

namespace cfm {
namespace {

// One warp per row; two-pass mean/variance in fp32 (rows are <= 4 KB and stay in L1).
// (x and y may alias - the block-final LayerNorm runs in place - so neither is __restrict__.)
template <typename TOut>
__global__ void __launch_bounds__(256) layer_norm_kernel(const float* x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, TOut* y, int64_t rows, int d) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * d;
  float s = 0.f;
  for (int c = lane; c < d; c += 32) s += xr[c];
  const float mean = warp_sum(s) / static_cast<float>(d);
  float v = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float t = xr[c] - mean;
    v = fmaf(t, t, v);
  }
  const float rstd = rsqrtf(warp_sum(v) / static_cast<float>(d) + 1e-5f);
  TOut* yr = y + row * d;
  for (int c = lane; c < d; c += 32) yr[c] = from_f32<TOut>((xr[c] - mean) * rstd * gamma[c] + beta[c]);
}

// Vectorised variant for d = 128 * NV: each lane keeps its 4*NV values in registers (one pass over HBM),
// loads them as float4 (coalesced 512 B per warp request) and stores 8-byte (bf16) or 16-byte (fp32) vectors.
template <typename TOut, int NV>
__global__ void __launch_bounds__(256) layer_norm_vec_kernel(const float* x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, TOut* y, int64_t rows) {
  pdl_wait();
  constexpr int D = 128 * NV;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    const float o0 = v[i].x * rstd * g4.x + b4.x, o1 = v[i].y * rstd * g4.y + b4.y;
    const float o2 = v[i].z * rstd * g4.z + b4.z, o3 = v[i].w * rstd * g4.w + b4.w;
    if (sizeof(TOut) == 4) {
      reinterpret_cast<float4*>(y + row * D)[i * 32 + lane] = make_float4(o0, o1, o2, o3);
    } else {
      reinterpret_cast<uint2*>(y + row * D)[i * 32 + lane] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
    }
  }
}

// Same, RPW rows per warp with every row's loads issued before the first reduction: 4 x 1 KB in flight per warp keeps
// the HBM pipe fuller than one row at a time (measured on B200 at [23936, 256]: 8.4 -> see DESIGN.md).
template <typename TOut, int NV, int RPW>
__global__ void __launch_bounds__(256) layer_norm_vec_rows_kernel(const float* x, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, TOut* y, int64_t rows) {
  constexpr int D = 128 * NV;
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  if (row0 >= rows) return;
  float4 v[RPW][NV];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const bool ok = row0 + r < rows;
    const float4* xr = reinterpret_cast<const float4*>(x + (row0 + (ok ? r : 0)) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[r][i] = xr[i * 32 + lane];
  }
  float4 g4[NV], b4[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g4[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    b4[i] = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    if (row0 + r >= rows) break;                                 // warp-uniform
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[r][i].x -= mean; v[r][i].y -= mean; v[r][i].z -= mean; v[r][i].w -= mean;
      q += (v[r][i].x * v[r][i].x + v[r][i].y * v[r][i].y) + (v[r][i].z * v[r][i].z + v[r][i].w * v[r][i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float o0 = v[r][i].x * rstd * g4[i].x + b4[i].x, o1 = v[r][i].y * rstd * g4[i].y + b4[i].y;
      const float o2 = v[r][i].z * rstd * g4[i].z + b4[i].z, o3 = v[r][i].w * rstd * g4[i].w + b4[i].w;
      if (sizeof(TOut) == 4) {
        reinterpret_cast<float4*>(y + (row0 + r) * D)[i * 32 + lane] = make_float4(o0, o1, o2, o3);
      } else {
        reinterpret_cast<uint2*>(y + (row0 + r) * D)[i * 32 + lane] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
      }
    }
  }
}

// Two LayerNorms in one pass over a row of d = 256:  y = LN(x; g1, b1) (fp32: the block output, block.py:27) and
// a = LN(y; g2, b2) (bf16: the normalised input of the NEXT block's first half-step FFN, ffn.py:16).  The fused FFN
// kernel then starts from one TMA load of `a` instead of every CTA pulling 128 KB of fp32 rows and normalising them.
__global__ void __launch_bounds__(256) layer_norm_dual_kernel(const float* x, const float* __restrict__ g1,
                                                              const float* __restrict__ b1, float* y,
                                                              const float* __restrict__ g2, const float* __restrict__ b2,
                                                              bf16* __restrict__ a, int64_t rows) {
  constexpr int D = 256, RPW = 4;                                  // rows per warp: all rows' loads are in flight together
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
  if (row0 >= rows) return;
  float v[RPW][8];                                                 // lane owns columns 8*lane .. 8*lane+7
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    if (row0 + r < rows) {
      const float4* xr = reinterpret_cast<const float4*>(x + (row0 + r) * D) + 2 * lane;
      const float4 p0 = xr[0], p1 = xr[1];
      v[r][0] = p0.x; v[r][1] = p0.y; v[r][2] = p0.z; v[r][3] = p0.w; v[r][4] = p1.x; v[r][5] = p1.y; v[r][6] = p1.z; v[r][7] = p1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[r][i] = 0.f;
    }
  }
  float gg[2][8], be[2][8];
  {
    const float* gs[2] = {g1, g2};
    const float* bs[2] = {b1, b2};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float4 ga = __ldg(reinterpret_cast<const float4*>(gs[k]) + 2 * lane), gb = __ldg(reinterpret_cast<const float4*>(gs[k]) + 2 * lane + 1);
      const float4 ba = __ldg(reinterpret_cast<const float4*>(bs[k]) + 2 * lane), bb = __ldg(reinterpret_cast<const float4*>(bs[k]) + 2 * lane + 1);
      gg[k][0] = ga.x; gg[k][1] = ga.y; gg[k][2] = ga.z; gg[k][3] = ga.w; gg[k][4] = gb.x; gg[k][5] = gb.y; gg[k][6] = gb.z; gg[k][7] = gb.w;
      be[k][0] = ba.x; be[k][1] = ba.y; be[k][2] = ba.z; be[k][3] = ba.w; be[k][4] = bb.x; be[k][5] = bb.y; be[k][6] = bb.z; be[k][7] = bb.w;
    }
  }
  auto normalise = [&](float (&w)[8], int k) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += w[i];
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] -= mean; q += w[i] * w[i]; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = w[i] * rstd * gg[k][i] + be[k][i];
  };
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    if (row0 + r >= rows) break;                                   // warp-uniform
    normalise(v[r], 0);
    float4* yr = reinterpret_cast<float4*>(y + (row0 + r) * D) + 2 * lane;
    yr[0] = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
    yr[1] = make_float4(v[r][4], v[r][5], v[r][6], v[r][7]);
    normalise(v[r], 1);
    reinterpret_cast<uint4*>(a + (row0 + r) * D)[lane] = make_uint4(pack_bf16x2(v[r][0], v[r][1]), pack_bf16x2(v[r][2], v[r][3]),
                                                                    pack_bf16x2(v[r][4], v[r][5]), pack_bf16x2(v[r][6], v[r][7]));
  }
}

template <typename T>
__global__ void rel_pos_table_kernel(const float* __restrict__ div_term, T* __restrict__ pe, int Tn, int d) {
  const int half = d >> 1;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(2 * Tn - 1) * half;
  if (idx >= total) return;
  const int p = static_cast<int>(idx / half), k = static_cast<int>(idx - static_cast<int64_t>(p) * half);
  // reference: angles = position * div_term in fp32, sin/cos of +-angle (position.py:15-21)
  const float pos = static_cast<float>(Tn - 1 - p);
  const float ang = pos * div_term[k];
  pe[static_cast<int64_t>(p) * d + 2 * k] = from_f32<T>(sinf(ang));
  pe[static_cast<int64_t>(p) * d + 2 * k + 1] = from_f32<T>(cosf(ang));
}

__global__ void lengths_mask_kernel(const int64_t* __restrict__ len_in, int64_t* __restrict__ len_out,
                                    uint8_t* __restrict__ key_pad, int B, int Tn, bool subsample,
                                    bool mark_valid) {
  pdl_wait();
  const int b = blockIdx.x;
  int64_t L = len_in[b];
  if (subsample) {
    // floor division like torch's // on int64 (convolution.py:55)
    auto fdiv2 = [](int64_t v) { return (v >= 0) ? (v >> 1) : -((-v + 1) >> 1); };
    L = fdiv2(fdiv2(L - 1) - 1);
  }
  if (threadIdx.x == 0 && len_out) len_out[b] = L;
  if (key_pad)
    for (int t = threadIdx.x; t < Tn; t += blockDim.x) key_pad[static_cast<int64_t>(b) * Tn + t] = ((t >= L) != mark_valid) ? 1 : 0;
}

// A sample whose keys are ALL masked: the reference fills every score with finfo.min, so its softmax is uniform over
// all T keys (attention.py:60-66) and the context of every query is the mean of V.  The attention kernels give masked
// keys weight exactly 0 (right whenever one key is valid), so such a sample is patched here: one block per sample, which
// returns at once when it finds a valid key.
template <typename T>
__global__ void __launch_bounds__(256) all_masked_mean_v_kernel(const uint8_t* __restrict__ key_pad, const T* __restrict__ qkv,
                                                                T* __restrict__ ctx, int Tn, int d) {
  const int b = blockIdx.x;
  int valid = 0;
  for (int t = threadIdx.x; t < Tn; t += blockDim.x) valid |= key_pad[static_cast<int64_t>(b) * Tn + t] == 0;
  if (__syncthreads_or(valid)) return;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < Tn; ++t) s += to_f32(qkv[(static_cast<int64_t>(b) * Tn + t) * 3 * d + 2 * d + c]);
    const T m = from_f32<T>(s / static_cast<float>(Tn));
    for (int t = 0; t < Tn; ++t) ctx[(static_cast<int64_t>(b) * Tn + t) * d + c] = m;
  }
}

__global__ void swish_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = swish_f(x[i]);
}
__global__ void glu_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int d) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  const int64_t r = i / d;
  const int c = static_cast<int>(i - r * d);
  y[i] = x[r * 2 * d + c] * sigmoid_f(x[r * 2 * d + d + c]);
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16_rn(x[i]);
}
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __bfloat162float(x[i]);
}
__global__ void pack_conv2_kernel(const float* __restrict__ w, float* __restrict__ out, int C) {
  // w[n][c][kh][kw] -> out[n][(kw*3+kh)*C + c]
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(C) * C * 9;
  if (i >= total) return;
  const int kw = static_cast<int>(i % 3), kh = static_cast<int>((i / 3) % 3);
  const int c = static_cast<int>((i / 9) % C), n = static_cast<int>(i / (9 * static_cast<int64_t>(C)));
  out[static_cast<int64_t>(n) * 9 * C + static_cast<int64_t>(kw * 3 + kh) * C + c] = w[i];
}
__global__ void pack_linear_kernel(const float* __restrict__ w, float* __restrict__ out, int d, int C, int F2) {
  // w[n][c*F2+f] -> out[n][f*C+c]   (flatten order of convolution.py:51-52)
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t K = static_cast<int64_t>(C) * F2;
  if (i >= K * d) return;
  const int64_t n = i / K, r = i - n * K;
  const int c = static_cast<int>(r / F2), f = static_cast<int>(r - static_cast<int64_t>(c) * F2);
  out[n * K + static_cast<int64_t>(f) * C + c] = w[i];
}
__global__ void fold_dwconv_bn_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                      const float* __restrict__ bn_w, const float* __restrict__ bn_b,
                                      const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                      float* __restrict__ w_out, float* __restrict__ b_out, int d, int k) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  // BatchNorm1d in eval mode is an affine map per channel (convolution.py:27)
  const float s = bn_w[c] / sqrtf(rv[c] + eps);
  for (int j = 0; j < k; ++j) w_out[static_cast<int64_t>(j) * d + c] = w[static_cast<int64_t>(c) * k + j] * s;
  b_out[c] = (b[c] - rm[c]) * s + bn_b[c];
}

__global__ void pack_glu_rows_kernel(const float* __restrict__ w, float* __restrict__ out, int d, int K) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= 2LL * d * K) return;
  const int r = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<int64_t>(r) * K);
  const int grp = r >> 6, in_grp = r & 63;                   // packed row r = 64*grp + (gate ? 32 : 0) + j
  const int c = grp * 32 + (in_grp & 31);
  const int src = (in_grp >= 32 ? d : 0) + c;
  out[i] = w[static_cast<int64_t>(src) * K + k];
}

inline unsigned blocks_for(int64_t n, int t) { return static_cast<unsigned>((n + t - 1) / t); }

}  // namespace

template <typename TOut>
int layer_norm(const float* x, const float* gamma, const float* beta, TOut* y, int64_t rows, int d,
               cudaStream_t stream) {
  CFM_REQUIRE(rows > 0 && d > 0, "empty layer_norm");
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                         reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  if (aligned && d == 256 && rows >= 4096)
    CFM_CUDA_OK(launch_pdl(layer_norm_vec_rows_kernel<TOut, 2, 4>, dim3(blocks_for(rows, 32)), dim3(256), 0, stream, 1, x, gamma, beta, y,
                           rows));
  else if (aligned && d == 256)
    CFM_CUDA_OK(launch_pdl<sizeof(TOut) == 2>(layer_norm_vec_kernel<TOut, 2>, dim3(blocks_for(rows, 8)), dim3(256), 0, stream, 1, x, gamma, beta,
                                             y, rows));
  else if (aligned && d == 512)
    CFM_CUDA_OK(launch_pdl<sizeof(TOut) == 2>(layer_norm_vec_kernel<TOut, 4>, dim3(blocks_for(rows, 8)), dim3(256), 0, stream, 1, x, gamma, beta,
                                             y, rows));
  else if (aligned && d == 128)
    CFM_CUDA_OK(launch_pdl<sizeof(TOut) == 2>(layer_norm_vec_kernel<TOut, 1>, dim3(blocks_for(rows, 8)), dim3(256), 0, stream, 1, x, gamma, beta,
                                             y, rows));
  else
    CFM_CUDA_OK(launch_pdl<sizeof(TOut) == 2>(layer_norm_kernel<TOut>, dim3(blocks_for(rows, 8)), dim3(256), 0, stream, 1, x, gamma, beta, y, rows, d));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
int layer_norm_dual(const float* x, const float* g1, const float* b1, float* y, const float* g2, const float* b2, bf16* a,
                    int64_t rows, int d, cudaStream_t stream) {
  CFM_REQUIRE(d == 256, "layer_norm_dual is specialised for d_model = 256");
  CFM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(a) |
                reinterpret_cast<uintptr_t>(g1) | reinterpret_cast<uintptr_t>(b1) | reinterpret_cast<uintptr_t>(g2) |
                reinterpret_cast<uintptr_t>(b2)) & 15) == 0, "layer_norm_dual operands must be 16-byte aligned");
  if (rows <= 0) return 0;
  CFM_CUDA_OK(launch_pdl(layer_norm_dual_kernel, dim3(blocks_for(rows, 32)), dim3(256), 0, stream, 1, x, g1, b1, y, g2, b2, a, rows));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
template int layer_norm<float>(const float*, const float*, const float*, float*, int64_t, int, cudaStream_t);
template int layer_norm<bf16>(const float*, const float*, const float*, bf16*, int64_t, int, cudaStream_t);

template <typename T>
int rel_pos_table(const float* div_term, T* pe, int Tn, int d, cudaStream_t stream) {
  CFM_REQUIRE(Tn > 0 && d > 0 && d % 2 == 0, "bad positional table shape");
  const int64_t total = static_cast<int64_t>(2 * Tn - 1) * (d / 2);
  rel_pos_table_kernel<T><<<blocks_for(total, 256), 256, 0, stream>>>(div_term, pe, Tn, d);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
template int rel_pos_table<float>(const float*, float*, int, int, cudaStream_t);
template int rel_pos_table<bf16>(const float*, bf16*, int, int, cudaStream_t);

int lengths_and_mask(const int64_t* lengths_in, int64_t* lengths_out, uint8_t* key_pad, int B, int Tn,
                     bool subsample, bool mark_valid, cudaStream_t stream) {
  CFM_REQUIRE(B > 0 && lengths_in != nullptr, "lengths_and_mask needs lengths");
  CFM_CUDA_OK(launch_pdl(lengths_mask_kernel, dim3(B), dim3(128), 0, stream, 1, lengths_in, lengths_out, key_pad, B, Tn, subsample,
                         mark_valid));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
template <typename T>
int all_masked_mean_v(const uint8_t* key_pad, const T* qkv, T* ctx, int B, int Tn, int d, cudaStream_t stream) {
  CFM_REQUIRE(key_pad != nullptr && B > 0 && Tn > 0, "all_masked_mean_v needs a key mask");
  all_masked_mean_v_kernel<T><<<B, 256, 0, stream>>>(key_pad, qkv, ctx, Tn, d);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
template int all_masked_mean_v<float>(const uint8_t*, const float*, float*, int, int, int, cudaStream_t);
template int all_masked_mean_v<bf16>(const uint8_t*, const bf16*, bf16*, int, int, int, cudaStream_t);

int swish_f32(const float* x, float* y, int64_t n, cudaStream_t stream) {
  swish_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(x, y, n);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
int glu_lastdim_f32(const float* x, float* y, int64_t rows, int d, cudaStream_t stream) {
  glu_kernel<<<blocks_for(rows * d, 256), 256, 0, stream>>>(x, y, rows, d);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
int cast_f32_to_bf16(const float* x, bf16* y, int64_t n, cudaStream_t stream) {
  f32_to_bf16_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(x, y, n);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
int cast_bf16_to_f32(const bf16* x, float* y, int64_t n, cudaStream_t stream) {
  bf16_to_f32_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(x, y, n);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
int pack_conv2_weight(const float* w, float* out, int C, cudaStream_t stream) {
  pack_conv2_kernel<<<blocks_for(static_cast<int64_t>(C) * C * 9, 256), 256, 0, stream>>>(w, out, C);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
int pack_linear_weight(const float* w, float* out, int d, int C, int F2, cudaStream_t stream) {
  pack_linear_kernel<<<blocks_for(static_cast<int64_t>(d) * C * F2, 256), 256, 0, stream>>>(w, out, d, C, F2);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
int pack_glu_rows(const float* w, float* out, int d, int K, cudaStream_t stream) {
  CFM_REQUIRE(d % 32 == 0, "GLU row packing needs channels to be a multiple of 32");
  pack_glu_rows_kernel<<<blocks_for(2LL * d * K, 256), 256, 0, stream>>>(w, out, d, K);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
int fold_dwconv_bn(const float* w, const float* b, const float* bn_w, const float* bn_b, const float* rm,
                   const float* rv, float eps, float* w_out, float* b_out, int d, int k, cudaStream_t stream) {
  fold_dwconv_bn_kernel<<<blocks_for(d, 128), 128, 0, stream>>>(w, b, bn_w, bn_b, rm, rv, eps, w_out, b_out, d, k);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
