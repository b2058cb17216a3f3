// Relative-position multi-head self-attention on CUDA cores with fp32 arithmetic: the parity
// path (and the path for head sizes the tensor-core kernel does not take).
//
// Restates scaled_dot_product_relative_attention (model/utils/attention.py:47-72) flash-style:
//   score[i,j] = ((q_i+u).k_j + (q_i+v).p[T-1+j-i]) / sqrt(dh)
// The Transformer-XL pad/view/slice shift (_relative_shift, attention.py:94-102) is the index
// identity pos[i, T-1+j-i]; for a 32x32 score tile it touches one contiguous window of 63 rows of
// the projected table, which is staged in shared memory - no [T,2T-1] tensor is ever formed.
// Masked keys (and keys past T) get weight exactly 0, as masked_fill(finfo.min)+softmax gives in
// the reference whenever a row has at least one valid key.
#include <math_constants.h>

#include "ops.h"

namespace cfm {
namespace {

constexpr int BQ = 32, BKV = 32, DHMAX = 64, LD = DHMAX + 1;
constexpr int ATT_SMEM_FLOATS = BQ * LD * 2 + BKV * LD + BKV * DHMAX + (BQ + BKV - 1) * LD + BQ * (BKV + 1);

template <typename T>
__global__ void __launch_bounds__(128) rel_attention_f32_kernel(const T* __restrict__ qkv, const T* __restrict__ pos,
                                                                int ldp, const float* __restrict__ ub,
                                                                const float* __restrict__ vb,
                                                                const uint8_t* __restrict__ key_pad,
                                                                T* __restrict__ ctx, int Tn, int H, int dh) {
  extern __shared__ float sm[];
  float* Qu = sm;                         // [BQ][LD]   q + content_bias
  float* Qv = Qu + BQ * LD;               // [BQ][LD]   q + position_bias
  float* Ks = Qv + BQ * LD;               // [BKV][LD]
  float* Vs = Ks + BKV * LD;              // [BKV][DHMAX]
  float* Pw = Vs + BKV * DHMAX;           // [BQ+BKV-1][LD]  window of the projected table
  float* Ss = Pw + (BQ + BKV - 1) * LD;   // [BQ][BKV+1]
  const int tid = threadIdx.x;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * BQ;
  const int d = H * dh;
  const int64_t row_stride = 3 * static_cast<int64_t>(d);
  const T* qbase = qkv + static_cast<int64_t>(b) * Tn * row_stride + h * dh;

  for (int idx = tid; idx < BQ * dh; idx += 128) {
    const int i = idx / dh, c = idx - i * dh;
    float q = 0.f;
    if (i0 + i < Tn) q = to_f32(qbase[(i0 + i) * row_stride + c]);
    Qu[i * LD + c] = q + ub[h * dh + c];
    Qv[i * LD + c] = q + vb[h * dh + c];
  }
  const int i = tid >> 2, q4 = tid & 3;
  float m_run = -CUDART_INF_F, l_run = 0.f;
  float acc[DHMAX / 4];
#pragma unroll
  for (int cc = 0; cc < DHMAX / 4; ++cc) acc[cc] = 0.f;
  const float inv_sqrt = 1.0f / sqrtf(static_cast<float>(dh));

  for (int j0 = 0; j0 < Tn; j0 += BKV) {
    __syncthreads();                                       // previous tile fully consumed (also covers Q staging)
    for (int idx = tid; idx < BKV * dh; idx += 128) {
      const int j = idx / dh, c = idx - j * dh;
      float kv = 0.f, vv = 0.f;
      if (j0 + j < Tn) {
        kv = to_f32(qbase[(j0 + j) * row_stride + d + c]);
        vv = to_f32(qbase[(j0 + j) * row_stride + 2 * d + c]);
      }
      Ks[j * LD + c] = kv;
      Vs[j * DHMAX + c] = vv;
    }
    const int r0 = Tn - 1 + j0 - i0 - (BQ - 1);            // table row of window entry 0
    for (int idx = tid; idx < (BQ + BKV - 1) * dh; idx += 128) {
      const int w = idx / dh, c = idx - w * dh;
      const int r = r0 + w;
      Pw[w * LD + c] = (r >= 0 && r <= 2 * Tn - 2) ? to_f32(pos[static_cast<int64_t>(r) * ldp + h * dh + c]) : 0.f;
    }
    __syncthreads();

    float s[BKV / 4];
#pragma unroll
    for (int jj = 0; jj < BKV / 4; ++jj) s[jj] = 0.f;
    for (int c = 0; c < dh; ++c) {
      const float qu = Qu[i * LD + c], qv = Qv[i * LD + c];
#pragma unroll
      for (int jj = 0; jj < BKV / 4; ++jj) {
        const int j = q4 + 4 * jj;
        s[jj] = fmaf(qu, Ks[j * LD + c], s[jj]);
        s[jj] = fmaf(qv, Pw[(j - i + BQ - 1) * LD + c], s[jj]);
      }
    }
    float tmax = -CUDART_INF_F;
    bool dead[BKV / 4];
#pragma unroll
    for (int jj = 0; jj < BKV / 4; ++jj) {
      const int j = j0 + q4 + 4 * jj;
      dead[jj] = (j >= Tn) || (key_pad != nullptr && key_pad[static_cast<int64_t>(b) * Tn + j] != 0);
      s[jj] = s[jj] * inv_sqrt;
      if (!dead[jj]) tmax = fmaxf(tmax, s[jj]);
    }
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
    tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 2));
    const float m_new = fmaxf(m_run, tmax);
    const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
    const float alpha = expf(m_run - m_use);               // 0 when nothing accumulated yet
    float psum = 0.f;
#pragma unroll
    for (int jj = 0; jj < BKV / 4; ++jj) {
      const float p = dead[jj] ? 0.f : expf(s[jj] - m_use);
      psum += p;
      Ss[i * (BKV + 1) + q4 + 4 * jj] = p;
    }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    l_run = l_run * alpha + psum;
    m_run = m_new;
    __syncwarp();                                          // a row's 4 threads are in one warp
#pragma unroll
    for (int cc = 0; cc < DHMAX / 4; ++cc) acc[cc] *= alpha;
    for (int j = 0; j < BKV; ++j) {
      const float p = Ss[i * (BKV + 1) + j];
#pragma unroll
      for (int cc = 0; cc < DHMAX / 4; ++cc) {
        const int c = q4 + 4 * cc;
        if (c < dh) acc[cc] = fmaf(p, Vs[j * DHMAX + c], acc[cc]);
      }
    }
  }
  if (i0 + i < Tn) {
    const float inv_l = (l_run > 0.f) ? 1.0f / l_run : 0.f;
    T* orow = ctx + (static_cast<int64_t>(b) * Tn + i0 + i) * d + h * dh;
#pragma unroll
    for (int cc = 0; cc < DHMAX / 4; ++cc) {
      const int c = q4 + 4 * cc;
      if (c < dh) orow[c] = from_f32<T>(acc[cc] * inv_l);
    }
  }
}

}  // namespace

template <typename T>
int rel_attention(const T* qkv, const T* pos, int ldp, const float* u, const float* v, const uint8_t* key_pad,
                  T* ctx, int B, int Tn, int H, int dh, cudaStream_t stream) {
  CFM_REQUIRE(dh >= 1 && dh <= DHMAX, "SIMT attention supports head size <= 64");
  CFM_REQUIRE(B > 0 && Tn > 0 && H > 0 && H <= 65535 && B <= 65535, "bad attention shape");
  const size_t smem = ATT_SMEM_FLOATS * sizeof(float);
  static std::atomic<uint64_t> attr_done{0};        // one per template instantiation
  CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(rel_attention_f32_kernel<T>), static_cast<int>(smem)));
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, BQ)), static_cast<unsigned>(H), static_cast<unsigned>(B));
  rel_attention_f32_kernel<T><<<grid, 128, smem, stream>>>(qkv, pos, ldp, u, v, key_pad, ctx, Tn, H, dh);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
template int rel_attention<float>(const float*, const float*, int, const float*, const float*, const uint8_t*, float*, int, int, int, int, cudaStream_t);
template int rel_attention<bf16>(const bf16*, const bf16*, int, const float*, const float*, const uint8_t*, bf16*, int, int, int, int, cudaStream_t);

}  // namespace cfm
