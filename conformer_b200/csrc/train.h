// Internal declarations of the training step (train.cu, train_kernels.cu, attention_train.cu).
#pragma once
#include "ops.h"

namespace cfm {

// nn.Dropout as a stateless hash: element i of dropout site `site` is kept with probability 1 - p and scaled by
// 1 / (1 - p); the same (seed, site, i) gives the same decision in the forward and in the backward pass.
// Stateless dropout: keep/drop of element i of dropout site `site` is a hash of (seed, site, i), so the backward
// regenerates the forward's mask instead of storing it.  The seed is read from DEVICE memory (two 32-bit words written by
// cfm_encoder_train_forward before the step's kernels): kernel arguments then do not change from step to step, which is
// what lets a whole training forward / block backward replay as a CUDA graph.
struct DropK {                          // kernel-side: the seed in registers (Drop::load, once per thread)
  float p = 0.f, inv_keep = 1.f;
  uint32_t lo = 0, hi = 0, site = 0;
  // one 32-bit hash per PAIR of elements, 16 bits each (kernels handle 4 consecutive elements per thread, so the hash
  // of a pair is computed once)
  __device__ __forceinline__ float factor(int64_t i) const {
    if (p <= 0.f) return 1.f;
    const uint64_t j = static_cast<uint64_t>(i) >> 1;
    uint32_t h = lo ^ (site * 0x9E3779B9u) ^ static_cast<uint32_t>(j) ^ (static_cast<uint32_t>(j >> 32) * 0x85EBCA6Bu);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    h ^= hi; h *= 0x9E3779B1u; h ^= h >> 15;
    const uint32_t u = (i & 1) ? (h >> 16) : (h & 0xFFFFu);
    return (static_cast<float>(u) * (1.0f / 65536.0f)) >= p ? inv_keep : 0.f;
  }
};
struct Drop {                           // kernel argument
  float p = 0.f, inv_keep = 1.f;
  const uint32_t* seed = nullptr;       // device: {lo, hi}
  uint32_t site = 0;
  __device__ __forceinline__ DropK load() const {
    DropK k;
    k.p = p; k.inv_keep = inv_keep; k.site = site;
    if (p > 0.f) {
      k.lo = __ldg(seed);
      k.hi = __ldg(seed + 1);
    }
    return k;
  }
};
inline Drop make_drop(float p, const uint32_t* seed_dev, uint32_t site) {
  Drop d;
  d.p = p;
  d.inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  d.seed = seed_dev;
  d.site = site;
  return d;
}

// ---- train_kernels.cu
template <typename TOut> int cast_scale_drop(const float* in, TOut* out, int64_t n, float scale, Drop dr, cudaStream_t st);
template <typename TA> int swish_drop_fwd(const TA* hp, TA* out, int64_t n, Drop dr, cudaStream_t st);
template <typename TA> int swish_drop_bwd(const float* dh, const TA* hp, TA* out, int64_t n, Drop dr, cudaStream_t st);
int residual_drop_add(const float* x, const float* t, float* out, int64_t n, float scale, Drop dr, cudaStream_t st);
int drop_inplace(float* x, int64_t n, Drop dr, cudaStream_t st);
template <typename TA> int glu_fwd(const TA* u, TA* g, int64_t rows, int d, cudaStream_t st);
template <typename TA> int glu_bwd(const float* dg, const TA* u, TA* du, int64_t rows, int d, cudaStream_t st);
template <typename TA> int relu_bwd(const float* dy, const TA* y, TA* out, int64_t n, cudaStream_t st);
template <typename T> int transpose_pad(const T* in, int ld, T* out, int M, int N, int Mp, cudaStream_t st);
template <typename T> int colsum_add(const T* in, int ld, float* out, int M, int N, cudaStream_t st);
int ln_bwd(const float* dy, const float* x, const float* gamma, float* dx, bool accumulate, float* dgamma, float* dbeta, int64_t rows,
           int d, cudaStream_t st);
template <typename TA> int dwconv_fwd(const TA* g, const float* w, const float* bias, float* z, int B, int Tn, int d, int ks, cudaStream_t st);
int dwconv_bwd_input(const float* dz, const float* w, float* dg, int B, int Tn, int d, int ks, cudaStream_t st);
template <typename TA> int dwconv_bwd_weight(const float* dz, const TA* g, float* dw, float* db, int B, int Tn, int d, int ks, cudaStream_t st);
int bn_train_stats(const float* z, float* acc2d, float* stats2d, float* running_mean, float* running_var, int M, int d, float momentum,
                   cudaStream_t st);
template <typename TA> int bn_swish_fwd(const float* z, const float* stats, const float* gamma, const float* beta, TA* s, int64_t M, int d, cudaStream_t st);
int bn_swish_bwd(float* ds_to_dz, const float* z, const float* stats, const float* gamma, const float* beta, float* acc2d, float* dgamma,
                 float* dbeta, int M, int d, cudaStream_t st);
template <typename TA> int im2col_t(const TA* out1, TA* colT, const ConvGeom& g, int64_t M2, int64_t M2p, cudaStream_t st);
int im2col_rows(const bf16* out1, bf16* col /*[M2, 9C]*/, const ConvGeom& g, int64_t M2, cudaStream_t st);
template <typename TA> int col2im_relu(const TA* dcol, const TA* out1, float* dout1, const ConvGeom& g, cudaStream_t st);
template <typename TA>
int conv1_wgrad_fused(const TA* dcol, const TA* out1, const float* mel, float* dw, float* db, const ConvGeom& g, int n_mel, int Tin, int T1,
                      int num_sms, cudaStream_t st);
int conv1_wgrad(const float* dout1, const float* mel, float* dw, float* db, int B, int n_mel, int Tin, int T1, int T1p, int F1, int C,
                cudaStream_t st);

// ---- attention_train.cu: relative-position attention with saved log-sum-exp, and its backward
// qkv [B,T,3d] (q|k|v, head h at columns h*dh), pp [2T-1, d] projected positional table, u/v content/position bias [H,dh],
// key_pad [B,T] (1 = masked) or null.  ctx [B,T,d]; lse fp32 [B,H,T] (natural log of the softmax normaliser of the scaled
// scores).  Attention-weight dropout (attention.py:67) through `dr` (element index ((b*H+h)*T+i)*T+j).
template <typename TA>
int attn_train_fwd(const TA* qkv, const TA* pp, const float* u, const float* v, const uint8_t* key_pad, TA* ctx, float* lse, int B, int Tn,
                   int H, int dh, Drop dr, cudaStream_t st);
// dctx [B,T,d] fp32 -> dqkv TA [B,T,3d]; dpp fp32 [2T-1,d] (+=, summed over the batch), du/dv fp32 [H,dh] (+=).
// scratch: fp32 [B,H,T] (row term D_i)
template <typename TA>
int attn_train_bwd(const TA* qkv, const TA* pp, const float* u, const float* v, const uint8_t* key_pad, const TA* ctx, const float* lse,
                   const float* dctx, TA* dqkv, float* dpp, float* du, float* dv, float* scratch, int B, int Tn, int H, int dh, Drop dr,
                   cudaStream_t st);

// attention_train_mma.cu: the backward with tensor-core products (bf16); D = the row term computed by attn_train_bwd
int attn_train_bwd_mma(const bf16* qkv, const bf16* pp, const float* u, const float* v, const uint8_t* key_pad, const float* lse,
                       const float* dctx, bf16* dqkv, float* dpp, float* du, float* dv, const float* scratch_D, int B, int Tn, int H, int dh,
                       Drop dr, cudaStream_t st);
// attention_tc.cu: the tcgen05 forward kernel in training form (head size 64, bf16)
int rel_attention_bf16_tc_train(const bf16* qkv, const bf16* pos, int ldp, const float* u, const float* v, const uint8_t* key_pad,
                                bf16* ctx, float* lse, Drop dr, int B, int Tn, int H, int dh, cudaStream_t stream);

}  // namespace cfm
