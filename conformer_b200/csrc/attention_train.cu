// Relative-position attention for the TRAINING step: forward with the saved softmax normaliser, and the backward pass
// (attention.py:47-72 under autograd; BASELINE config 4).
//
//   s_ij = ((q_i + u) . k_j + (q_i + v) . p_{T-1+j-i}) / sqrt(dh)        (the relative shift as an index)
//   P = softmax_j(s) with masked keys at weight 0,  Pd = dropout(P),  ctx_i = sum_j Pd_ij v_j
//
//   backward (flash-attention style, nothing of size T x T is stored):  D_i = dctx_i . ctx_i,
//     dPd_ij = dctx_i . v_j,   ds_ij = P_ij (f_ij dPd_ij - D_i) / sqrt(dh)          (f = dropout factor)
//     dq_i = sum_j ds_ij (k_j + p_r),  dk_j = sum_i ds_ij (q_i + u),  dv_j = sum_i Pd_ij dctx_i,
//     dp_r = sum_{j-i = r-(T-1)} ds_ij (q_i + v),  du = sum_i sum_j ds_ij k_j,  dv_bias = sum_i sum_j ds_ij p_r
//
// All three kernels work on 64 x 64 (query x key) tiles with 256 threads; thread (ti, tj) owns the 4 x 4 entries
// i = ti + 16 a, j = tj + 16 e and computes their two dot products from shared-memory tiles with 128-bit loads (the
// positional operand of entry (i, j) is row j - i + 63 of a 127-row window of the projected table, so no shifted copy
// exists); the "P x something" products re-read the score tile from shared memory with each thread accumulating a
// [4 rows x 4 channels] block.  fp32 arithmetic on the CUDA cores throughout: this is the first, correctness-first
// version of the backward (the forward of the training step at head size 64 is the tcgen05 kernel of attention_tc.cu, the
// forward kernel here serves the other head sizes and the fp32 mode); its cost is reported next to the GEMMs'.
#include <math_constants.h>

#include "train.h"

namespace cfm {
namespace {

constexpr int TQ = 64, TK = 64, DHP = 64, LDS_ = DHP + 4, WIN = TQ + TK - 1;      // 127-row positional window
constexpr int THREADS = 256;
// Operand tiles (q+u, q+v, dctx, k, v, positional window) are held in shared memory in the ACTIVATION type: fp32 in the
// gradient-parity mode, bf16 in bf16 mode, where they halve the footprint (2 CTAs per SM instead of 1: the second CTA's
// FMAs hide the first one's shared-memory latency).  Score tiles (P, ds) stay fp32.  Row pitch: 68 floats / 72 bf16 -
// both conflict-free for the 4-element accesses of consecutive rows.
template <typename TS> struct Tile;
template <> struct Tile<float> {
  static constexpr int LD = DHP + 4;
  static __device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Tile<bf16> {
  static constexpr int LD = DHP + 8;
  static __device__ __forceinline__ float4 ld4(const bf16* p) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    return make_float4(__uint_as_float(t.x << 16), __uint_as_float(t.x & 0xffff0000u), __uint_as_float(t.y << 16),
                       __uint_as_float(t.y & 0xffff0000u));
  }
  static __device__ __forceinline__ void st4(bf16* p, float4 v) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
};

struct AttnArgs {
  const void* qkv;
  const void* pp;
  const float* u;
  const float* v;
  const uint8_t* key_pad;
  int Tn, H, dh;
  float scale;
  Drop dr;
};

template <typename TA, typename TS>
__device__ __forceinline__ void load_rows(TS* dst, const TA* src, int64_t row_stride, int row0, int nrows_valid_lo, int nrows_valid_hi,
                                          int rows, int dh, const float* add) {
  // dst[r][c] = src[(row0 + r) * row_stride + c] (+ add[c]) for row0 + r in [lo, hi), zeros elsewhere and for c >= dh
  for (int idx = threadIdx.x; idx < rows * (DHP / 4); idx += THREADS) {
    const int r = idx / (DHP / 4), c4 = (idx - r * (DHP / 4)) * 4;
    const int row = row0 + r;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row >= nrows_valid_lo && row < nrows_valid_hi && c4 < dh) {
      const TA* p = src + static_cast<int64_t>(row) * row_stride + c4;
      o = make_float4(to_f32(p[0]), to_f32(p[1]), to_f32(p[2]), to_f32(p[3]));
      if (add != nullptr) { o.x += add[c4]; o.y += add[c4 + 1]; o.z += add[c4 + 2]; o.w += add[c4 + 3]; }
    }
    Tile<TS>::st4(dst + r * Tile<TS>::LD + c4, o);
  }
}
template <typename TS>
__device__ __forceinline__ void load_rows_f32(TS* dst, const float* src, int64_t row_stride, int row0, int hi, int rows, int dh) {
  for (int idx = threadIdx.x; idx < rows * (DHP / 4); idx += THREADS) {
    const int r = idx / (DHP / 4), c4 = (idx - r * (DHP / 4)) * 4;
    const int row = row0 + r;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < hi && c4 < dh) o = *reinterpret_cast<const float4*>(src + static_cast<int64_t>(row) * row_stride + c4);
    Tile<TS>::st4(dst + r * Tile<TS>::LD + c4, o);
  }
}

__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}

// s[a][e] = scale * (Qc_i . K_j + Qp_i . Pw_{j-i+63}), i = ti + 16a, j = tj + 16e
template <typename TS>
__device__ __forceinline__ void score_tile(const TS* Qc, const TS* Qp, const TS* Kt, const TS* Pw, int ti, int tj, float scale,
                                           float (&s)[4][4]) {
  constexpr int LD = Tile<TS>::LD;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[a][e] = 0.f;
  const int wbase = tj - ti + (TQ - 1);
#pragma unroll 4
  for (int c4 = 0; c4 < DHP; c4 += 4) {
    float4 qc[4], qp[4], kk[4], pw[7];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      qc[a] = Tile<TS>::ld4(Qc + (ti + 16 * a) * LD + c4);
      qp[a] = Tile<TS>::ld4(Qp + (ti + 16 * a) * LD + c4);
      kk[a] = Tile<TS>::ld4(Kt + (tj + 16 * a) * LD + c4);
    }
#pragma unroll
    for (int dlt = 0; dlt < 7; ++dlt) pw[dlt] = Tile<TS>::ld4(Pw + (wbase + 16 * (dlt - 3)) * LD + c4);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[a][e] = dot4(qp[a], pw[e - a + 3], dot4(qc[a], kk[e], s[a][e]));
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[a][e] *= scale;
}
// plain [4][4] block of A_i . B_j
template <typename TS>
__device__ __forceinline__ void dot_tile(const TS* A, const TS* Bm, int ti, int tj, float (&s)[4][4]) {
  constexpr int LD = Tile<TS>::LD;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[a][e] = 0.f;
#pragma unroll 4
  for (int c4 = 0; c4 < DHP; c4 += 4) {
    float4 av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      av[a] = Tile<TS>::ld4(A + (ti + 16 * a) * LD + c4);
      bv[a] = Tile<TS>::ld4(Bm + (tj + 16 * a) * LD + c4);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[a][e] = dot4(av[a], bv[e], s[a][e]);
  }
}
// acc[a][0..3] += sum_j S[i][j] * Bm[row(j)][4 tj ..], i = ti + 16a; row(j) = j + row_shift(i) (0 for K/V, j - i + 63 for the window)
template <bool WINDOW, typename TS>
__device__ __forceinline__ void row_matmul(const float* S, const TS* Bm, int ti, int tj, float (&acc)[4][4]) {
  constexpr int LD = Tile<TS>::LD;
  for (int j = 0; j < TK; ++j) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = ti + 16 * a;
      const float sv = S[i * LDS_ + j];
      const float4 b4 = Tile<TS>::ld4(Bm + (WINDOW ? (j - i + TQ - 1) : j) * LD + 4 * tj);
      acc[a][0] = fmaf(sv, b4.x, acc[a][0]); acc[a][1] = fmaf(sv, b4.y, acc[a][1]);
      acc[a][2] = fmaf(sv, b4.z, acc[a][2]); acc[a][3] = fmaf(sv, b4.w, acc[a][3]);
    }
  }
}
// acc[a][0..3] += sum_i S[i][j] * A[i][4 tj ..], j = ti + 16a   (the transposed product of the key-side kernel)
template <typename TS>
__device__ __forceinline__ void col_matmul(const float* S, const TS* A, int ti, int tj, float (&acc)[4][4]) {
  constexpr int LD = Tile<TS>::LD;
  for (int i = 0; i < TQ; ++i) {
    const float4 a4 = Tile<TS>::ld4(A + i * LD + 4 * tj);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float sv = S[i * LDS_ + ti + 16 * a];
      acc[a][0] = fmaf(sv, a4.x, acc[a][0]); acc[a][1] = fmaf(sv, a4.y, acc[a][1]);
      acc[a][2] = fmaf(sv, a4.z, acc[a][2]); acc[a][3] = fmaf(sv, a4.w, acc[a][3]);
    }
  }
}
__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ forward
template <typename TA>
__global__ void __launch_bounds__(THREADS) attn_fwd_kernel(AttnArgs p, TA* __restrict__ ctx, float* __restrict__ lse) {
  pdl_wait();
  extern __shared__ __align__(16) uint8_t sm_raw[];
  constexpr int LD = Tile<TA>::LD;
  float* Ps = reinterpret_cast<float*>(sm_raw);                  // [TQ][LDS_] fp32
  TA* Qc = reinterpret_cast<TA*>(Ps + TQ * LDS_);
  TA* Qp = Qc + TQ * LD;
  TA* Kt = Qp + TQ * LD;
  TA* Vt = Kt + TK * LD;
  TA* Pw = Vt + TK * LD;                                         // [WIN + 1][LD]
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TQ;
  const int Tn = p.Tn, dh = p.dh, d = p.H * dh;
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const TA* qkv = static_cast<const TA*>(p.qkv) + static_cast<int64_t>(b) * Tn * 3 * d + h * dh;
  const TA* pp = static_cast<const TA*>(p.pp) + h * dh;
  load_rows<TA>(Qc, qkv, 3 * d, i0, 0, Tn, TQ, dh, p.u + h * dh);
  load_rows<TA>(Qp, qkv, 3 * d, i0, 0, Tn, TQ, dh, p.v + h * dh);
  float m[4], l[4], o[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m[a] = -CUDART_INF_F;
    l[a] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) o[a][c] = 0.f;
  }
  const DropK dk = p.dr.load();
  for (int j0 = 0; j0 < Tn; j0 += TK) {
    __syncthreads();
    load_rows<TA>(Kt, qkv + d, 3 * d, j0, 0, Tn, TK, dh, nullptr);
    load_rows<TA>(Vt, qkv + 2 * d, 3 * d, j0, 0, Tn, TK, dh, nullptr);
    load_rows<TA>(Pw, pp, d, (Tn - 1) + j0 - i0 - (TQ - 1), 0, 2 * Tn - 1, WIN + 1, dh, nullptr);
    __syncthreads();
    float s[4][4];
    score_tile(Qc, Qp, Kt, Pw, ti, tj, p.scale, s);
    bool dead[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j0 + tj + 16 * e;
      dead[e] = j >= Tn || (p.key_pad != nullptr && p.key_pad[static_cast<int64_t>(b) * Tn + j] != 0);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int e = 0; e < 4; ++e) mx = dead[e] ? mx : fmaxf(mx, s[a][e]);
      mx = group16_max(mx);
      const float mn = fmaxf(m[a], mx);
      const float corr = (mn == -CUDART_INF_F) ? 1.f : __expf(m[a] - mn);
      float rs = 0.f;
      const int i = i0 + ti + 16 * a;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = dead[e] ? 0.f : __expf(s[a][e] - mn);
        rs += pv;
        const int j = j0 + tj + 16 * e;
        const float f = dk.factor((static_cast<int64_t>(b * p.H + h) * Tn + i) * Tn + j);
        Ps[(ti + 16 * a) * LDS_ + tj + 16 * e] = pv * f;
      }
      l[a] = l[a] * corr + group16_sum(rs);
      m[a] = mn;
#pragma unroll
      for (int c = 0; c < 4; ++c) o[a][c] *= corr;
    }
    __syncthreads();
    row_matmul<false>(Ps, Vt, ti, tj, o);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ti + 16 * a;
    if (i >= Tn) continue;
    const float inv = l[a] > 0.f ? 1.f / l[a] : 0.f;
    if (4 * tj < dh) {
      TA* dst = ctx + (static_cast<int64_t>(b) * Tn + i) * d + h * dh + 4 * tj;
#pragma unroll
      for (int c = 0; c < 4; ++c) dst[c] = from_f32<TA>(o[a][c] * inv);
    }
    if (tj == 0) lse[(static_cast<int64_t>(b) * p.H + h) * Tn + i] = m[a] + __logf(fmaxf(l[a], 1e-30f));
  }
}

// D[b,h,i] = dctx_i . ctx_i over the head's channels
template <typename TA>
__global__ void attn_delta_kernel(const float* __restrict__ dctx, const TA* __restrict__ ctx, float* __restrict__ D, int B, int Tn, int H, int dh) {
  pdl_wait();
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);     // (b, i, h)
  const int lane = threadIdx.x & 31;
  if (row >= static_cast<int64_t>(B) * Tn * H) return;
  const int h = static_cast<int>(row % H);
  const int64_t bi = row / H;
  float acc = 0.f;
  for (int c = lane; c < dh; c += 32) acc = fmaf(dctx[bi * H * dh + h * dh + c], to_f32(ctx[bi * H * dh + h * dh + c]), acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    const int64_t b = bi / Tn, i = bi - b * Tn;
    D[(b * H + h) * Tn + i] = acc;
  }
}

// shared by the two backward kernels: P (undropped), Pd = f P and ds (scaled) of this thread's 4 x 4 entries
__device__ __forceinline__ void bwd_entries(const AttnArgs& p, const float (&s)[4][4], const float (&dpd)[4][4], const float* lse_s,
                                            const float* D_s, int b, int h, int i0, int j0, int ti, int tj, float* Ps, float* DSs) {
  const DropK dk = p.dr.load();
  const int Tn = p.Tn;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int j = j0 + tj + 16 * e;
    const bool dead = j >= Tn || (p.key_pad != nullptr && p.key_pad[static_cast<int64_t>(b) * Tn + j] != 0);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int il = ti + 16 * a, i = i0 + il;
      float pd = 0.f, ds = 0.f;
      if (!dead && i < Tn) {
        const float pv = __expf(s[a][e] - lse_s[il]);
        const float f = dk.factor((static_cast<int64_t>(b * p.H + h) * Tn + i) * Tn + j);
        pd = pv * f;
        ds = pv * (f * dpd[a][e] - D_s[il]) * p.scale;
      }
      Ps[il * LDS_ + tj + 16 * e] = pd;
      DSs[il * LDS_ + tj + 16 * e] = ds;
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward, query side
// one CTA = 64 queries of one (b, h): dq (content + positional parts), the window gradient of the projected table, du, dv
template <typename TA>
__global__ void __launch_bounds__(THREADS, 2) attn_bwd_q_kernel(AttnArgs p, const float* __restrict__ lse, const float* __restrict__ D,
                                                             const float* __restrict__ dctx, TA* __restrict__ dqkv, float* __restrict__ dpp,
                                                             float* __restrict__ du, float* __restrict__ dv) {
  pdl_wait();
  extern __shared__ __align__(16) uint8_t sm_raw[];
  constexpr int LD = Tile<TA>::LD;
  float* Ps = reinterpret_cast<float*>(sm_raw);                  // fp32 score tiles and small vectors first
  float* DSs = Ps + TQ * LDS_;
  float* lse_s = DSs + TQ * LDS_;                                // [TQ]
  float* D_s = lse_s + TQ;
  float* red = D_s + TQ;                                         // [2][DHP]
  TA* Qc = reinterpret_cast<TA*>(red + 2 * DHP);
  TA* Qp = Qc + TQ * LD;
  TA* dO = Qp + TQ * LD;
  TA* Kt = dO + TQ * LD;
  TA* Vt = Kt + TK * LD;
  TA* Pw = Vt + TK * LD;                                         // [WIN + 1][LD]
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TQ;
  const int Tn = p.Tn, dh = p.dh, d = p.H * dh;
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const TA* qkv = static_cast<const TA*>(p.qkv) + static_cast<int64_t>(b) * Tn * 3 * d + h * dh;
  const TA* pp = static_cast<const TA*>(p.pp) + h * dh;
  load_rows<TA>(Qc, qkv, 3 * d, i0, 0, Tn, TQ, dh, p.u + h * dh);
  load_rows<TA>(Qp, qkv, 3 * d, i0, 0, Tn, TQ, dh, p.v + h * dh);
  load_rows_f32(dO, dctx + static_cast<int64_t>(b) * Tn * d + h * dh, d, i0, Tn, TQ, dh);
  for (int i = threadIdx.x; i < TQ; i += THREADS) {
    const bool ok = i0 + i < Tn;
    lse_s[i] = ok ? lse[(static_cast<int64_t>(b) * p.H + h) * Tn + i0 + i] : 0.f;
    D_s[i] = ok ? D[(static_cast<int64_t>(b) * p.H + h) * Tn + i0 + i] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * DHP; i += THREADS) red[i] = 0.f;
  float dqc[4][4], dqp[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) dqc[a][c] = dqp[a][c] = 0.f;
  // window-gradient accumulators: rows w = ti + 16 m (m = 0..7), channels 4 tj .. 4 tj + 3; rows >= 64 carry over to the
  // next key tile, whose window starts 64 table rows later
  float dw[8][4];
#pragma unroll
  for (int mm = 0; mm < 8; ++mm)
#pragma unroll
    for (int c = 0; c < 4; ++c) dw[mm][c] = 0.f;
  for (int j0 = 0; j0 < Tn; j0 += TK) {
    __syncthreads();
    load_rows<TA>(Kt, qkv + d, 3 * d, j0, 0, Tn, TK, dh, nullptr);
    load_rows<TA>(Vt, qkv + 2 * d, 3 * d, j0, 0, Tn, TK, dh, nullptr);
    const int r0 = (Tn - 1) + j0 - i0 - (TQ - 1);
    load_rows<TA>(Pw, pp, d, r0, 0, 2 * Tn - 1, WIN + 1, dh, nullptr);
    __syncthreads();
    float s[4][4], dpd[4][4];
    score_tile(Qc, Qp, Kt, Pw, ti, tj, p.scale, s);
    dot_tile(dO, Vt, ti, tj, dpd);
    bwd_entries(p, s, dpd, lse_s, D_s, b, h, i0, j0, ti, tj, Ps, DSs);
    __syncthreads();
    row_matmul<false>(DSs, Kt, ti, tj, dqc);
    row_matmul<true>(DSs, Pw, ti, tj, dqp);
    // dPw[w][c] += sum_i ds[i][i + w - 63] Qp[i][c]
#pragma unroll
    for (int mm = 0; mm < 8; ++mm) {
      const int w = ti + 16 * mm;
      const int ilo = max(0, (TQ - 1) - w), ihi = min(TQ, TK + (TQ - 1) - w);      // 0 <= i + w - 63 < 64
      for (int i = ilo; i < ihi; ++i) {
        const float sv = DSs[i * LDS_ + i + w - (TQ - 1)];
        const float4 q4 = Tile<TA>::ld4(Qp + i * LD + 4 * tj);
        dw[mm][0] = fmaf(sv, q4.x, dw[mm][0]); dw[mm][1] = fmaf(sv, q4.y, dw[mm][1]);
        dw[mm][2] = fmaf(sv, q4.z, dw[mm][2]); dw[mm][3] = fmaf(sv, q4.w, dw[mm][3]);
      }
    }
    // rows w < 64 are complete: flush them; rows w >= 64 become rows w - 64 of the next window
    const bool last = j0 + TK >= Tn;
#pragma unroll
    for (int mm = 0; mm < 8; ++mm) {
      if (mm >= 4 && !last) break;
      const int r = r0 + ti + 16 * mm;
      if (r >= 0 && r < 2 * Tn - 1 && 4 * tj < dh) {
        if ((d & 3) == 0 && (dh & 3) == 0) {                       // one 16-byte vector reduction per 4 channels
          atomicAdd(reinterpret_cast<float4*>(dpp + static_cast<int64_t>(r) * d + h * dh + 4 * tj), make_float4(dw[mm][0], dw[mm][1], dw[mm][2], dw[mm][3]));
        } else {
          float* dst = dpp + static_cast<int64_t>(r) * d + h * dh + 4 * tj;
#pragma unroll
          for (int c = 0; c < 4; ++c) atomicAdd(dst + c, dw[mm][c]);
        }
      }
    }
#pragma unroll
    for (int mm = 0; mm < 4; ++mm)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        dw[mm][c] = dw[mm + 4][c];
        dw[mm + 4][c] = 0.f;
      }
  }
  // dq, and the bias gradients du = sum_i dq_content, dv = sum_i dq_positional
  float su[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ti + 16 * a;
    if (i >= Tn) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      su[c] += dqc[a][c];
      sv[c] += dqp[a][c];
    }
    if (4 * tj < dh) {
      TA* dst = dqkv + (static_cast<int64_t>(b) * Tn + i) * 3 * d + h * dh + 4 * tj;
#pragma unroll
      for (int c = 0; c < 4; ++c) dst[c] = from_f32<TA>(dqc[a][c] + dqp[a][c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    atomicAdd(&red[4 * tj + c], su[c]);
    atomicAdd(&red[DHP + 4 * tj + c], sv[c]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < dh; c += THREADS) {
    atomicAdd(du + h * dh + c, red[c]);
    atomicAdd(dv + h * dh + c, red[DHP + c]);
  }
}

// ------------------------------------------------------------------------------------------------ backward, key side
// one CTA = 64 keys of one (b, h): dk and dv
template <typename TA>
__global__ void __launch_bounds__(THREADS, 2) attn_bwd_kv_kernel(AttnArgs p, const float* __restrict__ lse, const float* __restrict__ D,
                                                              const float* __restrict__ dctx, TA* __restrict__ dqkv) {
  pdl_wait();
  extern __shared__ __align__(16) uint8_t sm_raw[];
  constexpr int LD = Tile<TA>::LD;
  float* Ps = reinterpret_cast<float*>(sm_raw);
  float* DSs = Ps + TQ * LDS_;
  float* lse_s = DSs + TQ * LDS_;
  float* D_s = lse_s + TQ;
  TA* Qc = reinterpret_cast<TA*>(D_s + TQ);
  TA* Qp = Qc + TQ * LD;
  TA* dO = Qp + TQ * LD;
  TA* Kt = dO + TQ * LD;
  TA* Vt = Kt + TK * LD;
  TA* Pw = Vt + TK * LD;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * TK;
  const int Tn = p.Tn, dh = p.dh, d = p.H * dh;
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const TA* qkv = static_cast<const TA*>(p.qkv) + static_cast<int64_t>(b) * Tn * 3 * d + h * dh;
  const TA* pp = static_cast<const TA*>(p.pp) + h * dh;
  load_rows<TA>(Kt, qkv + d, 3 * d, j0, 0, Tn, TK, dh, nullptr);
  load_rows<TA>(Vt, qkv + 2 * d, 3 * d, j0, 0, Tn, TK, dh, nullptr);
  float dk[4][4], dvv[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) dk[a][c] = dvv[a][c] = 0.f;
  for (int i0 = 0; i0 < Tn; i0 += TQ) {
    __syncthreads();
    load_rows<TA>(Qc, qkv, 3 * d, i0, 0, Tn, TQ, dh, p.u + h * dh);
    load_rows<TA>(Qp, qkv, 3 * d, i0, 0, Tn, TQ, dh, p.v + h * dh);
    load_rows_f32(dO, dctx + static_cast<int64_t>(b) * Tn * d + h * dh, d, i0, Tn, TQ, dh);
    load_rows<TA>(Pw, pp, d, (Tn - 1) + j0 - i0 - (TQ - 1), 0, 2 * Tn - 1, WIN + 1, dh, nullptr);
    for (int i = threadIdx.x; i < TQ; i += THREADS) {
      const bool ok = i0 + i < Tn;
      lse_s[i] = ok ? lse[(static_cast<int64_t>(b) * p.H + h) * Tn + i0 + i] : 0.f;
      D_s[i] = ok ? D[(static_cast<int64_t>(b) * p.H + h) * Tn + i0 + i] : 0.f;
    }
    __syncthreads();
    float s[4][4], dpd[4][4];
    score_tile(Qc, Qp, Kt, Pw, ti, tj, p.scale, s);
    dot_tile(dO, Vt, ti, tj, dpd);
    bwd_entries(p, s, dpd, lse_s, D_s, b, h, i0, j0, ti, tj, Ps, DSs);
    __syncthreads();
    col_matmul(Ps, dO, ti, tj, dvv);                              // dv_j += sum_i Pd_ij dctx_i
    col_matmul(DSs, Qc, ti, tj, dk);                              // dk_j += sum_i ds_ij (q_i + u)
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = j0 + ti + 16 * a;
    if (j >= Tn || 4 * tj >= dh) continue;
    TA* dst = dqkv + (static_cast<int64_t>(b) * Tn + j) * 3 * d + h * dh + 4 * tj;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      dst[d + c] = from_f32<TA>(dk[a][c]);
      dst[2 * d + c] = from_f32<TA>(dvv[a][c]);
    }
  }
}

template <typename TA> constexpr int fwd_smem() { return TQ * LDS_ * 4 + (4 * TQ + WIN + 1) * Tile<TA>::LD * static_cast<int>(sizeof(TA)); }
template <typename TA> constexpr int bwd_smem() {
  return (2 * TQ * LDS_ + 2 * TQ + 2 * DHP) * 4 + (5 * TQ + WIN + 1) * Tile<TA>::LD * static_cast<int>(sizeof(TA));
}

}  // namespace

template <typename TA>
int attn_train_fwd(const TA* qkv, const TA* pp, const float* u, const float* v, const uint8_t* key_pad, TA* ctx, float* lse, int B, int Tn,
                   int H, int dh, Drop dr, cudaStream_t st) {
  CFM_REQUIRE(dh >= 4 && dh <= DHP && dh % 4 == 0, "training attention supports head sizes that are multiples of 4 up to 64");
  CFM_REQUIRE(B > 0 && Tn > 0 && H > 0 && H <= 65535 && B <= 65535, "bad attention shape");
  static std::atomic<uint64_t> done{0};
  CFM_TRY(smem_attr_once(done, reinterpret_cast<const void*>(attn_fwd_kernel<TA>), fwd_smem<TA>()));
  AttnArgs a{qkv, pp, u, v, key_pad, Tn, H, dh, 1.0f / sqrtf(static_cast<float>(dh)), dr};
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, TQ)), H, B);
  CFM_CUDA_OK(launch_pdl(attn_fwd_kernel<TA>, dim3(grid), dim3(THREADS), fwd_smem<TA>(), st, 1, a, ctx, lse));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

// bf16 mode with a head size that is a multiple of 8: the products run as warp-level tensor-core MMAs
// (attention_train_mma.cu); the fp32 gradient-parity mode (and odd head sizes) use the CUDA-core kernels above
inline bool bwd_on_tensor_cores(const float*, const float*, const float*, const float*, const uint8_t*, const float*, const float*, float*,
                                float*, float*, float*, const float*, int, int, int, int, Drop, cudaStream_t, int*) {
  return false;
}
inline bool bwd_on_tensor_cores(const bf16* qkv, const bf16* pp, const float* u, const float* v, const uint8_t* key_pad, const float* lse,
                                const float* dctx, bf16* dqkv, float* dpp, float* du, float* dv, const float* D, int B, int Tn, int H, int dh,
                                Drop dr, cudaStream_t st, int* status) {
  if (dh % 8 != 0) return false;
  *status = attn_train_bwd_mma(qkv, pp, u, v, key_pad, lse, dctx, dqkv, dpp, du, dv, D, B, Tn, H, dh, dr, st);
  return true;
}

template <typename TA>
int attn_train_bwd(const TA* qkv, const TA* pp, const float* u, const float* v, const uint8_t* key_pad, const TA* ctx, const float* lse,
                   const float* dctx, TA* dqkv, float* dpp, float* du, float* dv, float* scratch, int B, int Tn, int H, int dh, Drop dr,
                   cudaStream_t st) {
  int status = 0;
  CFM_REQUIRE(dh >= 4 && dh <= DHP && dh % 4 == 0, "training attention supports head sizes that are multiples of 4 up to 64");
  static std::atomic<uint64_t> done_q{0}, done_kv{0};
  CFM_TRY(smem_attr_once(done_q, reinterpret_cast<const void*>(attn_bwd_q_kernel<TA>), bwd_smem<TA>()));
  CFM_TRY(smem_attr_once(done_kv, reinterpret_cast<const void*>(attn_bwd_kv_kernel<TA>), bwd_smem<TA>()));
  const int64_t rows = static_cast<int64_t>(B) * Tn * H;
  CFM_CUDA_OK(launch_pdl(attn_delta_kernel<TA>, dim3(static_cast<unsigned>(ceil_div(rows, 8))), dim3(256), 0, st, 1, dctx, ctx, scratch, B, Tn, H, dh));
  CFM_CUDA_OK(cudaGetLastError());
  if (bwd_on_tensor_cores(qkv, pp, u, v, key_pad, lse, dctx, dqkv, dpp, du, dv, scratch, B, Tn, H, dh, dr, st, &status)) return status;
  AttnArgs a{qkv, pp, u, v, key_pad, Tn, H, dh, 1.0f / sqrtf(static_cast<float>(dh)), dr};
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, TQ)), H, B);
  CFM_CUDA_OK(launch_pdl(attn_bwd_q_kernel<TA>, dim3(grid), dim3(THREADS), bwd_smem<TA>(), st, 1, a, lse, scratch, dctx, dqkv, dpp, du, dv));
  CFM_CUDA_OK(cudaGetLastError());
  CFM_CUDA_OK(launch_pdl(attn_bwd_kv_kernel<TA>, dim3(grid), dim3(THREADS), bwd_smem<TA>(), st, 1, a, lse, scratch, dctx, dqkv));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

template int attn_train_fwd<float>(const float*, const float*, const float*, const float*, const uint8_t*, float*, float*, int, int, int, int, Drop, cudaStream_t);
template int attn_train_fwd<bf16>(const bf16*, const bf16*, const float*, const float*, const uint8_t*, bf16*, float*, int, int, int, int, Drop, cudaStream_t);
template int attn_train_bwd<float>(const float*, const float*, const float*, const float*, const uint8_t*, const float*, const float*, const float*, float*, float*, float*, float*, float*, int, int, int, int, Drop, cudaStream_t);
template int attn_train_bwd<bf16>(const bf16*, const bf16*, const float*, const float*, const uint8_t*, const bf16*, const float*, const float*, bf16*, float*, float*, float*, float*, int, int, int, int, Drop, cudaStream_t);

}  // namespace cfm
