// Relative-position multi-head self-attention on tcgen05 tensor cores (head size 64), flash style.
//
//   score[i,j] = ((q_i+u).k_j + (q_i+v).p[T-1+j-i]) / sqrt(dh)         (attention.py:47-56, shift :94-102)
//
// One CTA = 128 query rows of one (batch, head); keys walked in tiles of 64.  Warp roles:
//   warp 0, 3   TMA producers: K tile [64x64] + the 192-row window of the projected positional table this
//               (query tile, key tile) pair can touch (2-D map, negative / past-the-end rows zero-filled), and the
//               V tile [64x64] (3-D maps over q|k|v, rows past T zero-filled), as two 2-stage rings;
//   warp 1      MMA issuer: S = (q+u) K^T (128x64) and PS = (q+v) Pwin^T (128x192) into TMEM, then O += P V
//               (128x64, V is the MN-major B operand) - O ACCUMULATES IN TMEM over all key tiles.  S/PS of tile
//               j+1 are issued as soon as every softmax thread has pulled tile j into registers;
//   warps 4-11  softmax, 256 threads: thread = (query row, half of the tile's keys); warps w+4 and w+8 share
//               TMEM lane quadrant w and take keys 0-31 / 32-63.  The Transformer-XL shift is a per-row column
//               offset 127-i into PS: its warp-uniform part is folded into the tcgen05.ld column address, the
//               per-lane part (31-lane) goes through a private shared-memory row (float4 stores, reads at a
//               lane-dependent offset, both bank-conflict free) - no [T,2T-1] tensor is ever formed.
//               Online softmax in fp32 with exp2 (scale folded into Q) and LAZY rescaling: the reference maximum
//               of a row only moves when the tile maximum exceeds it by more than 2^8, and only then is the row
//               of O in TMEM rescaled (tcgen05.ld / st), so the common tile costs no accumulator traffic at all.
//               P is written as bf16 straight into the MMA's 128B-swizzled A layout.
// Masked keys (and keys past T) get weight exactly 0.
#include <math_constants.h>

#include "train.h"

namespace cfm {
namespace {

#ifndef TRACE_Z
#define TRACE_Z 30      // debug timeline: which batch index's first CTA is traced (30 = a CTA of the third wave at C2)
#endif

constexpr int BQ = 128, BKV = 64, DH = 64;
constexpr int WIN = BQ + BKV;                         // 192 window rows per (query tile, key tile)
constexpr int Q_BYTES = BQ * 128;                     // 16 KB: [128 x 64] bf16, 128B swizzle
constexpr int KV_BYTES = BKV * 128;                   // 8 KB
constexpr int WIN_BYTES = WIN * 128;                  // 24 KB
constexpr int KP_BYTES = KV_BYTES + WIN_BYTES;        // K | window = 32 KB per stage
constexpr int P_BYTES = BQ * 128;                     // 16 KB: [128 x 64 keys] bf16
constexpr int NUM_SM_WARPS = 8, NUM_SM_THREADS = NUM_SM_WARPS * 32;
constexpr int SKEW_LD = 68;                           // floats per private row: 64 used; pitch 68 -> conflict free
constexpr int SKEW_BYTES = NUM_SM_THREADS * SKEW_LD * 4;
constexpr int XCHG_BYTES = 3 * 2 * BQ * 4;            // tile-max exchange (double buffered) + row-sum exchange
constexpr int SMEM_BYTES = 1024 + 2 * Q_BYTES + 2 * KP_BYTES + 2 * KV_BYTES + P_BYTES + SKEW_BYTES + XCHG_BYTES + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "attention shared memory budget");
constexpr int NUM_THREADS = 128 + NUM_SM_THREADS;     // warps 0,3: TMA; 1: MMA; 2: TMEM alloc; 4-11: softmax
constexpr int TM_S = 0, TM_PS = 64, TM_O = 256;       // TMEM columns: S (64) | PS (192) | O (64)
constexpr float RESCALE_THRESHOLD = 8.0f;             // log2 units: P stays below 2^8

struct AttnParams {
  const bf16* qkv;
  const float* ub;
  const float* vb;
  const uint8_t* key_pad;
  bf16* ctx;
  int Tn, H;
  float scale_log2e;
  unsigned long long* trace;     // debug: globaltimer stamps of CTA (0,0,0)
  float* lse;                    // TRAIN: natural-log softmax normaliser of every query row, [B,H,T] (for the backward pass)
  Drop dr;                       // TRAIN: attention-weight dropout (attention.py:67), element ((b*H+h)*T+i)*T+j
};
__device__ __forceinline__ unsigned long long atime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define ATRACE(slot)                                                                                              \
  do {                                                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == TRACE_Z && lane == 0 && (slot) < 64) \
      p.trace[64 + (slot)] = atime_ns();                                                                               \
  } while (0)

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// TRAIN: also writes the softmax normaliser and applies attention-weight dropout to P (the row sums stay undropped)
template <bool TRAIN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
rel_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmPos, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* qu_s = smem;
  uint8_t* qv_s = qu_s + Q_BYTES;
  uint8_t* kp_ring = qv_s + Q_BYTES;                   // 2 x (K | window)
  uint8_t* v_ring = kp_ring + 2 * KP_BYTES;            // 2 x V
  uint8_t* p_s = v_ring + 2 * KV_BYTES;
  float* skew = reinterpret_cast<float*>(p_s + P_BYTES);
  float* xmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(skew) + SKEW_BYTES);   // [2 buffers][2 halves][128]
  float* xsum = xmax + 2 * 2 * BQ;                                                          // [2 halves][128]
  uint64_t* kp_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xmax) + XCHG_BYTES);
  uint64_t* kp_empty = kp_full + 2;
  uint64_t* v_full = kp_empty + 2;
  uint64_t* v_empty = v_full + 2;
  uint64_t* sp_full = v_empty + 2;
  uint64_t* s_read = sp_full + 1;
  uint64_t* p_full = s_read + 1;
  uint64_t* o_done = p_full + 1;
  uint64_t* q_full = o_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * BQ;
  const int Tn = p.Tn, d = p.H * DH;
  const int ntiles = (Tn + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmPos);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kp_full[s], 1);
      mbar_init(&kp_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(sp_full, 1);
    mbar_init(s_read, NUM_SM_WARPS);
    mbar_init(p_full, NUM_SM_WARPS);
    mbar_init(o_done, 1);
    mbar_init(q_full, NUM_SM_WARPS);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_whole_sm(tmem_slot);
  pdl_wait();                                              // q|k|v exist once the projection GEMM has completed

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: K + positional window
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      mbar_wait(&kp_empty[s], ((j >> 1) & 1) ^ 1);
      if (lane == 0) {
        uint8_t* st = kp_ring + s * KP_BYTES;
        const int j0 = j * BKV;
        mbar_expect_tx(&kp_full[s], KP_BYTES);
        tma_load_3d(st, &tmQKV, &kp_full[s], d + h * DH, j0, b);                    // K rows j0..j0+63
        // window entry w <-> table row (T-1 + j0 - i0 - 127) + w
        tma_load_2d(st + KV_BYTES, &tmPos, &kp_full[s], h * DH, Tn - 1 + j0 - i0 - (BQ - 1));
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ TMA producer: V
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      mbar_wait(&v_empty[s], ((j >> 1) & 1) ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&v_full[s], KV_BYTES);
        tma_load_3d(v_ring + s * KV_BYTES, &tmQKV, &v_full[s], 2 * d + h * DH, j * BKV, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV);          // 128 x 64
    constexpr uint32_t idesc_ps = umma_idesc_bf16(BQ, WIN);         // 128 x 192
    constexpr uint32_t idesc_pv = umma_idesc_bf16_bmn(BQ, DH);      // 128 x 64, B (=V) MN-major
    auto issue_sp = [&](int j) {
      const int s = j & 1;
      mbar_wait(&kp_full[s], (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sk = umma_desc_lo(smem_u32(kp_ring + s * KP_BYTES)), sw = sk + (KV_BYTES >> 4);
        const uint32_t squ = umma_desc_lo(smem_u32(qu_s)), sqv = umma_desc_lo(smem_u32(qv_s));
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16_lo(tmem_base + TM_S, squ + k * 2, sk + k * 2, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16_lo(tmem_base + TM_PS, sqv + k * 2, sw + k * 2, idesc_ps, k != 0 ? 1u : 0u);
        tc_commit(sp_full);
        tc_commit(&kp_empty[s]);                                    // K + window may be refilled
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    issue_sp(0);
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      if (j + 1 < ntiles) {
        mbar_wait(s_read, j & 1);                                   // every softmax thread holds tile j in registers
        tc_fence_after();
        issue_sp(j + 1);
      }
      mbar_wait(p_full, j & 1);                                     // P_j written (and O rescaled where needed)
      mbar_wait(&v_full[s], (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sp = umma_desc_lo(smem_u32(p_s));
        const uint32_t sv = umma_desc_lo(smem_u32(v_ring + s * KV_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)                                 // K = 64 keys in steps of 16 (V: 16 rows = 2048 B)
          tc_mma_f16_lo(tmem_base + TM_O, sp + k * 2, sv + k * (2048 >> 4), idesc_pv, (j | k) != 0 ? 1u : 0u);
        tc_commit(o_done);
        tc_commit(&v_empty[s]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax: thread = (query row, key half)
    const int w = warp & 3, grp = (warp - 4) >> 2;       // TMEM lane quadrant (hardware rule: warp % 4), key half
    const int row = w * 32 + lane;
    const int qi = i0 + row;
    const int sw = row & 7;
    if (warp == 4) ATRACE(60);
    // ---- Q operands: (q+u)*c and (q+v)*c as bf16 in the MMA's A layout, c = log2(e)/sqrt(dh); half a row per thread
    {
      const bf16* qrow = p.qkv + (static_cast<int64_t>(b) * Tn + (qi < Tn ? qi : 0)) * (3 * d) + h * DH;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int ch = grp * 4 + c4;
        uint4 raw = make_uint4(0u, 0u, 0u, 0u);
        if (qi < Tn) raw = *reinterpret_cast<const uint4*>(qrow + ch * 8);
        const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
        const float4 ua = __ldg(reinterpret_cast<const float4*>(p.ub + h * DH + ch * 8));
        const float4 ub2 = __ldg(reinterpret_cast<const float4*>(p.ub + h * DH + ch * 8 + 4));
        const float4 va = __ldg(reinterpret_cast<const float4*>(p.vb + h * DH + ch * 8));
        const float4 vb2 = __ldg(reinterpret_cast<const float4*>(p.vb + h * DH + ch * 8 + 4));
        const float uu[8] = {ua.x, ua.y, ua.z, ua.w, ub2.x, ub2.y, ub2.z, ub2.w};
        const float vv[8] = {va.x, va.y, va.z, va.w, vb2.x, vb2.y, vb2.z, vb2.w};
        uint32_t ou[4], ov[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float q0 = __low2float(q2[e]), q1 = __high2float(q2[e]);
          ou[e] = pack_bf16x2((q0 + uu[2 * e]) * p.scale_log2e, (q1 + uu[2 * e + 1]) * p.scale_log2e);
          ov[e] = pack_bf16x2((q0 + vv[2 * e]) * p.scale_log2e, (q1 + vv[2 * e + 1]) * p.scale_log2e);
        }
        const uint32_t off = row * 128 + ((ch ^ sw) << 4);
        *reinterpret_cast<uint4*>(qu_s + off) = make_uint4(ou[0], ou[1], ou[2], ou[3]);
        *reinterpret_cast<uint4*>(qv_s + off) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
      if (warp == 4) ATRACE(61);
    }
    float* myrow = skew + (grp * BQ + row) * SKEW_LD;
    const int shift = 31 - lane;                         // per-lane part of the relative shift
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(w * 32) << 16);
    const uint32_t t_s = tlane + TM_S + 32 * grp;                      // content scores of this thread's 32 keys
    const uint32_t t_ps = tlane + TM_PS + 32 * grp + 32 * (3 - w);     // 64-column window holding their positional scores
    const uint32_t t_o = tlane + TM_O + 32 * grp;                      // the half of the O row this thread rescales
    uint8_t* prow = p_s + row * 128;
    const DropK dk = p.dr.load();                  // TRAIN: dropout seed in registers
    float m_ref = -CUDART_INF_F, l_part = 0.f;

    for (int j = 0; j < ntiles; ++j) {
      // ---- usable-key bitmask of this thread's 32 keys (warp-uniform)
      uint32_t vm = 0xffffffffu;
      {
        const int jk = j * BKV + 32 * grp;
        if (jk + 32 > Tn || p.key_pad != nullptr) {
          const int ja = jk + lane;
          const bool va = ja < Tn && (p.key_pad == nullptr || p.key_pad[static_cast<int64_t>(b) * Tn + ja] == 0);
          vm = __ballot_sync(0xffffffffu, va);
        }
      }
      if (warp == 4) ATRACE(8 * j + 0);
      mbar_wait(sp_full, j & 1);
      tc_fence_after();
      if (warp == 4) ATRACE(8 * j + 1);
      float sc[32];
      float tmax = -CUDART_INF_F;
      {
        uint32_t r0[32], r1[32], s0[32];
        tmem_ld_32x32(t_ps, r0);
        tmem_ld_32x32(t_ps + 32, r1);
        tmem_ld_32x32(t_s, s0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_read);              // the S/PS columns may be overwritten by tile j+1
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          *reinterpret_cast<float4*>(myrow + 4 * e) = make_float4(__uint_as_float(r0[4 * e]), __uint_as_float(r0[4 * e + 1]),
                                                                  __uint_as_float(r0[4 * e + 2]), __uint_as_float(r0[4 * e + 3]));
          *reinterpret_cast<float4*>(myrow + 32 + 4 * e) = make_float4(__uint_as_float(r1[4 * e]), __uint_as_float(r1[4 * e + 1]),
                                                                       __uint_as_float(r1[4 * e + 2]), __uint_as_float(r1[4 * e + 3]));
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          float v = __uint_as_float(s0[jj]) + myrow[shift + jj];
          if (((vm >> jj) & 1u) == 0) v = -CUDART_INF_F;
          sc[jj] = v;
          tmax = fmaxf(tmax, v);
        }
      }
      if (warp == 4) ATRACE(8 * j + 2);
      // ---- tile maximum of the whole row: exchange with the thread that holds the other 32 keys
      float* xm = xmax + (j & 1) * 2 * BQ;
      xm[grp * BQ + row] = tmax;
      named_bar_sync(1 + w, 64);
      const float t_row = fmaxf(tmax, xm[(grp ^ 1) * BQ + row]);
      // ---- lazy rescaling: move the reference maximum only when the tile exceeds it by more than the threshold
      const float m_new = (t_row > m_ref + RESCALE_THRESHOLD) ? t_row : m_ref;
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);                  // P.V of tile j-1 retired: P buffer free, O stable
        if (__any_sync(0xffffffffu, m_new != m_ref)) {
          tc_fence_after();
          const float f = ex2f(m_ref - m_new);           // 1 for rows whose reference did not move
          uint32_t o[32];
          tmem_ld_32x32(t_o, o);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
          tmem_st_32x32(t_o, o);
          tmem_st_wait();
          l_part *= f;
        }
      }
      m_ref = m_new;
      const float m_use = (m_ref == -CUDART_INF_F) ? 0.f : m_ref;
      if (warp == 4) ATRACE(8 * j + 3);
      // ---- P = exp2(s - m) as bf16 into the A layout of the P.V MMA (this thread: 4 chunks of 8 keys)
      float psum = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = ex2f(sc[c4 * 8 + 2 * e] - m_use), p1 = ex2f(sc[c4 * 8 + 2 * e + 1] - m_use);
          psum += p0 + p1;
          if (TRAIN) {
            const int64_t e0 = (static_cast<int64_t>(b * p.H + h) * Tn + qi) * Tn + j * BKV + 32 * grp + c4 * 8 + 2 * e;
            o[e] = pack_bf16x2(p0 * dk.factor(e0), p1 * dk.factor(e0 + 1));
          } else {
            o[e] = pack_bf16x2(p0, p1);
          }
        }
        *reinterpret_cast<uint4*>(prow + (((grp * 4 + c4) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      l_part += psum;
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == 4) ATRACE(8 * j + 4);
    }
    // ---- epilogue: O / l, this thread's 32 of the 64 output columns
    xsum[grp * BQ + row] = l_part;
    mbar_wait(o_done, (ntiles - 1) & 1);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld_32x32(t_o, o);
    tmem_ld_wait();
    named_bar_sync(1 + w, 64);
    const float l_tot = l_part + xsum[(grp ^ 1) * BQ + row];
    const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
    if (TRAIN && grp == 0 && qi < Tn)                    // scores are in log2 units: ln sum e^s = m ln2 + ln l
      p.lse[(static_cast<int64_t>(b) * p.H + h) * Tn + qi] = (m_ref == -CUDART_INF_F ? 0.f : m_ref) * 0.6931471805599453f + __logf(fmaxf(l_tot, 1e-30f));
    if (qi < Tn) {
      bf16* orow = p.ctx + (static_cast<int64_t>(b) * Tn + qi) * d + h * DH + 32 * grp;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint4 v4;
        v4.x = pack_bf16x2(__uint_as_float(o[ch * 8 + 0]) * inv, __uint_as_float(o[ch * 8 + 1]) * inv);
        v4.y = pack_bf16x2(__uint_as_float(o[ch * 8 + 2]) * inv, __uint_as_float(o[ch * 8 + 3]) * inv);
        v4.z = pack_bf16x2(__uint_as_float(o[ch * 8 + 4]) * inv, __uint_as_float(o[ch * 8 + 5]) * inv);
        v4.w = pack_bf16x2(__uint_as_float(o[ch * 8 + 6]) * inv, __uint_as_float(o[ch * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + ch * 8) = v4;
      }
    }
    if (warp == 4) ATRACE(62);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

extern unsigned long long* g_gemm_trace;

static int launch_attention_tc(const bf16* qkv, const bf16* pos, int ldp, const float* u, const float* v, const uint8_t* key_pad, bf16* ctx,
                               float* lse, Drop dr, bool train, int B, int Tn, int H, int dh, cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(dh == DH, "tcgen05 attention is specialised for head size 64");
  CFM_REQUIRE(B > 0 && Tn > 0 && H > 0 && H <= 65535 && B <= 65535, "bad attention shape");
  CFM_REQUIRE(ldp % 8 == 0 && ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(pos) |
                                reinterpret_cast<uintptr_t>(ctx)) & 15) == 0, "attention operands must be 16-byte aligned");
  const int d = H * dh;
  CUtensorMap tmQKV, tmPos;
  {
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(3 * d), static_cast<cuuint64_t>(Tn), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(3 * d) * 2, static_cast<cuuint64_t>(Tn) * 3 * d * 2};
    cuuint32_t box[3] = {DH, BKV, 1};
    CFM_TRY(encode_tmap(&tmQKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv, dims, strides, box, "attention qkv"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(ldp), static_cast<cuuint64_t>(2 * Tn - 1)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldp) * 2};
    cuuint32_t box[2] = {DH, WIN};
    CFM_TRY(encode_tmap(&tmPos, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pos, dims, strides, box, "attention pos"));
  }
  static std::atomic<uint64_t> attr_done{0}, attr_done_t{0};
  AttnParams p{qkv, u, v, key_pad, ctx, Tn, H, 1.4426950408889634f / sqrtf(static_cast<float>(dh)), g_gemm_trace, lse, dr};
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, BQ)), static_cast<unsigned>(H), static_cast<unsigned>(B));
  if (train) {
    CFM_TRY(smem_attr_once(attr_done_t, reinterpret_cast<const void*>(rel_attention_tc_kernel<true>), SMEM_BYTES));
    CFM_CUDA_OK(launch_pdl(rel_attention_tc_kernel<true>, grid, dim3(NUM_THREADS), SMEM_BYTES, stream, 1, tmQKV, tmPos, p));
  } else {
    CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(rel_attention_tc_kernel<false>), SMEM_BYTES));
    CFM_CUDA_OK(launch_pdl(rel_attention_tc_kernel<false>, grid, dim3(NUM_THREADS), SMEM_BYTES, stream, 1, tmQKV, tmPos, p));
  }
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

int rel_attention_bf16_tc(const bf16* qkv, const bf16* pos, int ldp, const float* u, const float* v,
                          const uint8_t* key_pad, bf16* ctx, int B, int Tn, int H, int dh, cudaStream_t stream) {
  return launch_attention_tc(qkv, pos, ldp, u, v, key_pad, ctx, nullptr, Drop{}, false, B, Tn, H, dh, stream);
}
// training forward: the same kernel, also emitting the log-sum-exp of every row and applying attention-weight dropout
int rel_attention_bf16_tc_train(const bf16* qkv, const bf16* pos, int ldp, const float* u, const float* v, const uint8_t* key_pad,
                                bf16* ctx, float* lse, Drop dr, int B, int Tn, int H, int dh, cudaStream_t stream) {
  CFM_REQUIRE(lse != nullptr, "training attention needs the log-sum-exp buffer");
  return launch_attention_tc(qkv, pos, ldp, u, v, key_pad, ctx, lse, dr, true, B, Tn, H, dh, stream);
}

}  // namespace cfm
