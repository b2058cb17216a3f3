// Relative-position multi-head self-attention for d_model = 256, head size 64, with the SINUSOIDAL positional
// term factorised so that no relative shift is needed at all (attention.py:47-56, :94-102; position.py:11-28).
//
// The reference adds  pos[i,j] = (q_i+v) . Wp_h PE(i-j)  where PE(r) = [sin(r w_k), cos(r w_k)]_k  (rows of the table
// built by RelativePositionalEncoding, gathered by _relative_shift).  With (a_ik, b_ik) = Wp_h^T (q_i+v) split into
// the sin / cos input columns of pos_proj, the angle-difference identities give
//
//     pos[i,j] = sum_k  A_ik cos(j w_k) + B_ik sin(j w_k),      A = a sin(i w) + b cos(i w),   B = b sin(i w) - a cos(i w)
//
// i.e. a plain [T x 256] . [256 x T] product of a rotated query Q'_i = [A_i | B_i] with the ABSOLUTE key table
// E_j = [cos(j w) | sin(j w)] (shared by every layer, head and utterance) - the Transformer-XL shift, the [T,2T-1]
// score tensor and the projected [2T-1,d] table all disappear, and the positional scores accumulate into the same
// TMEM columns as the content scores.  (The pos_proj bias adds a per-query constant that softmax cancels, SURVEY.md
// 5.7-5.)  bf16 rounding now falls on Q' and E instead of on the projected table; measured score error is the same
// as the reference's own autocast path.
//
// One CTA = 128 query rows of one (batch, head); keys in tiles of 64.
//   warp 0, 3   TMA producers: pos_proj slice of the head (once), then per key tile K [64x64] + E [64x256], and V;
//   warp 1      MMA issuer: prologue ab = (q+v) Wp_h (128x256 into TMEM); per tile S = (q+u) K^T + Q' E^T (K = 64 + 256,
//               double-buffered in TMEM) and O += P V, O accumulating in TMEM over all key tiles;
//   warps 4-11  256 softmax threads, thread = (query row, half of the tile's keys): rotate ab into Q' (prologue), then
//               online softmax in fp32 with exp2 (scale folded into Q) and lazy rescaling of O (only when the row
//               maximum grows by more than 2^8); P goes to shared memory as the bf16 A operand of P V.
// Masked keys (and keys past T) get weight exactly 0.
#include <cuda_fp16.h>
#include <math_constants.h>

#include "ops.h"

namespace cfm {
namespace {

#ifndef TRACE_Z
#define TRACE_Z 30      // debug timeline: which batch index's first CTA is traced (30 = a CTA of the third wave at C2)
#endif

constexpr int D = 256, DH = 64, NF = D / 2;           // model width, head size, number of sinusoid frequencies
constexpr int BQ = 128, BKV = 64;
constexpr int Q_BYTES = BQ * 128;                     // 16 KB: [128 x 64] bf16, 128B swizzle
constexpr int QP_BYTES = (D / 64) * Q_BYTES;          // 64 KB: Q' as 4 k-blocks
constexpr int KV_BYTES = BKV * 128;                   // 8 KB
constexpr int E_BYTES = (D / 64) * KV_BYTES;          // 32 KB: E tile as 4 k-blocks of [64 keys x 64]
constexpr int KP_BYTES = KV_BYTES + E_BYTES;          // K | E = 40 KB per stage
constexpr int NUM_SM_WARPS = 8, NUM_SM_THREADS = NUM_SM_WARPS * 32;
constexpr int XCHG_BYTES = 3 * 2 * BQ * 4;            // tile-max exchange (double buffered) + row-sum exchange
constexpr int SMEM_BYTES = 1024 + Q_BYTES + QP_BYTES + Q_BYTES + 2 * KP_BYTES + 2 * KV_BYTES + XCHG_BYTES + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "attention shared memory budget");
static_assert(D * 128 <= E_BYTES, "the pos_proj slice must fit the E region of a ring stage");
constexpr int NUM_THREADS = 128 + NUM_SM_THREADS;     // warps 0,3: TMA; 1: MMA; 2: TMEM alloc; 4-11: softmax
constexpr int TM_S = 0, TM_O = 128, TM_AB = 256;      // TMEM columns: S0 | S1 (64 each), O (64), ab (256)
constexpr float RESCALE_THRESHOLD = 8.0f;             // log2 units: P stays below 2^8

struct FactParams {
  const bf16* qkv;
  const float* ub;
  const float* vb;
  const uint8_t* key_pad;
  const __half* rot;    // [T][256] fp16: cos(i w_k) | sin(i w_k) (query rotation; abs error 2^-12, far below bf16 Q')
  bf16* ctx;
  int Tn, H;
  float scale_log2e;
  unsigned long long* trace;     // debug: globaltimer stamps of CTA (0,0,0)
};
__device__ __forceinline__ unsigned long long ftime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define FTRACE(slot)                                                                                              \
  do {                                                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == TRACE_Z && lane == 0 && (slot) < 64) \
      p.trace[64 + (slot)] = ftime_ns();                                                                               \
  } while (0)

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
rel_attention_fact_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmE,
                          const __grid_constant__ CUtensorMap tmWp, const __grid_constant__ CUtensorMap tmRot, FactParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* qu_s = smem;
  uint8_t* qp_s = qu_s + Q_BYTES;                      // Q' (4 k-blocks)
  uint8_t* p_s = qp_s + QP_BYTES;                      // (q+v) during the prologue, P afterwards
  uint8_t* kp_ring = p_s + Q_BYTES;                    // 2 x (K | E)
  uint8_t* v_ring = kp_ring + 2 * KP_BYTES;            // 2 x V
  float* xmax = reinterpret_cast<float*>(v_ring + 2 * KV_BYTES);    // [2 buffers][2 halves][128]
  float* xsum = xmax + 2 * 2 * BQ;                                   // [2 halves][128]
  uint64_t* kp_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xmax) + XCHG_BYTES);
  uint64_t* kp_empty = kp_full + 2;
  uint64_t* v_full = kp_empty + 2;
  uint64_t* v_empty = v_full + 2;
  uint64_t* s_full = v_empty + 2;       // [2]
  uint64_t* s_read = s_full + 2;        // [2]
  uint64_t* p_full = s_read + 2;
  uint64_t* o_done = p_full + 1;
  uint64_t* q_full = o_done + 1;
  uint64_t* ab_full = q_full + 1;
  uint64_t* qp_full = ab_full + 1;
  uint64_t* rot_full = qp_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rot_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * BQ;
  const int Tn = p.Tn, d = p.H * DH;
  const int ntiles = (Tn + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmE);
    tma_prefetch_desc(&tmWp);
    tma_prefetch_desc(&tmRot);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kp_full[s], 1);
      mbar_init(&kp_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_read[s], NUM_SM_WARPS);
    }
    mbar_init(p_full, NUM_SM_WARPS);
    mbar_init(o_done, 1);
    mbar_init(q_full, NUM_SM_WARPS);
    mbar_init(ab_full, 1);
    mbar_init(qp_full, NUM_SM_WARPS);
    mbar_init(rot_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_whole_sm(tmem_slot);

  // ring stage 1 is filled once more than stage 0: its first fill is the pos_proj slice of the prologue
  auto fill_index = [](int j) { return (j & 1) ? (j >> 1) + 1 : (j >> 1); };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: pos_proj slice, then K + E tiles
    if (lane == 0) {
      mbar_expect_tx(&kp_full[1], D * 128);
      tma_load_2d(kp_ring + KP_BYTES + KV_BYTES, &tmWp, &kp_full[1], 0, h * D);       // [256 rows (m') x 64 (c)]
      // rotation factors of the 128 query rows, laid out exactly like Q' (k-block kb = table columns 64*kb .. +63):
      // every softmax thread later overwrites, in place, the two rows it reads
      mbar_expect_tx(rot_full, QP_BYTES);
#pragma unroll
      for (int kb = 0; kb < D / 64; ++kb) tma_load_2d(qp_s + kb * Q_BYTES, &tmRot, rot_full, kb * 64, i0);
    }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1, n = fill_index(j);
      mbar_wait(&kp_empty[s], (n & 1) ^ 1);
      if (lane == 0) {
        uint8_t* st = kp_ring + s * KP_BYTES;
        const int j0 = j * BKV;
        mbar_expect_tx(&kp_full[s], KP_BYTES);
        tma_load_3d(st, &tmQKV, &kp_full[s], d + h * DH, j0, b);                      // K rows j0..j0+63
#pragma unroll
        for (int kb = 0; kb < D / 64; ++kb) tma_load_2d(st + KV_BYTES + kb * KV_BYTES, &tmE, &kp_full[s], kb * 64, j0);
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
    constexpr uint32_t idesc_ab = umma_idesc_bf16(BQ, D);           // 128 x 256
    constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV);          // 128 x 64
    constexpr uint32_t idesc_pv = umma_idesc_bf16_bmn(BQ, DH);      // 128 x 64, B (=V) MN-major
    // ---- prologue: ab = (q+v) Wp_h   [128 x 256], K = 64
    mbar_wait(q_full, 0);
    mbar_wait(&kp_full[1], 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t sa = umma_desc_lo(smem_u32(p_s)), sb = umma_desc_lo(smem_u32(kp_ring + KP_BYTES + KV_BYTES));
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_f16_lo(tmem_base + TM_AB, sa + k * 2, sb + k * 2, idesc_ab, k != 0 ? 1u : 0u);
      tc_commit(ab_full);
      tc_commit(&kp_empty[1]);
    }
    __syncwarp();
    auto issue_s = [&](int j) {
      const int s = j & 1, n = fill_index(j);
      mbar_wait(&kp_full[s], n & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sk = umma_desc_lo(smem_u32(kp_ring + s * KP_BYTES)), se = sk + (KV_BYTES >> 4);
        const uint32_t squ = umma_desc_lo(smem_u32(qu_s)), sqp = umma_desc_lo(smem_u32(qp_s));
        const uint32_t ts = tmem_base + TM_S + s * BKV;
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16_lo(ts, squ + k * 2, sk + k * 2, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int kb = 0; kb < D / 64; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16_lo(ts, sqp + kb * (Q_BYTES >> 4) + k * 2, se + kb * (KV_BYTES >> 4) + k * 2, idesc_s, 1u);
        tc_commit(&s_full[s]);
        tc_commit(&kp_empty[s]);                                    // K + E may be refilled
      }
      __syncwarp();
    };
    mbar_wait(qp_full, 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      if (j + 1 < ntiles) {
        if (j >= 1) {
          mbar_wait(&s_read[(j + 1) & 1], ((j - 1) >> 1) & 1);      // tile j-1 left this S buffer
          tc_fence_after();
        }
        issue_s(j + 1);
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
    const int w = warp & 3, grp = (warp - 4) >> 2;       // TMEM lane quadrant (hardware rule: warp % 4), half
    const int row = w * 32 + lane;
    const int qi = i0 + row;
    const int sw = row & 7;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(w * 32) << 16);
    if (warp == 4) FTRACE(0);
    // ---- Q operands: (q+u)*c and (q+v)*c as bf16 in the MMA's A layout, c = log2(e)/sqrt(dh); half a row per thread
    {
      const bf16* qrow = p.qkv + (static_cast<int64_t>(b) * Tn + (qi < Tn ? qi : 0)) * (3 * d) + h * DH;
      uint4 raw[4];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        raw[c4] = make_uint4(0u, 0u, 0u, 0u);
        if (qi < Tn) raw[c4] = *reinterpret_cast<const uint4*>(qrow + (grp * 4 + c4) * 8);
      }
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int ch = grp * 4 + c4;
        const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&raw[c4]);
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
        *reinterpret_cast<uint4*>(p_s + off) = make_uint4(ov[0], ov[1], ov[2], ov[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
    }
    // ---- rotate ab into Q' = [A | B]; this thread: frequencies 64*grp .. +63 -> k-blocks grp (A) and 2+grp (B)
    {
      if (warp == 4) FTRACE(1);
      mbar_wait(ab_full, 0);
      tc_fence_after();
      if (warp == 4) FTRACE(2);
      mbar_wait(rot_full, 0);
      uint8_t* arow = qp_s + grp * Q_BYTES + row * 128;          // holds cos(i w_k), k = 64*grp .. +63; becomes A
      uint8_t* brow = qp_s + (2 + grp) * Q_BYTES + row * 128;    // holds sin(i w_k);                    becomes B
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {                   // 16 frequencies at a time
        const int k0 = grp * 64 + cc * 16;
        uint32_t ar[16], br[16];
        tmem_ld_32x16(tlane + TM_AB + k0, ar);
        tmem_ld_32x16(tlane + TM_AB + NF + k0, br);
        uint4* pa0 = reinterpret_cast<uint4*>(arow + (((2 * cc) ^ sw) << 4));
        uint4* pa1 = reinterpret_cast<uint4*>(arow + (((2 * cc + 1) ^ sw) << 4));
        uint4* pb0 = reinterpret_cast<uint4*>(brow + (((2 * cc) ^ sw) << 4));
        uint4* pb1 = reinterpret_cast<uint4*>(brow + (((2 * cc + 1) ^ sw) << 4));
        const uint4 c0 = *pa0, c1 = *pa1, s0 = *pb0, s1 = *pb1;
        const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const uint32_t sw8[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        tmem_ld_wait();
        uint32_t oa[8], ob[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float2 cs = __half22float2(*reinterpret_cast<const __half2*>(&cw[e]));
          const float2 sn = __half22float2(*reinterpret_cast<const __half2*>(&sw8[e]));
          const float a0 = __uint_as_float(ar[2 * e]), a1 = __uint_as_float(ar[2 * e + 1]);
          const float b0 = __uint_as_float(br[2 * e]), b1 = __uint_as_float(br[2 * e + 1]);
          oa[e] = pack_bf16x2(a0 * sn.x + b0 * cs.x, a1 * sn.y + b1 * cs.y);
          ob[e] = pack_bf16x2(b0 * sn.x - a0 * cs.x, b1 * sn.y - a1 * cs.y);
        }
        *pa0 = make_uint4(oa[0], oa[1], oa[2], oa[3]);
        *pa1 = make_uint4(oa[4], oa[5], oa[6], oa[7]);
        *pb0 = make_uint4(ob[0], ob[1], ob[2], ob[3]);
        *pb1 = make_uint4(ob[4], ob[5], ob[6], ob[7]);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(qp_full);
      if (warp == 4) FTRACE(3);
    }
    const uint32_t t_o = tlane + TM_O + 32 * grp;        // the half of the O row this thread rescales
    uint8_t* prow = p_s + row * 128;
    float m_ref = -CUDART_INF_F, l_part = 0.f;

    // usable-key bitmask of this thread's 32 keys of tile j (warp-uniform)
    auto key_mask = [&](int j) -> uint32_t {
      const int jk = j * BKV + 32 * grp;
      if (jk + 32 <= Tn && p.key_pad == nullptr) return 0xffffffffu;
      const int ja = jk + lane;
      const bool va = ja < Tn && (p.key_pad == nullptr || p.key_pad[static_cast<int64_t>(b) * Tn + ja] == 0);
      return __ballot_sync(0xffffffffu, va);
    };
    // pull this thread's 32 scores of tile j out of TMEM and hand the S buffer back
    auto load_scores = [&](int j, uint32_t (&s0)[32]) {
      const int buf = j & 1;
      mbar_wait(&s_full[buf], (j >> 1) & 1);
      tc_fence_after();
      tmem_ld_32x32(tlane + TM_S + buf * BKV + 32 * grp, s0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_read[buf]);          // this S buffer may be overwritten by tile j+2
    };
    auto mask_and_max = [&](uint32_t (&s0)[32], uint32_t vm) -> float {
      if (vm != 0xffffffffu) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          if (((vm >> jj) & 1u) == 0) s0[jj] = __float_as_uint(-CUDART_INF_F);
      }
      float m0 = max3f(__uint_as_float(s0[0]), __uint_as_float(s0[1]), __uint_as_float(s0[2]));
      float m1 = max3f(__uint_as_float(s0[3]), __uint_as_float(s0[4]), __uint_as_float(s0[5]));
#pragma unroll
      for (int jj = 6; jj < 30; jj += 4) {
        m0 = max3f(m0, __uint_as_float(s0[jj]), __uint_as_float(s0[jj + 1]));
        m1 = max3f(m1, __uint_as_float(s0[jj + 2]), __uint_as_float(s0[jj + 3]));
      }
      return max3f(m0, m1, fmaxf(__uint_as_float(s0[30]), __uint_as_float(s0[31])));
    };

    uint32_t sc[32];
    if (warp == 4) FTRACE(8);
    load_scores(0, sc);
    float tmax = mask_and_max(sc, key_mask(0));
    for (int j = 0; j < ntiles; ++j) {
      const int buf = j & 1;
      if (warp == 4) FTRACE(9 + 4 * j);
      // ---- tile maximum of the whole row: exchange with the thread that holds the other 32 keys
      float* xm = xmax + buf * 2 * BQ;
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
      if (warp == 4) FTRACE(10 + 4 * j);
      // ---- P = exp2(s - m) as bf16 into the A layout of the P.V MMA (this thread: 4 chunks of 8 keys)
      float psum0 = 0.f, psum1 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = ex2f(__uint_as_float(sc[c4 * 8 + 2 * e]) - m_use);
          const float p1 = ex2f(__uint_as_float(sc[c4 * 8 + 2 * e + 1]) - m_use);
          psum0 += p0;
          psum1 += p1;
          o[e] = pack_bf16x2(p0, p1);
        }
        *reinterpret_cast<uint4*>(prow + (((grp * 4 + c4) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      l_part += psum0 + psum1;
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == 4) FTRACE(11 + 4 * j);
      if (j + 1 < ntiles) {
        const uint32_t vm_next = key_mask(j + 1);
        load_scores(j + 1, sc);
        tmax = mask_and_max(sc, vm_next);
      }
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
    if (warp == 4) FTRACE(4);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// E[t] = [cos(t w_k) | sin(t w_k)], k < d/2, as bf16 (B operand of the positional product) and fp16 (query rotation)
__global__ void pos_fact_tables_kernel(const float* __restrict__ div_term, bf16* __restrict__ e16, __half* __restrict__ r16,
                                       int Tn, int nf) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Tn * nf) return;
  const int t = idx / nf, k = idx - t * nf;
  float s, c;
  sincosf(static_cast<float>(t) * div_term[k], &s, &c);
  const int64_t base = static_cast<int64_t>(t) * 2 * nf;
  e16[base + k] = __float2bfloat16_rn(c);
  e16[base + nf + k] = __float2bfloat16_rn(s);
  r16[base + k] = __float2half_rn(c);
  r16[base + nf + k] = __float2half_rn(s);
}

// WpT[h][m'][c] = Wp[h*dh + c][m(m')], m(m') = 2 m' (sin input columns) for m' < d/2, else 2 (m' - d/2) + 1 (cos)
__global__ void pack_pos_fact_kernel(const float* __restrict__ wp, bf16* __restrict__ out, int d, int dh) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(d) * d) return;
  const int c = static_cast<int>(idx % dh);
  const int mp = static_cast<int>((idx / dh) % d);
  const int h = static_cast<int>(idx / (static_cast<int64_t>(dh) * d));
  const int m = mp < d / 2 ? 2 * mp : 2 * (mp - d / 2) + 1;
  out[idx] = __float2bfloat16_rn(wp[static_cast<int64_t>(h * dh + c) * d + m]);
}

}  // namespace

extern unsigned long long* g_gemm_trace;

bool rel_attention_fact_supported(int d, int dh) { return d == D && dh == DH; }

int pos_fact_tables(const float* div_term, bf16* e16, void* r16, int Tn, int d, cudaStream_t stream) {
  CFM_REQUIRE(Tn > 0 && d > 0 && d % 2 == 0, "bad positional table shape");
  const int n = Tn * (d / 2);
  pos_fact_tables_kernel<<<(n + 255) / 256, 256, 0, stream>>>(div_term, e16, static_cast<__half*>(r16), Tn, d / 2);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

int pack_pos_fact(const float* wp, bf16* out, int d, int dh, cudaStream_t stream) {
  const int64_t n = static_cast<int64_t>(d) * d;
  pack_pos_fact_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(wp, out, d, dh);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

int rel_attention_bf16_fact(const bf16* qkv, const bf16* wpT, const bf16* e16, const void* r16, const float* u,
                            const float* v, const uint8_t* key_pad, bf16* ctx, int B, int Tn, int H, int dh,
                            cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(dh == DH && H * dh == D, "factorised attention is specialised for d_model 256, head size 64");
  CFM_REQUIRE(B > 0 && Tn > 0 && B <= 65535, "bad attention shape");
  CFM_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(wpT) | reinterpret_cast<uintptr_t>(e16) |
                reinterpret_cast<uintptr_t>(r16) | reinterpret_cast<uintptr_t>(ctx)) & 15) == 0,
              "attention operands must be 16-byte aligned");
  CUtensorMap tmQKV, tmE, tmWp;
  {
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(3 * D), static_cast<cuuint64_t>(Tn), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(3 * D) * 2, static_cast<cuuint64_t>(Tn) * 3 * D * 2};
    cuuint32_t box[3] = {DH, BKV, 1};
    CFM_TRY(encode_tmap(&tmQKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qkv, dims, strides, box, "attention qkv"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(Tn)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * 2};
    cuuint32_t box[2] = {64, BKV};
    CFM_TRY(encode_tmap(&tmE, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, e16, dims, strides, box, "attention E"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(DH), static_cast<cuuint64_t>(H) * D};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(DH) * 2};
    cuuint32_t box[2] = {DH, D};
    CFM_TRY(encode_tmap(&tmWp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wpT, dims, strides, box, "attention pos_proj"));
  }
  CUtensorMap tmRot;
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(Tn)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * 2};
    cuuint32_t box[2] = {64, BQ};
    CFM_TRY(encode_tmap(&tmRot, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, r16, dims, strides, box, "attention rotation table"));
  }
  static bool attr_done = false;
  if (!attr_done) {
    CFM_CUDA_OK(cudaFuncSetAttribute(rel_attention_fact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  FactParams p{qkv, u, v, key_pad, static_cast<const __half*>(r16), ctx, Tn, H, 1.4426950408889634f / sqrtf(static_cast<float>(dh)), g_gemm_trace};
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, BQ)), static_cast<unsigned>(H), static_cast<unsigned>(B));
  rel_attention_fact_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQKV, tmE, tmWp, tmRot, p);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
