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
// The query block of the fused q|k|v projection is packed for this kernel as  Qs = c (q + u),  c = log2(e)/sqrt(dh)
// (content bias and softmax scale folded into the projection's weight and bias, pack_q_fact), so the query tile is used
// exactly as TMA delivers it; the positional bias enters as the per-head constant  c (v - u) Wp_h  added to ab.
//
// One CTA = 128 query rows of one (batch, head); keys in tiles of 64.  Every A operand of the score / context MMAs
// lives in TENSOR MEMORY (Qs, Q', P): with A in shared memory an M=128, N=64 step needs 192 B/clk of operand reads, above
// the 128 B/clk port (measured 85 clk per step); with A in TMEM only B (K, E, V) is read and a step takes its 32 clk.
//   warp 0, 3   TMA producers: Qs tile, pos_proj slice of the head, rotation factors (once), then per key tile
//               K [64x64] + E [64x256] through a 3-stage ring, and V;
//   warp 1      MMA issuer: prologue ab = Qs Wp_h (128x256 into TMEM); per tile S = Qs K^T + Q' E^T (K = 64 + 256,
//               double-buffered in TMEM) and O += P V, O accumulating in TMEM over all key tiles;
//   warps 4-11  256 softmax threads, thread = (query row, half of the tile's keys): copy Qs into TMEM and rotate ab
//               into Q' (prologue), then online softmax in fp32 with exp2 and lazy rescaling of O (only when the row
//               maximum grows by more than 2^8); each key half keeps its own maximum, sum and O accumulator (merged in
//               the epilogue), so the two threads of a row never synchronise inside the key loop.  P overwrites, as bf16,
//               columns of its own S buffer in TMEM and is the TMEM-resident A operand of P V.
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
constexpr int ROT_BYTES = (D / 64) * Q_BYTES;         // 64 KB: rotation factors of the 128 query rows, 4 k-blocks
constexpr int KV_BYTES = BKV * 128;                   // 8 KB
constexpr int E_BYTES = (D / 64) * KV_BYTES;          // 32 KB: E tile as 4 k-blocks of [64 keys x 64]
constexpr int KP_BYTES = KV_BYTES + E_BYTES;          // K | E = 40 KB per stage
constexpr int NS_RING = 3;                            // K | E stages with their own shared memory
constexpr int NS = 4;                                 // K | E stages in all: the last one reuses the rotation-factor buffer,
                                                      // dead after the prologue
constexpr int WP_STAGE = 2;                           // stage whose first fill is the pos_proj slice of the prologue
constexpr int NSB = 3;                                // S buffers in TMEM: S runs two tiles ahead of the softmax
constexpr int NUM_SM_WARPS = 8, NUM_SM_THREADS = NUM_SM_WARPS * 32;
constexpr int XCHG_BYTES = 3 * 2 * BQ * 4;            // tile-max exchange (double buffered) + row-sum exchange
constexpr int SMEM_BYTES = 1024 + Q_BYTES + ROT_BYTES + NS_RING * KP_BYTES + 2 * KV_BYTES + XCHG_BYTES + 256;
static_assert(KP_BYTES <= ROT_BYTES, "the extra ring stage lives in the rotation-factor buffer");
static_assert(SMEM_BYTES <= 227 * 1024, "attention shared memory budget");
static_assert(D * 128 <= E_BYTES, "the pos_proj slice must fit the E region of a ring stage");
constexpr int NUM_THREADS = 128 + NUM_SM_THREADS;     // warps 0,3: TMA; 1: MMA; 2: TMEM alloc; 4-11: softmax
// TMEM columns.  Prologue: ab (256 fp32 columns).  Main loop, over the dead ab: S0 | S1 | S2 (64 each; each key half of
// P_j overwrites, as bf16 pairs, the first 16 columns of its own half of S_j) and one O accumulator per key half (64
// each).  Throughout: Q' (256 bf16 = 128 columns) and Qs (64 bf16 = 32).  480 of the 512 columns.
// Three S buffers because P_j.V must precede S_{j+NSB} in the (in-order) tensor pipe: with two, S_{j+2} could only be
// issued after P_j was ready and the softmax warps idled ~0.25 us per tile waiting for the next scores.
constexpr int TM_AB = 0, TM_S = 0, TM_O = NSB * 64, TM_QP = TM_O + 128, TM_Q = TM_QP + 128;
static_assert(TM_Q + 32 <= 512 && TM_QP >= 256, "TMEM budget; Q' must not alias ab");
constexpr float RESCALE_THRESHOLD = 8.0f;             // log2 units: P stays below 2^8

struct FactParams {
  const float* pos_const;   // [H][256]: c (v_h - u_h) Wp_h in the packed column order of wpT
  const uint8_t* key_pad;
  bf16* ctx;
  int Tn, H;
  uint32_t tmem_zero;       // always 0: the TMEM base as a kernel parameter keeps every TMEM operand address a uniform
                            // register + immediate (a literal base makes ptxas build each one in a vector register
                            // and move it over with R2UR: ~56 clk per issued MMA instead of < 32)
  unsigned long long* trace;     // debug: globaltimer stamps of one CTA
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
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* q_s = smem;                                 // Qs tile as TMA delivers it
  uint8_t* rot_s = q_s + Q_BYTES;                      // rotation factors, laid out like a [128 x 256] K-major operand
  uint8_t* kp_ring = rot_s + ROT_BYTES;                // NS x (K | E)
  uint8_t* v_ring = kp_ring + NS_RING * KP_BYTES;      // 2 x V
  float* xmax = reinterpret_cast<float*>(v_ring + 2 * KV_BYTES);    // [2 buffers][2 halves][128]
  float* xsum = xmax + 2 * 2 * BQ;                                   // [2 halves][128]
  uint64_t* kp_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xmax) + XCHG_BYTES);
  uint64_t* kp_empty = kp_full + NS;
  uint64_t* v_full = kp_empty + NS;
  uint64_t* v_empty = v_full + 2;
  uint64_t* s_full = v_empty + 2;       // [NSB]
  uint64_t* p_full = s_full + NSB;
  uint64_t* o_done = p_full + 1;
  uint64_t* q_full = o_done + 1;
  uint64_t* ab_full = q_full + 1;
  uint64_t* qp_full = ab_full + 1;
  uint64_t* rot_full = qp_full + 1;
  uint64_t* o_final = rot_full + 1;     // completes once, after the last P.V (a parity wait on o_done could alias)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_final + 1);

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
    for (int s = 0; s < NS; ++s) {
      mbar_init(&kp_full[s], 1);
      mbar_init(&kp_empty[s], 1);
    }
    for (int s = 0; s < NSB; ++s) mbar_init(&s_full[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(p_full, NUM_SM_WARPS);
    mbar_init(o_done, 1);
    mbar_init(q_full, 1);
    mbar_init(ab_full, 1);
    mbar_init(qp_full, NUM_SM_WARPS);
    mbar_init(rot_full, 1);
    mbar_init(o_final, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_whole_sm(tmem_slot) + p.tmem_zero;
  pdl_wait();                                              // q|k|v exist once the projection GEMM has completed

  // one ring stage is filled once more than the others: its first fill is the pos_proj slice of the prologue
  auto fill_index = [](int j) { return j / NS + (j % NS == WP_STAGE ? 1 : 0); };
  auto stage_ptr = [&](int st) { return st < NS_RING ? kp_ring + st * KP_BYTES : rot_s; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: pos_proj slice, then K + E tiles
    if (lane == 0) {
      mbar_expect_tx(q_full, Q_BYTES);                                                 // rows past T arrive as zeros
      tma_load_3d(q_s, &tmQKV, q_full, h * DH, i0, b);
      tma_load_3d(q_s + KV_BYTES, &tmQKV, q_full, h * DH, i0 + BKV, b);
      mbar_expect_tx(&kp_full[WP_STAGE], D * 128);
      tma_load_2d(kp_ring + WP_STAGE * KP_BYTES + KV_BYTES, &tmWp, &kp_full[WP_STAGE], 0, h * D);   // [256 rows (m') x 64 (c)]
      // rotation factors of the 128 query rows (k-block kb = table columns 64*kb .. +63)
      mbar_expect_tx(rot_full, ROT_BYTES);
#pragma unroll
      for (int kb = 0; kb < D / 64; ++kb) tma_load_2d(rot_s + kb * Q_BYTES, &tmRot, rot_full, kb * 64, i0);
    }
    __syncwarp();
    for (int j = 0; j < ntiles; ++j) {
      const int s = j % NS, n = fill_index(j);
      if (j == NS - 1) mbar_wait(qp_full, 0);              // the rotation factors have been consumed: their buffer is a stage now
      mbar_wait(&kp_empty[s], (n & 1) ^ 1);
      if (lane == 0) {
        uint8_t* st = stage_ptr(s);
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
    // ---- prologue: ab = Qs Wp_h   [128 x 256], K = 64 (both operands in shared memory: N = 256 amortises the A reads)
    mbar_wait(q_full, 0);
    mbar_wait(&kp_full[WP_STAGE], 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t sa = umma_desc_lo(smem_u32(q_s)), sb = umma_desc_lo(smem_u32(kp_ring + WP_STAGE * KP_BYTES + KV_BYTES));
#pragma unroll
      for (int k = 0; k < 4; ++k) tc_mma_f16_lo(tmem_base + TM_AB, sa + k * 2, sb + k * 2, idesc_ab, k != 0 ? 1u : 0u);
      tc_commit(ab_full);
      tc_commit(&kp_empty[WP_STAGE]);
    }
    __syncwarp();
    auto issue_s = [&](int j) {
      const int s = j % NS, n = fill_index(j);
      mbar_wait(&kp_full[s], n & 1);
      tc_fence_after();
      if (j >= NSB) FTRACE(39 + 4 * (j - NSB));                  // K | E of tile j landed (stamped in the row of tile j-NSB)
      if (elect_one()) {
        const uint32_t sk = umma_desc_lo(smem_u32(stage_ptr(s))), se = sk + (KV_BYTES >> 4);
        const uint32_t ts = tmem_base + TM_S + (j % NSB) * BKV;
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16_ts(ts, tmem_base + TM_Q + k * 8, sk + k * 2, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int kb = 0; kb < D / 64; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16_ts(ts, tmem_base + TM_QP + (kb * 4 + k) * 8, se + kb * (KV_BYTES >> 4) + k * 2, idesc_s, 1u);
        tc_commit(&s_full[j % NSB]);
        tc_commit(&kp_empty[s]);                                    // K + E may be refilled
      }
      __syncwarp();
    };
    mbar_wait(qp_full, 0);
    tc_fence_after();
    // S runs NSB tiles ahead: S_0 .. S_{NSB-1} up front, then per tile  wait P_j -> P_j.V -> S_{j+NSB}.  P_j.V is issued
    // the moment P_j is ready (the softmax warps' o_done handshake depends on it), and S_{j+NSB} goes to the buffer P_j
    // has just vacated: the tensor pipe executes in issue order, so no further handshake is needed.  While the issuing
    // thread is held up by the back-pressure of S_{j+NSB} (~0.45 us), the softmax warps work on S_{j+1}, complete long ago.
    for (int j = 0; j < NSB && j < ntiles; ++j) issue_s(j);
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      FTRACE(36 + 4 * j);
      mbar_wait(p_full, j & 1);                                     // P_j in TMEM (and O rescaled where needed)
      FTRACE(38 + 4 * j);
      mbar_wait(&v_full[s], (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t tp = tmem_base + TM_S + (j % NSB) * BKV;
        const uint32_t sv = umma_desc_lo(smem_u32(v_ring + s * KV_BYTES));
#pragma unroll
        for (int g = 0; g < 2; ++g)                                 // key half g: keys 32g .. 32g+31 -> its own O
#pragma unroll
          for (int k = 0; k < 2; ++k)                               // K = 32 keys in steps of 16 (V: 16 rows = 2048 B)
            tc_mma_f16_ts(tmem_base + TM_O + 64 * g, tp + 32 * g + k * 8, sv + (2 * g + k) * (2048 >> 4), idesc_pv,
                          (j | k) != 0 ? 1u : 0u);
        tc_commit(o_done);
        tc_commit(&v_empty[s]);
        if (j == ntiles - 1) tc_commit(o_final);
      }
      __syncwarp();
      FTRACE(37 + 4 * j);
      if (j + NSB < ntiles) issue_s(j + NSB);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax: thread = (query row, key half)
    const int w = warp & 3, grp = (warp - 4) >> 2;       // TMEM lane quadrant (hardware rule: warp % 4), half
    const int row = w * 32 + lane;
    const int qi = i0 + row;
    const int sw = row & 7;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(w * 32) << 16);
    if (warp == 4) {
      FTRACE(0);
      if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == TRACE_Z && lane == 0) p.trace[64 + 6] = clock64();
    }
    // ---- Qs: shared memory (as delivered by TMA; the prologue MMA reads it there) -> TMEM, half a row per thread
    {
      mbar_wait(q_full, 0);
      uint32_t qr[16];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint4 v4 = *reinterpret_cast<const uint4*>(q_s + row * 128 + (((grp * 4 + c4) ^ sw) << 4));
        qr[4 * c4 + 0] = v4.x; qr[4 * c4 + 1] = v4.y; qr[4 * c4 + 2] = v4.z; qr[4 * c4 + 3] = v4.w;
      }
      tmem_st_32x16(tlane + TM_Q + 16 * grp, qr);
    }
    // ---- rotate ab into Q' = [A | B] (TMEM); this thread: frequencies 64*grp .. +63 of its row
    {
      if (warp == 4) FTRACE(1);
      mbar_wait(rot_full, 0);
      mbar_wait(ab_full, 0);
      tc_fence_after();
      if (warp == 4) FTRACE(2);
      const uint8_t* crow = rot_s + grp * Q_BYTES + row * 128;          // cos(i w_k), k = 64*grp .. +63
      const uint8_t* srow = rot_s + (2 + grp) * Q_BYTES + row * 128;    // sin(i w_k)
      const float* pc = p.pos_const + h * D;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {                   // 16 frequencies at a time
        const int k0 = grp * 64 + cc * 16;
        uint32_t ar[16], br[16];
        tmem_ld_32x16(tlane + TM_AB + k0, ar);
        tmem_ld_32x16(tlane + TM_AB + NF + k0, br);
        const uint4 c0 = *reinterpret_cast<const uint4*>(crow + (((2 * cc) ^ sw) << 4));
        const uint4 c1 = *reinterpret_cast<const uint4*>(crow + (((2 * cc + 1) ^ sw) << 4));
        const uint4 s0 = *reinterpret_cast<const uint4*>(srow + (((2 * cc) ^ sw) << 4));
        const uint4 s1 = *reinterpret_cast<const uint4*>(srow + (((2 * cc + 1) ^ sw) << 4));
        const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const uint32_t sw8[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        float ka[16], kb2[16];                           // positional-bias constants (warp-uniform addresses)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 va = __ldg(reinterpret_cast<const float4*>(pc + k0) + e);
          const float4 vb = __ldg(reinterpret_cast<const float4*>(pc + NF + k0) + e);
          ka[4 * e] = va.x; ka[4 * e + 1] = va.y; ka[4 * e + 2] = va.z; ka[4 * e + 3] = va.w;
          kb2[4 * e] = vb.x; kb2[4 * e + 1] = vb.y; kb2[4 * e + 2] = vb.z; kb2[4 * e + 3] = vb.w;
        }
        tmem_ld_wait();
        uint32_t oa[8], ob[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float2 cs = __half22float2(*reinterpret_cast<const __half2*>(&cw[e]));
          const float2 sn = __half22float2(*reinterpret_cast<const __half2*>(&sw8[e]));
          const float a0 = __uint_as_float(ar[2 * e]) + ka[2 * e], a1 = __uint_as_float(ar[2 * e + 1]) + ka[2 * e + 1];
          const float b0 = __uint_as_float(br[2 * e]) + kb2[2 * e], b1 = __uint_as_float(br[2 * e + 1]) + kb2[2 * e + 1];
          oa[e] = pack_bf16x2(a0 * sn.x + b0 * cs.x, a1 * sn.y + b1 * cs.y);
          ob[e] = pack_bf16x2(b0 * sn.x - a0 * cs.x, b1 * sn.y - a1 * cs.y);
        }
        tmem_st_32x8(tlane + TM_QP + (k0 >> 1), oa);
        tmem_st_32x8(tlane + TM_QP + ((NF + k0) >> 1), ob);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qp_full);
      if (warp == 4) FTRACE(3);
    }
    // Each key half of a row runs its own online softmax (reference maximum, row sum, O accumulator): the two threads
    // of a row never talk inside the key loop; the halves are merged once, in the epilogue.
    const uint32_t t_o = tlane + TM_O + 64 * grp;        // this thread's O accumulator (64 columns)
    float m_ref = -CUDART_INF_F, l_part = 0.f;

    // usable-key bitmask of this thread's 32 keys of tile j (warp-uniform)
    auto key_mask = [&](int j) -> uint32_t {
      const int jk = j * BKV + 32 * grp;
      if (jk + 32 <= Tn && p.key_pad == nullptr) return 0xffffffffu;
      const int ja = jk + lane;
      const bool va = ja < Tn && (p.key_pad == nullptr || p.key_pad[static_cast<int64_t>(b) * Tn + ja] == 0);
      return __ballot_sync(0xffffffffu, va);
    };
    // rescale this key half's O accumulator and row sum when its reference maximum moves (P.V of tile jdone retired)
    auto move_reference = [&](float m_to, int jdone) {
      mbar_wait(o_done, jdone & 1);
      tc_fence_after();
      const float f = (m_to == m_ref) ? 1.f : ex2f(m_ref - m_to);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t o[32];
        tmem_ld_32x32(t_o + 32 * hh, o);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
        tmem_st_32x32(t_o + 32 * hh, o);
      }
      tmem_st_wait();
      l_part *= f;
      m_ref = m_to;
    };
    // Exact tile: all 32 scores first, their maximum, the lazy-rescale rule, then P.  Used for the first tile and
    // whenever a streamed tile meets a score far above the reference.
    auto exact_tile = [&](int j, uint32_t vm, uint32_t (&pk)[16], float& psum) {
      uint32_t sc[32];
      tmem_ld_32x32(tlane + TM_S + (j % NSB) * BKV + 32 * grp, sc);
      tmem_ld_wait();
      if (vm != 0xffffffffu) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj)
          if (((vm >> jj) & 1u) == 0) sc[jj] = __float_as_uint(-CUDART_INF_F);
      }
      float m0 = max3f(__uint_as_float(sc[0]), __uint_as_float(sc[1]), __uint_as_float(sc[2]));
      float m1 = max3f(__uint_as_float(sc[3]), __uint_as_float(sc[4]), __uint_as_float(sc[5]));
#pragma unroll
      for (int jj = 6; jj < 30; jj += 4) {
        m0 = max3f(m0, __uint_as_float(sc[jj]), __uint_as_float(sc[jj + 1]));
        m1 = max3f(m1, __uint_as_float(sc[jj + 2]), __uint_as_float(sc[jj + 3]));
      }
      const float tmax = max3f(m0, m1, fmaxf(__uint_as_float(sc[30]), __uint_as_float(sc[31])));
      const float m_to = (tmax > m_ref + RESCALE_THRESHOLD) ? tmax : m_ref;
      if (j > 0) {
        if (__any_sync(0xffffffffu, m_to != m_ref)) move_reference(m_to, j - 1);
      }
      m_ref = m_to;
      const float m_use = (m_ref == -CUDART_INF_F) ? 0.f : m_ref;
      float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float p0 = ex2f(__uint_as_float(sc[2 * e]) - m_use);
        const float p1 = ex2f(__uint_as_float(sc[2 * e + 1]) - m_use);
        ps0 += p0;
        ps1 += p1;
        pk[e] = pack_bf16x2(p0, p1);
      }
      psum = ps0 + ps1;
    };

    // Key loop.  TMEM reads (64 B/clk) and the exponentials (MUFU) each take about 512 clocks per tile, so from the second
    // tile on the scores are STREAMED in two chunks of 16 columns against the reference maximum the half already has: the
    // next chunk's tcgen05.ld is in flight while the current one goes through ex2.  The tile maximum falls out of the
    // same pass and moves the reference (lazily, by the 2^8 rule) before the NEXT tile.  A score more than 2^60 above
    // the reference (P would overflow; also the case of a half that has seen no usable key yet) redoes the tile exactly.
    constexpr float STREAM_GUARD = 60.0f;
    float m_pend = m_ref;
    if (warp == 4) FTRACE(8);
    for (int j = 0; j < ntiles; ++j) {
      const int buf = j % NSB;
      const uint32_t vm = key_mask(j);
      const uint32_t t_s = tlane + TM_S + buf * BKV + 32 * grp;
      uint32_t pk[16];
      float psum;
      mbar_wait(&s_full[buf], (j / NSB) & 1);
      tc_fence_after();
      // (S_j complete implies P.V of tile j-2 complete - the tensor pipe runs in issue order - so the parity wait on
      // o_done for tile j-1 inside move_reference cannot alias an older phase, however far this warp has run ahead)
      if (j > 0) {
        if (__any_sync(0xffffffffu, m_pend != m_ref)) move_reference(m_pend, j - 1);
      }
      if (warp == 4) FTRACE(9 + 4 * j);
      if (j == 0) {
        exact_tile(0, vm, pk, psum);
        m_pend = m_ref;
      } else {
        const float m_use = (m_ref == -CUDART_INF_F) ? 0.f : m_ref;
        float ps0 = 0.f, ps1 = 0.f, t0 = -CUDART_INF_F, t1 = -CUDART_INF_F;
        uint32_t ca[16], cb[16];
        auto chunk = [&](const uint32_t (&cur)[16], int c) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float s0 = __uint_as_float(cur[2 * e]), s1 = __uint_as_float(cur[2 * e + 1]);
            if (vm != 0xffffffffu) {
              if (((vm >> (16 * c + 2 * e)) & 1u) == 0) s0 = -CUDART_INF_F;
              if (((vm >> (16 * c + 2 * e + 1)) & 1u) == 0) s1 = -CUDART_INF_F;
            }
            t0 = fmaxf(t0, s0);
            t1 = fmaxf(t1, s1);
            const float p0 = ex2f(s0 - m_use), p1 = ex2f(s1 - m_use);
            ps0 += p0;
            ps1 += p1;
            pk[8 * c + e] = pack_bf16x2(p0, p1);
          }
        };
        tmem_ld_32x16(t_s, ca);
        tmem_ld_wait_for16(ca);
        tmem_ld_32x16(t_s + 16, cb);
        chunk(ca, 0);
        tmem_ld_wait_for16(cb);
        chunk(cb, 1);
        const float tmax = fmaxf(t0, t1);
        psum = ps0 + ps1;
        if (__any_sync(0xffffffffu, tmax > m_ref + STREAM_GUARD)) {
          exact_tile(j, vm, pk, psum);
          m_pend = m_ref;
        } else {
          m_pend = (tmax > m_ref + RESCALE_THRESHOLD) ? tmax : m_ref;
        }
      }
      if (warp == 4) FTRACE(10 + 4 * j);
      // ---- P_j of this key half (bf16 pairs) overwrites columns of S_j that only this thread reads
      tmem_st_32x16(t_s, pk);
      tmem_st_wait();
      l_part += psum;
      tc_fence_before();
      // Consumer -> producer handshake.  S_{j+1} is issued BEFORE the MMA warp waits for P_j, so without it the softmax
      // warps could complete the p_full phases of tiles j and j+1 while the MMA warp is still between its issue loop and
      // its poll - its parity wait would then alias and never return (seen once per ~10^5 CTAs).  P.V of tile j-1
      // complete implies the MMA warp has consumed phase j-1; it was issued a whole tile ago, so this wait is free, and
      // every warp follows the o_done phases one by one (no aliasing in move_reference either).
      if (j > 0) mbar_wait(o_done, (j - 1) & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == 4) FTRACE(11 + 4 * j);
    }
    // ---- epilogue: merge the two key halves, O = (f0 O0 + f1 O1) / (f0 l0 + f1 l1), f_g = 2^(m_g - max(m0, m1));
    //      this thread writes 32 of the row's 64 output columns
    xmax[grp * BQ + row] = m_ref;
    xsum[grp * BQ + row] = l_part;
    // (o_done cannot be used here: this warp has not followed its phases, and a parity wait issued after BOTH of the last
    // two P.V commits would wait for a phase that never completes)
    mbar_wait(o_final, 0);
    tc_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld_32x32(tlane + TM_O + 32 * grp, o0);
    tmem_ld_32x32(tlane + TM_O + 64 + 32 * grp, o1);
    tmem_ld_wait();
    named_bar_sync(1 + w, 64);
    const float m_oth = xmax[(grp ^ 1) * BQ + row], l_oth = xsum[(grp ^ 1) * BQ + row];
    const float m_all = fmaxf(m_ref, m_oth);
    const float f_me = (m_ref == -CUDART_INF_F) ? 0.f : ex2f(m_ref - m_all);
    const float f_oth = (m_oth == -CUDART_INF_F) ? 0.f : ex2f(m_oth - m_all);
    const float l_tot = f_me * l_part + f_oth * l_oth;
    const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
    const float w0 = (grp == 0 ? f_me : f_oth) * inv, w1 = (grp == 0 ? f_oth : f_me) * inv;
    if (qi < Tn) {
      bf16* orow = p.ctx + (static_cast<int64_t>(b) * Tn + qi) * d + h * DH + 32 * grp;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t r[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = ch * 8 + 2 * e;
          r[e] = pack_bf16x2(__uint_as_float(o0[c]) * w0 + __uint_as_float(o1[c]) * w1,
                             __uint_as_float(o0[c + 1]) * w0 + __uint_as_float(o1[c + 1]) * w1);
        }
        *reinterpret_cast<uint4*>(orow + ch * 8) = make_uint4(r[0], r[1], r[2], r[3]);
      }
    }
    if (warp == 4) {
      FTRACE(4);
      if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == TRACE_Z && lane == 0) p.trace[64 + 7] = clock64();
    }
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

// query rows of the fused q|k|v projection for this kernel: W <- c W, b <- c (b + u)      (in place, rows [0, d))
__global__ void pack_q_fact_kernel(float* __restrict__ w, float* __restrict__ bias, const float* __restrict__ u, int d, float c) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx < static_cast<int64_t>(d) * d) w[idx] *= c;
  if (idx < d) bias[idx] = c * (bias[idx] + u[idx]);
}

// out[h][m'] = c * sum_cc (v - u)[h*dh + cc] * Wp[h*dh + cc][m(m')]      (m(m') as in pack_pos_fact_kernel)
__global__ void pos_const_fact_kernel(const float* __restrict__ wp, const float* __restrict__ u, const float* __restrict__ v,
                                      float* __restrict__ out, int d, int dh, float c) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int H = d / dh;
  if (idx >= H * d) return;
  const int h = idx / d, mp = idx - h * d;
  const int m = mp < d / 2 ? 2 * mp : 2 * (mp - d / 2) + 1;
  float acc = 0.f;
  for (int cc = 0; cc < dh; ++cc) acc += (v[h * dh + cc] - u[h * dh + cc]) * wp[static_cast<int64_t>(h * dh + cc) * d + m];
  out[idx] = c * acc;
}

}  // namespace

extern unsigned long long* g_gemm_trace;

int pack_q_fact(float* wq, float* bq, const float* u, int d, int dh, cudaStream_t stream) {
  const int64_t n = static_cast<int64_t>(d) * d;
  pack_q_fact_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(wq, bq, u, d,
                                                                               1.4426950408889634f / sqrtf(static_cast<float>(dh)));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

int pos_const_fact(const float* wp, const float* u, const float* v, float* out, int d, int dh, cudaStream_t stream) {
  const int n = (d / dh) * d;
  pos_const_fact_kernel<<<(n + 255) / 256, 256, 0, stream>>>(wp, u, v, out, d, dh,
                                                             1.4426950408889634f / sqrtf(static_cast<float>(dh)));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

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

int rel_attention_bf16_fact(const bf16* qkv, const bf16* wpT, const bf16* e16, const void* r16, const float* pos_const,
                            const uint8_t* key_pad, bf16* ctx, int B, int Tn, int H, int dh, cudaStream_t stream) {
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
  static std::atomic<uint64_t> attr_done{0};
  CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(rel_attention_fact_kernel), SMEM_BYTES));
  FactParams p{pos_const, key_pad, ctx, Tn, H, 0u, g_gemm_trace};
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, BQ)), static_cast<unsigned>(H), static_cast<unsigned>(B));
  CFM_CUDA_OK(launch_pdl(rel_attention_fact_kernel, grid, dim3(NUM_THREADS), SMEM_BYTES, stream, 1, tmQKV, tmE, tmWp, tmRot, p));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
