// Relative-position multi-head self-attention on tcgen05 tensor cores (head size 64), flash style.
//
//   score[i,j] = ((q_i+u).k_j + (q_i+v).p[T-1+j-i]) / sqrt(dh)         (attention.py:47-56, shift :94-102)
//
// One CTA = 128 query rows of one (batch, head); keys walked in tiles of 64.  Warp roles:
//   warp 0      TMA producer: K tile [64x64], V tile [64x64] (3-D maps over q|k|v, rows past T zero-filled) and the
//               192-row window of the projected positional table this (query tile, key tile) pair can touch
//               (2-D map, negative / past-the-end rows zero-filled) into a 2-stage ring;
//   warp 1      MMA issuer: S = (q+u) K^T (128x64) and PS = (q+v) Pwin^T (128x192) into TMEM, later O_j = P V
//               (128x64, V is the MN-major B operand).  TMEM holds two {S|O, PS} buffers (2 x 256 columns) so the
//               MMAs of tile j+1 run while the softmax warps work on tile j;
//   warps 4-7   softmax, one thread = one query row (no cross-thread reductions).  The Transformer-XL shift is a
//               per-row column offset 127-i into PS: its warp-uniform part is folded into the tcgen05.ld column
//               address, the per-lane part (31-lane) goes through a private shared-memory row (written as float2,
//               read back at a lane-dependent offset, bank-conflict free) - no [T,2T-1] tensor is ever formed.
//               Online softmax in fp32 with exp2 (scale folded into Q); P is written as bf16 straight into the
//               MMA's 128B-swizzled A layout; O_j is read back from TMEM and merged into a register accumulator
//               (acc = acc*alpha + O_j), so TMEM never needs rescaling.
// Masked keys (and keys past T) get weight exactly 0.
#include <math_constants.h>

#include "ops.h"

namespace cfm {
namespace {

constexpr int BQ = 128, BKV = 64, DH = 64;
constexpr int WIN = BQ + BKV;                         // 192 window rows per (query tile, key tile)
constexpr int Q_BYTES = BQ * 128;                     // 16 KB: [128 x 64] bf16, 128B swizzle
constexpr int KV_BYTES = BKV * 128;                   // 8 KB
constexpr int WIN_BYTES = WIN * 128;                  // 24 KB
constexpr int STAGE_BYTES = 2 * KV_BYTES + WIN_BYTES; // K | V | window = 40 KB
constexpr int STAGES = 2;
constexpr int P_BYTES = BQ * 128;                     // 16 KB: [128 x 64 keys] bf16
constexpr int SKEW_LD = 98;                           // floats per private row: 96 used; 98 % 32 == 2 -> conflict free
constexpr int SKEW_BYTES = 4 * 32 * SKEW_LD * 4;      // 4 softmax warps
constexpr int SMEM_BYTES = 1024 + 2 * Q_BYTES + STAGES * STAGE_BYTES + 2 * P_BYTES + SKEW_BYTES + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "attention shared memory budget");
constexpr int NUM_THREADS = 256;                      // warps 0,1: TMA / MMA; 2: TMEM alloc; 3: idle; 4-7: softmax
constexpr int TM_BUF = 256;                           // TMEM columns per buffer: S|O at +0 (64), PS at +64 (192)

struct AttnParams {
  const bf16* qkv;
  const float* ub;
  const float* vb;
  const uint8_t* key_pad;
  bf16* ctx;
  int Tn, H;
  float scale_log2e;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
rel_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmPos, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* qu_s = smem;
  uint8_t* qv_s = qu_s + Q_BYTES;
  uint8_t* ring = qv_s + Q_BYTES;
  uint8_t* p_s = ring + STAGES * STAGE_BYTES;          // two P buffers
  float* skew = reinterpret_cast<float*>(p_s + 2 * P_BYTES);
  uint64_t* kv_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(skew) + SKEW_BYTES);
  uint64_t* kv_empty = kv_full + 2;
  uint64_t* sp_full = kv_empty + 2;
  uint64_t* spo_empty = sp_full + 2;
  uint64_t* p_full = spo_empty + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* q_full = o_full + 2;
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
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
      mbar_init(&sp_full[s], 1);
      mbar_init(&spo_empty[s], 4);
      mbar_init(&p_full[s], 4);
      mbar_init(&o_full[s], 1);
    }
    mbar_init(q_full, 4);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    for (int j = 0; j < ntiles; ++j) {
      const int s = j & 1;
      mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
      if (lane == 0) {
        uint8_t* st = ring + s * STAGE_BYTES;
        const int j0 = j * BKV;
        mbar_expect_tx(&kv_full[s], STAGE_BYTES);
        tma_load_3d(st, &tmQKV, &kv_full[s], d + h * DH, j0, b);                    // K rows j0..j0+63
        tma_load_3d(st + KV_BYTES, &tmQKV, &kv_full[s], 2 * d + h * DH, j0, b);     // V rows
        // window entry w <-> table row (T-1 + j0 - i0 - 127) + w
        tma_load_2d(st + 2 * KV_BYTES, &tmPos, &kv_full[s], h * DH, Tn - 1 + j0 - i0 - (BQ - 1));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_bf16(BQ, BKV);          // 128 x 64
    constexpr uint32_t idesc_ps = umma_idesc_bf16(BQ, WIN);         // 128 x 192
    constexpr uint32_t idesc_pv = umma_idesc_bf16_bmn(BQ, DH);      // 128 x 64, B (=V) MN-major
    auto issue_sp = [&](int j) {
      const int buf = j & 1, s = j & 1;
      mbar_wait(&spo_empty[buf], ((j >> 1) & 1) ^ 1);
      mbar_wait(&kv_full[s], (j >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sk = smem_u32(ring + s * STAGE_BYTES), sw = sk + 2 * KV_BYTES;
        const uint32_t squ = smem_u32(qu_s), sqv = smem_u32(qv_s);
        const uint32_t ts = tmem_base + buf * TM_BUF, tps = ts + BKV;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(ts, umma_desc_k_sw128(squ + k * 32), umma_desc_k_sw128(sk + k * 32), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc_mma_f16(tps, umma_desc_k_sw128(sqv + k * 32), umma_desc_k_sw128(sw + k * 32), idesc_ps, k != 0 ? 1u : 0u);
        tc_commit(&sp_full[buf]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    tc_fence_after();
    issue_sp(0);
    for (int j = 0; j < ntiles; ++j) {
      if (j + 1 < ntiles) issue_sp(j + 1);
      const int buf = j & 1, s = j & 1;
      mbar_wait(&p_full[buf], (j >> 1) & 1);           // P_j written (and S_j fully read: its columns become O_j)
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sp = smem_u32(p_s + buf * P_BYTES);
        const uint32_t sv = smem_u32(ring + s * STAGE_BYTES + KV_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)                      // K = 64 keys in steps of 16 (V: 16 rows = 2048 B)
          tc_mma_f16(tmem_base + buf * TM_BUF, umma_desc_k_sw128(sp + k * 32), umma_desc_mn_sw128(sv + k * 2048), idesc_pv,
                     k != 0 ? 1u : 0u);
        tc_commit(&o_full[buf]);
        tc_commit(&kv_empty[s]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax warps: thread = query row
    const int w = warp & 3;                              // TMEM lane quadrant
    const int row = w * 32 + lane;
    const int qi = i0 + row;
    const int sw = row & 7;
    // ---- Q operands: (q+u)*c and (q+v)*c as bf16 in the MMA's A layout, c = log2(e)/sqrt(dh)
    {
      const bf16* qrow = p.qkv + (static_cast<int64_t>(b) * Tn + (qi < Tn ? qi : 0)) * (3 * d) + h * DH;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
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
    }
    float* myrow = skew + (w * 32 + lane) * SKEW_LD;
    const int shift = 31 - lane;                         // per-lane part of the relative shift
    const uint32_t lane_addr = static_cast<uint32_t>(w * 32) << 16;
    float acc[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.f;
    float m_run = -CUDART_INF_F, l_run = 0.f, alpha_prev = 1.f;

    for (int j = 0; j <= ntiles; ++j) {
      // ---- merge O_{j-1} (its P.V finished) and hand its TMEM buffer back
      if (j > 0) {
        const int pb = (j - 1) & 1;
        mbar_wait(&o_full[pb], ((j - 1) >> 1) & 1);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(tmem_base + lane_addr + pb * TM_BUF, o0);
        tmem_ld_32x32(tmem_base + lane_addr + pb * TM_BUF + 32, o1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&spo_empty[pb]);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          acc[c] = fmaf(acc[c], alpha_prev, __uint_as_float(o0[c]));
          acc[32 + c] = fmaf(acc[32 + c], alpha_prev, __uint_as_float(o1[c]));
        }
      }
      if (j == ntiles) break;
      const int buf = j & 1;
      const int j0 = j * BKV;
      // ---- usable-key bitmask of this tile (warp-uniform)
      uint32_t vm0 = 0xffffffffu, vm1 = 0xffffffffu;
      bool need_mask = false;
      if (j0 + BKV > Tn || p.key_pad != nullptr) {
        const int ja = j0 + lane, jb = ja + 32;
        const bool va = ja < Tn && (p.key_pad == nullptr || p.key_pad[static_cast<int64_t>(b) * Tn + ja] == 0);
        const bool vb = jb < Tn && (p.key_pad == nullptr || p.key_pad[static_cast<int64_t>(b) * Tn + jb] == 0);
        vm0 = __ballot_sync(0xffffffffu, va);
        vm1 = __ballot_sync(0xffffffffu, vb);
        need_mask = (vm0 & vm1) != 0xffffffffu;
      }
      mbar_wait(&sp_full[buf], (j >> 1) & 1);
      tc_fence_after();
      // ---- positional scores: PS columns [32*(3-w), +96) -> private smem row (float2 stores)
      const uint32_t tps = tmem_base + lane_addr + buf * TM_BUF + BKV + 32 * (3 - w);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(tps + ch * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e)
          *reinterpret_cast<float2*>(myrow + ch * 32 + 2 * e) =
              make_float2(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
      }
      // ---- content scores + shifted positional scores, mask, tile max
      float sc[BKV];
      {
        uint32_t s0[32], s1[32];
        tmem_ld_32x32(tmem_base + lane_addr + buf * TM_BUF, s0);
        tmem_ld_32x32(tmem_base + lane_addr + buf * TM_BUF + 32, s1);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          sc[jj] = __uint_as_float(s0[jj]) + myrow[shift + jj];
          sc[32 + jj] = __uint_as_float(s1[jj]) + myrow[shift + 32 + jj];
        }
      }
      float tmax = -CUDART_INF_F;
      if (need_mask) {
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          if (((vm0 >> jj) & 1u) == 0) sc[jj] = -CUDART_INF_F;
          if (((vm1 >> jj) & 1u) == 0) sc[32 + jj] = -CUDART_INF_F;
        }
      }
#pragma unroll
      for (int jj = 0; jj < BKV; ++jj) tmax = fmaxf(tmax, sc[jj]);
      const float m_new = fmaxf(m_run, tmax);
      const float m_use = (m_new == -CUDART_INF_F) ? 0.f : m_new;
      const float alpha = ex2f(m_run - m_use);
      m_run = m_new;
      // ---- P = exp2(s - m) as bf16 into the A layout of the P.V MMA
      float psum = 0.f;
      uint8_t* prow = p_s + buf * P_BYTES + row * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = ex2f(sc[ch * 8 + 2 * e] - m_use), p1 = ex2f(sc[ch * 8 + 2 * e + 1] - m_use);
          psum += p0 + p1;
          o[e] = pack_bf16x2(p0, p1);
        }
        *reinterpret_cast<uint4*>(prow + ((ch ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      l_run = l_run * alpha + psum;
      alpha_prev = alpha;
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[buf]);
    }
    // ---- normalise and store this row of the context
    if (qi < Tn) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* orow = p.ctx + (static_cast<int64_t>(b) * Tn + qi) * d + h * DH;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 o;
        o.x = pack_bf16x2(acc[ch * 8 + 0] * inv, acc[ch * 8 + 1] * inv);
        o.y = pack_bf16x2(acc[ch * 8 + 2] * inv, acc[ch * 8 + 3] * inv);
        o.z = pack_bf16x2(acc[ch * 8 + 4] * inv, acc[ch * 8 + 5] * inv);
        o.w = pack_bf16x2(acc[ch * 8 + 6] * inv, acc[ch * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(orow + ch * 8) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int rel_attention_bf16_tc(const bf16* qkv, const bf16* pos, int ldp, const float* u, const float* v,
                          const uint8_t* key_pad, bf16* ctx, int B, int Tn, int H, int dh, cudaStream_t stream) {
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
  static bool attr_done = false;
  if (!attr_done) {
    CFM_CUDA_OK(cudaFuncSetAttribute(rel_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  AttnParams p{qkv, u, v, key_pad, ctx, Tn, H, 1.4426950408889634f / sqrtf(static_cast<float>(dh))};
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, BQ)), static_cast<unsigned>(H), static_cast<unsigned>(B));
  rel_attention_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQKV, tmPos, p);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
