// Fused ConvolutionModule for d_model = 256, kernel size 31 (convolution.py:21-32 + the residual of block.py:23):
//
//     out[b,t,:] = res[b,t,:] + W2 . swish( bn( dw31( glu( W1 . LN(x[b,t,:]) + b1 ) ) ) ) + b2
//
// One CTA = one tile of TO = 96 output frames of one utterance.  Nothing but x, the weights and `out` touches
// global memory; the [T,2d] pointwise output, the GLU output and the depthwise output never leave the SM:
//   prologue   8 compute warps LayerNorm the 126 frames the tile needs (96 + a 15-frame halo each side; frames
//              outside [0,T) are the Conv1d zero padding) and write the bf16 A operand in the MMA's swizzled layout;
//   pw1        tcgen05: [128 x 256] . W1p^T as two N = 256 halves into TMEM (2 x 256 columns); W1p rows are packed
//              per 64 as 32 value rows + their 32 gate rows, so a thread's GLU pair sits 32 columns apart;
//   GLU        thread = frame: value * sigmoid(gate) (+bias), zeroed outside [0,T), bf16 -> shared [126][256]
//              (row pitch 528 B: conflict-free for both the row-wise writes and the channel-wise reads);
//   depthwise  thread = channel pair x 48 frames: the 31 BN-folded taps live in registers, the frames slide through
//              an 8-output register window, MACs are packed fma.rn.f32x2 (measured: the fp32 FMA pipe, ~64 lanes per
//              clock per SM with three register operands, bounds this phase - scalar and packed FMAs time the same);
//              Swish; bf16 result is written straight into the A-operand layout of pw2 (re-using the LN(x) buffer);
//   pw2        tcgen05: [128 x 256] . W2^T, accumulating over the first TMEM half;
//   epilogue   accumulators staged as fp32 in shared memory, then one warp per frame adds b2 and the residual and
//              writes `out` with fully coalesced 16-byte accesses.
// W1p/W2 stream through a 3-stage TMA/mbarrier ring ([256 rows x 64 k] bf16 per stage).
#include "ops.h"

namespace cfm {
namespace {

constexpr int D = 256, KS = 31, HALF = 15, BM = 128, TO = 96, ROWS = TO + KS - 1;   // 126 frames per tile
constexpr int KB = 64;
constexpr int XS_BYTES = (D / KB) * BM * 128;              // 64 KB: A operand (LN(x), later the depthwise output)
constexpr int GS_LD = 2 * D + 16;                          // bytes per GLU row (bf16) + 16 B pad
constexpr int GS_BYTES = ROWS * GS_LD;                     // 66528
constexpr int GS_BYTES_AL = (GS_BYTES + 1023) & ~1023;
constexpr int STAGE_BYTES = 32 * 1024, STAGES = 3;
constexpr int NUM_CW = 8, NUM_CT = NUM_CW * 32, NUM_THREADS = 128 + NUM_CT;
constexpr int SMEM_BYTES = 1024 + XS_BYTES + GS_BYTES_AL + STAGES * STAGE_BYTES + 128;
static_assert(SMEM_BYTES <= 227 * 1024, "fused conv module shared memory budget");
constexpr int OUT_LD = D + 4;                              // fp32 staging row pitch (floats): conflict-free row-wise float4 stores
static_assert(TO * OUT_LD * 4 <= XS_BYTES + GS_BYTES_AL, "output staging must fit the dead operand buffers");
#ifndef CFM_DW_PACKED
#define CFM_DW_PACKED 1
#endif
constexpr int DW_RT = 8;                                   // outputs per register window
constexpr int DW_PER_THREAD = TO / 2;                      // 48 frames per thread (two time groups)

struct ConvParams {
  const float* x;
  const float* res;      // nullable
  const float* ln_g;
  const float* ln_b;
  const float* b1;       // packed like W1p
  const float* dw_w;     // [31][256] BN-folded
  const float* dw_b;     // [256]     BN-folded
  const float* b2;
  float* out;
  const float* ln2_g;    // nullable: also write a_out = LN(out; ln2_g, ln2_b) as bf16, the normalised input of the
  const float* ln2_b;    //           half-step FFN that follows the conv module in the block (ffn.py:16)
  bf16* a_out;
  int Tn;
  unsigned long long* trace;     // debug: globaltimer stamps of CTA (0,0)
};
__device__ __forceinline__ unsigned long long ctime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CTRACE(slot)                                                                                  \
  do {                                                                                                \
    if (p.trace != nullptr && blockIdx.x == 1 && blockIdx.y == 0 && threadIdx.x == 128) p.trace[(slot)] = ctime_ns(); \
  } while (0)

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ra, rb, rc, rd;
  ra = (static_cast<uint64_t>(__float_as_uint(a.y)) << 32) | __float_as_uint(a.x);
  rb = (static_cast<uint64_t>(__float_as_uint(b.y)) << 32) | __float_as_uint(b.x);
  rc = (static_cast<uint64_t>(__float_as_uint(c.y)) << 32) | __float_as_uint(c.x);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return make_float2(__uint_as_float(static_cast<uint32_t>(rd)), __uint_as_float(static_cast<uint32_t>(rd >> 32)));
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_module_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* xs = smem;
  uint8_t* gs = xs + XS_BYTES;
  uint8_t* ring = gs + GS_BYTES_AL;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* xs_full = empty_bar + STAGES;
  uint64_t* acc1_full = xs_full + 1;      // [2]
  uint64_t* hs_full = acc1_full + 2;
  uint64_t* acc2_full = hs_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, t_out0 = blockIdx.x * TO, t_in0 = t_out0 - HALF;
  const int Tn = p.Tn;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(xs_full, NUM_CW);
    mbar_init(&acc1_full[0], 1);
    mbar_init(&acc1_full[1], 1);
    mbar_init(hs_full, NUM_CW);
    mbar_init(acc2_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_whole_sm(tmem_slot);
  pdl_wait();                                              // x exists once the previous kernel has completed

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer: 8 W1p stages, 4 W2 stages
    for (uint32_t it = 0; it < 12; ++it) {
      const int s = it % STAGES;
      mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
      if (lane == 0) {
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        if (it < 8) tma_load_2d(ring + s * STAGE_BYTES, &tmW1, &full_bar[s], (it & 3) * KB, (it >> 2) * 256);
        else tma_load_2d(ring + s * STAGE_BYTES, &tmW2, &full_bar[s], (it & 3) * KB, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BM, 256);
    mbar_wait(xs_full, 0);
    tc_fence_after();
    for (uint32_t it = 0; it < 12; ++it) {
      const int s = it % STAGES, kb = it & 3, g = it >> 2;        // g: 0,1 = pw1 halves, 2 = pw2
      if (it == 8) {
        mbar_wait(hs_full, 0);                                    // depthwise output in place, both TMEM halves drained
        tc_fence_after();
      }
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = umma_desc_lo(smem_u32(xs + kb * (BM * 128))), sb = umma_desc_lo(smem_u32(ring + s * STAGE_BYTES));
        const uint32_t dcol = tmem_base + (g == 1 ? 256u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_f16_lo(dcol, sa + k * 2, sb + k * 2, idesc, (kb | k) != 0 ? 1u : 0u);
        tc_commit(&empty_bar[s]);
        if (kb == 3) tc_commit(g == 0 ? &acc1_full[0] : (g == 1 ? &acc1_full[1] : acc2_full));
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ compute warps
    const int cw = warp - 4, q = warp & 3, grp = cw >> 2;
    const int tid_c = threadIdx.x - 128;
    CTRACE(0);
    // ---- LayerNorm: warp cw normalises tile rows cw*16 .. +15; lane owns columns 8*lane .. +7
    {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_g) + 2 * lane);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.ln_g) + 2 * lane + 1);
      const float4 be0 = __ldg(reinterpret_cast<const float4*>(p.ln_b) + 2 * lane);
      const float4 be1 = __ldg(reinterpret_cast<const float4*>(p.ln_b) + 2 * lane + 1);
      float4 a[16], c[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = cw * 16 + i, t = t_in0 + r;
        if (r < ROWS && t >= 0 && t < Tn) {
          const float4* xr = reinterpret_cast<const float4*>(p.x + (static_cast<int64_t>(b) * Tn + t) * D);
          a[i] = xr[2 * lane];
          c[i] = xr[2 * lane + 1];
        } else {
          a[i] = c[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = cw * 16 + i;
        float s = (a[i].x + a[i].y) + (a[i].z + a[i].w) + (c[i].x + c[i].y) + (c[i].z + c[i].w);
        const float mean = warp_sum(s) * (1.0f / D);
        a[i].x -= mean; a[i].y -= mean; a[i].z -= mean; a[i].w -= mean;
        c[i].x -= mean; c[i].y -= mean; c[i].z -= mean; c[i].w -= mean;
        float v = (a[i].x * a[i].x + a[i].y * a[i].y) + (a[i].z * a[i].z + a[i].w * a[i].w) +
                  (c[i].x * c[i].x + c[i].y * c[i].y) + (c[i].z * c[i].z + c[i].w * c[i].w);
        const float rstd = rsqrtf(warp_sum(v) * (1.0f / D) + 1e-5f);
        uint4 o;
        o.x = pack_bf16x2(a[i].x * rstd * g0.x + be0.x, a[i].y * rstd * g0.y + be0.y);
        o.y = pack_bf16x2(a[i].z * rstd * g0.z + be0.z, a[i].w * rstd * g0.w + be0.w);
        o.z = pack_bf16x2(c[i].x * rstd * g1.x + be1.x, c[i].y * rstd * g1.y + be1.y);
        o.w = pack_bf16x2(c[i].z * rstd * g1.z + be1.z, c[i].w * rstd * g1.w + be1.w);
        *reinterpret_cast<uint4*>(xs + (lane >> 3) * (BM * 128) + r * 128 + (((lane & 7) ^ (r & 7)) << 4)) = o;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(xs_full);
    }
    CTRACE(1);
    // ---- depthwise taps of this thread's channel pair (latency hidden behind pw1)
    const int cp = tid_c & 127, tg = tid_c >> 7;
    float2 wt[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) wt[j] = __ldg(reinterpret_cast<const float2*>(p.dw_w + j * D) + cp);
    const float2 dwb = __ldg(reinterpret_cast<const float2*>(p.dw_b) + cp);

    // ---- GLU epilogue of pw1: thread = frame (tile row), 2 x 32 channels per half
    const int row = q * 32 + lane;
    const bool row_ok = row < ROWS && (t_in0 + row) >= 0 && (t_in0 + row) < Tn;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      mbar_wait(&acc1_full[h], 0);
      tc_fence_after();
      CTRACE(2 + h);
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        const int cg = grp * 2 + i;                                // 32-channel group within the half
        uint32_t va[32], ga[32];
        tmem_ld_32x32(tmem_base + lane_addr + h * 256 + cg * 64, va);
        tmem_ld_32x32(tmem_base + lane_addr + h * 256 + cg * 64 + 32, ga);
        tmem_ld_wait();
        if (row < ROWS) {
          const float4* bv = reinterpret_cast<const float4*>(p.b1 + h * 256 + cg * 64);
          uint8_t* dst = gs + row * GS_LD + (h * 128 + cg * 32) * 2;
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float4 b4 = __ldg(bv + c8 * 2 + e), g4 = __ldg(bv + 8 + c8 * 2 + e);
              const int k = c8 * 8 + e * 4;
              o[e * 4 + 0] = (__uint_as_float(va[k + 0]) + b4.x) * sigmoid_tanh(__uint_as_float(ga[k + 0]) + g4.x);
              o[e * 4 + 1] = (__uint_as_float(va[k + 1]) + b4.y) * sigmoid_tanh(__uint_as_float(ga[k + 1]) + g4.y);
              o[e * 4 + 2] = (__uint_as_float(va[k + 2]) + b4.z) * sigmoid_tanh(__uint_as_float(ga[k + 2]) + g4.z);
              o[e * 4 + 3] = (__uint_as_float(va[k + 3]) + b4.w) * sigmoid_tanh(__uint_as_float(ga[k + 3]) + g4.w);
            }
            uint4 v4 = make_uint4(0u, 0u, 0u, 0u);                 // Conv1d zero padding outside [0,T)
            if (row_ok) {
              v4.x = pack_bf16x2(o[0], o[1]); v4.y = pack_bf16x2(o[2], o[3]);
              v4.z = pack_bf16x2(o[4], o[5]); v4.w = pack_bf16x2(o[6], o[7]);
            }
            *reinterpret_cast<uint4*>(dst + c8 * 16) = v4;
          }
        }
      }
    }
    tc_fence_before();
    CTRACE(4);
    named_bar_sync(1, NUM_CT);                                     // GLU tile complete; LN(x) operand is dead (pw1 done)

    // ---- depthwise conv (k = 31) + folded BatchNorm + Swish -> A operand of pw2
    CTRACE(5);
    {
      const uint8_t* gcol = gs + cp * 4;
      uint8_t* hcol = xs + (cp >> 5) * (BM * 128) + (cp & 3) * 4;
      const int chunk = (cp & 31) >> 2;
#pragma unroll 1
      for (int blk = 0; blk < DW_PER_THREAD / DW_RT; ++blk) {
        const int o0 = tg * DW_PER_THREAD + blk * DW_RT;
        float2 win[DW_RT + KS - 1];
#pragma unroll
        for (int i = 0; i < DW_RT + KS - 1; ++i) {
          const uint32_t raw = *reinterpret_cast<const uint32_t*>(gcol + (o0 + i) * GS_LD);
          win[i] = make_float2(__uint_as_float(raw << 16), __uint_as_float(raw & 0xffff0000u));
        }
        float2 acc[DW_RT];
#pragma unroll
        for (int i = 0; i < DW_RT; ++i) acc[i] = dwb;
#pragma unroll
        for (int j = 0; j < KS; ++j)
#pragma unroll
          for (int i = 0; i < DW_RT; ++i) {
#if CFM_DW_PACKED
            acc[i] = ffma2(wt[j], win[i + j], acc[i]);
#else
            acc[i].x = fmaf(wt[j].x, win[i + j].x, acc[i].x);
            acc[i].y = fmaf(wt[j].y, win[i + j].y, acc[i].y);
#endif
          }
#pragma unroll
        for (int i = 0; i < DW_RT; ++i) {
          const int o = o0 + i;
          *reinterpret_cast<uint32_t*>(hcol + o * 128 + ((chunk ^ (o & 7)) << 4)) =
              pack_bf16x2(swish_tanh(acc[i].x), swish_tanh(acc[i].y));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(hs_full);
    }
    CTRACE(6);

    // ---- final epilogue: out = res + acc2 + b2, staged fp32 (frames past T are skipped)
    // The residual rows of this warp's frames (frame o = cw + 8 i; lane = columns 4 lane .. +3 and 128 + 4 lane .. +3, so
    // a warp reads 512 contiguous bytes per instruction from shared and global memory alike) are requested BEFORE the
    // wait on pw2: their latency hides behind the MMAs and the staging pass.
    constexpr int EP_ROWS = TO / NUM_CW, EP_EARLY = 8;             // the last rows are requested after the staging pass
    float4 r0[EP_ROWS], r1[EP_ROWS];                               // (its 32 TMEM registers and these do not fit together)
    auto load_res = [&](int i) {
      const int t = t_out0 + cw + i * NUM_CW;
      r0[i] = r1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.res != nullptr && t < Tn) {
        const float4* rr = reinterpret_cast<const float4*>(p.res + (static_cast<int64_t>(b) * Tn + t) * D);
        r0[i] = rr[lane];
        r1[i] = rr[lane + 32];
      }
    };
#pragma unroll
    for (int i = 0; i < EP_EARLY; ++i) load_res(i);
    mbar_wait(acc2_full, 0);
    tc_fence_after();
    CTRACE(7);
    uint8_t* stage_out = smem;                                     // aliases xs/gs: every MMA and the depthwise pass are done
    if (q < 3) {                                                   // tile rows 96..127 carry no output
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        const int col = grp * 128 + ch * 32;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + lane_addr + col, r);
        tmem_ld_wait();
        float* srow = reinterpret_cast<float*>(stage_out) + row * OUT_LD + col;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(srow + j * 4) = make_float4(__uint_as_float(r[j * 4 + 0]), __uint_as_float(r[j * 4 + 1]),
                                                                 __uint_as_float(r[j * 4 + 2]), __uint_as_float(r[j * 4 + 3]));
      }
    }
    tc_fence_before();
    CTRACE(8);
#pragma unroll
    for (int i = EP_EARLY; i < EP_ROWS; ++i) load_res(i);
    named_bar_sync(1, NUM_CT);
    // second pass, coalesced: one warp = one frame
    {
      const float4 b2a = __ldg(reinterpret_cast<const float4*>(p.b2) + lane);
      const float4 b2b = __ldg(reinterpret_cast<const float4*>(p.b2) + lane + 32);
      float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga, ba = ga, bb = ga;
      if (p.a_out != nullptr) {
        ga = __ldg(reinterpret_cast<const float4*>(p.ln2_g) + lane); gb = __ldg(reinterpret_cast<const float4*>(p.ln2_g) + lane + 32);
        ba = __ldg(reinterpret_cast<const float4*>(p.ln2_b) + lane); bb = __ldg(reinterpret_cast<const float4*>(p.ln2_b) + lane + 32);
      }
#pragma unroll
      for (int i = 0; i < EP_ROWS; ++i) {
        const int o = cw + i * NUM_CW, t = t_out0 + o;
        if (t < Tn) {
          const float4* srow = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(stage_out) + o * OUT_LD);
          float4 v0 = srow[lane], v1 = srow[lane + 32];
          const int64_t goff = (static_cast<int64_t>(b) * Tn + t) * D;
          v0.x = (v0.x + r0[i].x) + b2a.x; v0.y = (v0.y + r0[i].y) + b2a.y; v0.z = (v0.z + r0[i].z) + b2a.z; v0.w = (v0.w + r0[i].w) + b2a.w;
          v1.x = (v1.x + r1[i].x) + b2b.x; v1.y = (v1.y + r1[i].y) + b2b.y; v1.z = (v1.z + r1[i].z) + b2b.z; v1.w = (v1.w + r1[i].w) + b2b.w;
          float4* orow = reinterpret_cast<float4*>(p.out + goff);
          orow[lane] = v0;
          orow[lane + 32] = v1;
          if (p.a_out != nullptr) {                                // the whole frame is in this warp: LayerNorm it
            const float mean = warp_sum((v0.x + v0.y) + (v0.z + v0.w) + (v1.x + v1.y) + (v1.z + v1.w)) * (1.0f / D);
            v0.x -= mean; v0.y -= mean; v0.z -= mean; v0.w -= mean;
            v1.x -= mean; v1.y -= mean; v1.z -= mean; v1.w -= mean;
            const float rstd = rsqrtf(warp_sum((v0.x * v0.x + v0.y * v0.y) + (v0.z * v0.z + v0.w * v0.w) +
                                               (v1.x * v1.x + v1.y * v1.y) + (v1.z * v1.z + v1.w * v1.w)) * (1.0f / D) + 1e-5f);
            uint2* arow = reinterpret_cast<uint2*>(p.a_out + goff);
            arow[lane] = make_uint2(pack_bf16x2(v0.x * rstd * ga.x + ba.x, v0.y * rstd * ga.y + ba.y),
                                    pack_bf16x2(v0.z * rstd * ga.z + ba.z, v0.w * rstd * ga.w + ba.w));
            arow[lane + 32] = make_uint2(pack_bf16x2(v1.x * rstd * gb.x + bb.x, v1.y * rstd * gb.y + bb.y),
                                         pack_bf16x2(v1.z * rstd * gb.z + bb.z, v1.w * rstd * gb.w + bb.w));
          }
        }
      }
    }
    CTRACE(9);
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

bool conv_module_fused_supported(int d, int ksize) { return d == D && ksize == KS; }

int conv_module_fused_bf16(const float* x, const float* res, float* out, const float* ln_g, const float* ln_b,
                           const bf16* W1p, const float* b1p, const float* dw_w, const float* dw_b, const bf16* W2,
                           const float* b2, const float* ln2_g, const float* ln2_b, bf16* a_out, int B, int Tn, int d,
                           int ksize, cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(d == D && ksize == KS, "fused conv module kernel is specialised for d_model = 256, kernel size 31");
  CFM_REQUIRE(B > 0 && Tn > 0 && B <= 65535, "bad conv module shape");
  CFM_REQUIRE(out != x, "fused conv module cannot run in place (tiles read a halo of their neighbours' frames)");
  CFM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(out) |
                reinterpret_cast<uintptr_t>(W1p) | reinterpret_cast<uintptr_t>(W2) | reinterpret_cast<uintptr_t>(ln_g) |
                reinterpret_cast<uintptr_t>(ln_b) | reinterpret_cast<uintptr_t>(b1p) | reinterpret_cast<uintptr_t>(dw_w) |
                reinterpret_cast<uintptr_t>(dw_b) | reinterpret_cast<uintptr_t>(b2)) & 15) == 0,
              "fused conv module operands must be 16-byte aligned");
  CUtensorMap tmW1, tmW2;
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(2 * D)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * 2};
    cuuint32_t box[2] = {KB, 256};
    CFM_TRY(encode_tmap(&tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W1p, dims, strides, box, "conv module W1"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(D)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * 2};
    cuuint32_t box[2] = {KB, 256};
    CFM_TRY(encode_tmap(&tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W2, dims, strides, box, "conv module W2"));
  }
  static std::atomic<uint64_t> attr_done{0};
  CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(conv_module_fused_kernel), SMEM_BYTES));
  CFM_REQUIRE(a_out == nullptr || (ln2_g != nullptr && ln2_b != nullptr), "a_out needs the following LayerNorm's parameters");
  CFM_REQUIRE(((reinterpret_cast<uintptr_t>(a_out) | reinterpret_cast<uintptr_t>(ln2_g) | reinterpret_cast<uintptr_t>(ln2_b)) & 15) == 0,
              "fused conv module operands must be 16-byte aligned");
  ConvParams p{x, res, ln_g, ln_b, b1p, dw_w, dw_b, b2, out, ln2_g, ln2_b, a_out, Tn, g_gemm_trace};
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, TO)), static_cast<unsigned>(B));
  CFM_CUDA_OK(launch_pdl(conv_module_fused_kernel, grid, dim3(NUM_THREADS), SMEM_BYTES, stream, 1, tmW1, tmW2, p));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
