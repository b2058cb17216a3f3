// bf16 GEMM on the 5th-generation tensor cores: C[M,N] = epi(A[M,K] * W[N,K]^T), fp32 accumulate.
//
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a multi-stage shared-memory
//     ring guarded by full/empty mbarriers;
//   * tcgen05.mma (kind::f16, M=128, N=BLOCK_N, K=16) issued by one thread, accumulators in TMEM,
//     double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * 8 epilogue warps read the accumulator with tcgen05.ld (one thread = one output row), apply
//     bias / Swish / ReLU / (residual + scale), and write the tile into a 128-byte-swizzled staging
//     buffer (conflict-free 16-byte stores) from which ONE thread issues TMA stores - global traffic
//     is fully coalesced and M/N tails are clipped by the tensor map.  The residual epilogue never reads
//     the residual: the tile scale*(acc+bias) leaves through a TMA *reduction* store (cp.reduce.async.bulk
//     .add.f32), i.e. the in-place update x += ... is performed by the L2;
//   * persistent: grid = min(#tiles, #SMs), static round-robin tile schedule, N-tiles innermost so the
//     CTAs running together share the same rows of A in L2.
//
// CONV mode turns the second front-end convolution (Conv2d(d,d,3,stride 2), convolution.py:39,46)
// into an implicit GEMM: the A tile for tap (kh,kw) is a strided box of conv1's channels-last output
// fetched by a 5-D tensor map (the stride-2 time axis is split into (pair, parity)), so no im2col
// buffer exists; its output tile goes out through a 4-D map over [B][T2][F2][C].
#include <mutex>

#include "ops.h"

namespace cfm {
namespace {

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 16;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = 128 + NUM_EPI_THREADS;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int CBLK_BYTES = BLOCK_M * 128;                      // one staging block: 128 rows x 128 bytes
constexpr int SMEM_LIMIT = 227 * 1024;

constexpr bool epi_out_f32(int epi) { return epi == EPI_BIAS_RES || epi == EPI_BIAS_F32; }

template <int BLOCK_N, int EPI, bool PAIR = false> struct Cfg {
  static constexpr bool OUT_F32 = epi_out_f32(EPI);
  static constexpr bool GLU = (EPI == EPI_BIAS_GLU);
  static constexpr int OUT_N = GLU ? BLOCK_N / 2 : BLOCK_N;    // output columns per tile
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2 / (PAIR ? 2 : 1);   // a CTA of a pair holds half the rows of B
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int CBLK_COLS = OUT_F32 ? 32 : 64;          // columns per staging block
  // a 256-wide fp32 tile (128 KB) is staged in two passes of 4 blocks: pass p carries columns 64 g + 32 p .. +31 of every
  // 64-column warp group g
  static constexpr bool F32_WIDE = OUT_F32 && BLOCK_N == 256;
  static constexpr int PASSES = F32_WIDE ? 2 : 1;
  static constexpr int NUM_CBLK = F32_WIDE ? 4 : OUT_N / CBLK_COLS;
  static constexpr int CBUF_BYTES = NUM_CBLK * CBLK_BYTES;
  // bf16 128-wide tiles can afford two staging buffers, so a TMA store drains while the next tile is written
  static constexpr int CBUF_COUNT = (!OUT_F32 && OUT_N == 128) ? 2 : 1;
  static constexpr int AUX_BYTES = 1024 /*alignment slack*/ + 256 /*barriers*/ + BLOCK_N * 4 /*bias*/;
  static constexpr int STAGES_FIT = (SMEM_LIMIT - CBUF_COUNT * CBUF_BYTES - AUX_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > (PAIR ? 8 : 6) ? (PAIR ? 8 : 6) : STAGES_FIT;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;                // two accumulator stages
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + CBUF_COUNT * CBUF_BYTES + AUX_BYTES;
  static_assert(STAGES >= 2, "not enough shared memory for a pipeline");
};

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_kb;
  int num_m_pairs;                     // PAIR: 256-row tiles
  int k_splits, kb_per_split;          // split-K (reduce-add epilogue only): tile = (split, m_tile, n_tile)
  const float* bias;
  float res_scale;
  // conv mode
  int T2, F2, TCH, kb_per_tap;
  // optional timeline of CTA 0 (debug): 8 globaltimer stamps per tile
  unsigned long long* trace;
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(slot)                                                                         \
  do {                                                                                      \
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0) p.trace[(tile / ncta) * 8 + (slot)] = gtime(); \
  } while (0)

// TF32: A and W are fp32 in memory and multiplied as TF32 (kind::tf32, K = 8 per MMA, 32 elements per 128-byte row);
// callers wanting fp32 accuracy pass operands split into TF32-exact high and low parts concatenated along K (3xTF32).
// SPLITK (reduce-add epilogue only): the tile index gets a third component, the K split; off for every inference GEMM so
// that their tile loops stay free of the extra index arithmetic.
// TN: C[M,N] = A^T B with A [K rows, M cols] and B [K rows, N cols] row-major (the weight-gradient form dW = dY^T X, whose
// contraction runs over the rows).  Both operands are then MN-major: a stage holds [64 k-rows][64 mn] slabs of 8 KB
// (128-byte rows, one swizzle atom along MN), consecutive slabs = consecutive 64-wide MN atoms (descriptor LBO = 8 KB,
// SBO = 1 KB = 8 k-rows), and one K = 16 MMA step advances the start address by 16 rows = 2 KB.
// PAIR: the kernel runs on CTA pairs (cluster of 2, tcgen05 cta_group::2).  A pair owns a 256 x BLOCK_N tile: each CTA
// loads its own 128 rows of A and HALF the rows of the W tile, the pair's leader issues M = 256 MMAs for both SMs, every
// CTA drains its own 128 accumulator rows.  Per 64-wide K block a CTA then moves 16 + BLOCK_N/16 KB through its
// shared-memory port twice (TMA fill + MMA read) instead of 16 + BLOCK_N/8 KB: with BLOCK_N = 256 that is 125 B/clk
// against the 128 B/clk port - the single-CTA form needs 188 B/clk and tops out at ~70 % of the tensor pipe on long
// contractions.  Barriers the issuing warp waits on live in the leader (the peer's TMA loads and epilogue warps reach
// them through cluster addresses); MMA completions are multicast to both CTAs.
template <int BLOCK_N, int EPI, bool CONV, bool TF32 = false, bool SPLITK = false, bool TN = false, bool PAIR = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, GemmParams p) {
  constexpr bool OUT_F32 = epi_out_f32(EPI);
  constexpr bool GLU = (EPI == EPI_BIAS_GLU);
  static_assert(!GLU || BLOCK_N == 256, "the GLU epilogue pairs two 32-column chunks per warp");
  static_assert(!PAIR || (!TF32 && !SPLITK && !TN), "CTA pairs: plain bf16 GEMM and the implicit-GEMM convolution");
  using C_ = Cfg<BLOCK_N, EPI, PAIR>;
  constexpr int STAGES = C_::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* cbuf = smem + STAGES * C_::STAGE_BYTES;
  float* bias_s = reinterpret_cast<float*>(cbuf + C_::CBUF_COUNT * C_::CBUF_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + BLOCK_N);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = PAIR ? p.num_m_pairs * p.num_n_tiles : p.num_m_tiles * p.num_n_tiles * (SPLITK ? p.k_splits : 1);
  constexpr int BK = TF32 ? BLOCK_K / 2 : BLOCK_K;         // elements per 128-byte operand row
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0 = the pair's leader
  const int cta = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int ncta = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], PAIR ? 2 * NUM_EPI_WARPS : NUM_EPI_WARPS);
    }
    mbar_init(res_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {                                         // whole TMEM: the CTA owns the SM, base is 0
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();                  // both CTAs' barriers exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_whole_sm(tmem_slot);
  pdl_wait();                                              // set-up done; the operands exist once the previous kernel has

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t it = 0;
    for (int tile = cta; tile < num_tiles; tile += ncta) {
      const int mn = SPLITK ? tile % (p.num_m_tiles * p.num_n_tiles) : tile, ks = SPLITK ? tile / (p.num_m_tiles * p.num_n_tiles) : 0;
      const int m_tile = PAIR ? 2 * (mn / p.num_n_tiles) + static_cast<int>(rank) : mn / p.num_n_tiles;
      const int n_tile = mn - (mn / p.num_n_tiles) * p.num_n_tiles;
      const int kb0 = SPLITK ? ks * p.kb_per_split : 0, kb1 = SPLITK ? min(p.num_kb, kb0 + p.kb_per_split) : p.num_kb;
      int cb = 0, cf = 0, ct = 0;
      if (CONV) {
        ct = m_tile % p.TCH;
        cf = (m_tile / p.TCH) % p.F2;
        cb = m_tile / (p.TCH * p.F2);
      }
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (kb == kb0) TRACE(0);
        if (kb == kb1 - 1) TRACE(1);
        if (lane == 0) {
          uint8_t* sa = smem + s * C_::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          if constexpr (PAIR) {
            // own rows of A, own half of the W tile; every byte of both CTAs completes on the leader's barrier
            const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
            if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * C_::STAGE_BYTES);
            if (CONV) {
              const int tap = kb / p.kb_per_tap, c0 = (kb - tap * p.kb_per_tap) * BLOCK_K;
              const int kw = tap / 3, kh = tap - kw * 3;
              tma_load_5d_pair(sa, &tmA, bar, c0, 2 * cf + kh, kw & 1, ct * BLOCK_M + (kw >> 1), cb);
            } else {
              tma_load_2d_pair(sa, &tmA, bar, kb * BK, m_tile * BLOCK_M);
            }
            tma_load_2d_pair(sb, &tmB, bar, kb * BK, n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2));
          } else {
          mbar_expect_tx(&full_bar[s], C_::STAGE_BYTES);
          if (CONV) {
            const int tap = kb / p.kb_per_tap, c0 = (kb - tap * p.kb_per_tap) * BLOCK_K;
            const int kw = tap / 3, kh = tap - kw * 3;
            // dims: (c, f1, parity, pair, b); source time index 2*t2 + kw = 2*(t2 + (kw>>1)) + (kw&1)
            tma_load_5d(sa, &tmA, &full_bar[s], c0, 2 * cf + kh, kw & 1, ct * BLOCK_M + (kw >> 1), cb);
          } else if (TN) {
#pragma unroll
            for (int sl = 0; sl < BLOCK_M / 64; ++sl) tma_load_2d(sa + sl * 8192, &tmA, &full_bar[s], m_tile * BLOCK_M + sl * 64, kb * BLOCK_K);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m_tile * BLOCK_M);
          }
          if (TN) {
#pragma unroll
            for (int sl = 0; sl < BLOCK_N / 64; ++sl) tma_load_2d(sb + sl * 8192, &tmB, &full_bar[s], n_tile * BLOCK_N + sl * 64, kb * BLOCK_K);
          } else {
            tma_load_2d(sb, &tmB, &full_bar[s], kb * BK, n_tile * BLOCK_N);
          }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (PAIR: the leader's, for both SMs)
    constexpr uint32_t idesc = TF32 ? umma_idesc_tf32(BLOCK_M, BLOCK_N)
                                    : (TN ? (umma_idesc_bf16(BLOCK_M, BLOCK_N) | (1u << 15) | (1u << 16))
                                          : umma_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, BLOCK_N));
    uint32_t it = 0, acc_it = 0;
    for (int tile = cta; tile < num_tiles; tile += ncta, ++acc_it) {
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aph ^ 1);                     // epilogue has drained this accumulator
      tc_fence_after();
      TRACE(2);
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      const int ks = SPLITK ? tile / (p.num_m_tiles * p.num_n_tiles) : 0;
      const int kb0 = SPLITK ? ks * p.kb_per_split : 0, kb1 = SPLITK ? min(p.num_kb, kb0 + p.kb_per_split) : p.num_kb;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (kb == kb0) TRACE(3);
        if (kb == kb1 - 1) TRACE(4);
        if (elect_one()) {
          // descriptor low word = (address >> 4) | LBO field; MN-major operands: LBO = 8 KB between 64-wide MN atoms
          const uint32_t sa = TN ? (((smem_u32(smem + s * C_::STAGE_BYTES) & 0x3FFFFu) >> 4) | ((8192u >> 4) << 16))
                                 : umma_desc_lo(smem_u32(smem + s * C_::STAGE_BYTES));
          const uint32_t sb = sa + (A_STAGE_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {         // one MMA covers 32 bytes of K (K-major) or 16 k-rows = 2 KB (MN-major)
            if (TN) tc_mma_f16_lo(d_tmem, sa + k * (2048 >> 4), sb + k * (2048 >> 4), idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            else if (TF32) tc_mma_tf32_lo(d_tmem, sa + k * (UMMA_K * 2 >> 4), sb + k * (UMMA_K * 2 >> 4), idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            else if (PAIR) tc_mma_f16_lo_pair(d_tmem, sa + k * (UMMA_K * 2 >> 4), sb + k * (UMMA_K * 2 >> 4), idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
            else tc_mma_f16_lo(d_tmem, sa + k * (UMMA_K * 2 >> 4), sb + k * (UMMA_K * 2 >> 4), idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          if constexpr (PAIR) {
            tc_commit_pair(&empty_bar[s]);                     // both CTAs' slots are free once these MMAs retire
            if (kb == kb1 - 1) tc_commit_pair(&tfull_bar[as]);
          } else {
          tc_commit(&empty_bar[s]);                            // smem slot free once these MMAs retire
          if (kb == kb1 - 1) tc_commit(&tfull_bar[as]);        // accumulator complete
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (16 warps)
    // warp -> TMEM lane quadrant (warp % 4, a hardware restriction) x column group (BLOCK_N / 4 columns)
    const int ew = warp - 4;
    const int q = warp & 3;
    const int cgrp = ew >> 2;
    constexpr int COLS_PER_WARP = BLOCK_N / 4;
    constexpr int NCH = COLS_PER_WARP / 32;                    // 32-column chunks per warp (1 or 2)
    const int row = q * 32 + lane;                             // row inside the tile
    const int tid_e = threadIdx.x - 128;
    const bool leader = (tid_e == 0);
    const uint32_t row_off = row * 128;
    const int sw = row & 7;
    uint32_t acc_it = 0;
    const uint32_t tempty_remote[2] = {PAIR ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u, PAIR ? mapa_u32(smem_u32(&tempty_bar[1]), 0) : 0u};
    for (int tile = cta; tile < num_tiles; tile += ncta, ++acc_it) {
      const int mn = SPLITK ? tile % (p.num_m_tiles * p.num_n_tiles) : tile, ks = SPLITK ? tile / (p.num_m_tiles * p.num_n_tiles) : 0;
      const int m_tile = PAIR ? 2 * (mn / p.num_n_tiles) + static_cast<int>(rank) : mn / p.num_n_tiles;
      const int n_tile = mn - (mn / p.num_n_tiles) * p.num_n_tiles;
      const int n0 = n_tile * BLOCK_N;                        // first GEMM column of the tile
      const int o0 = n_tile * C_::OUT_N;                      // first OUTPUT column of the tile
      const int out_n = GLU ? (p.N >> 1) : p.N;
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      uint8_t* cb = cbuf + (C_::CBUF_COUNT == 2 ? (acc_it & 1) * C_::CBUF_BYTES : 0);
      // The staging buffer must have been drained by the TMA store that last used it.  Single-pass epilogues wait for
      // that only after the accumulator has been pulled and converted in registers (LATE_WAIT below): a 64 KB store
      // takes ~1.7 us to leave shared memory, which used to sit in front of every tile's tcgen05.ld.
      constexpr bool LATE_WAIT = !GLU && C_::PASSES == 1;
      if (!LATE_WAIT && leader) {
        if (C_::CBUF_COUNT == 2) bulk_wait_read1(); else bulk_wait_read0();
      }
      if (tid_e < BLOCK_N) bias_s[tid_e] = (p.bias != nullptr && ks == 0 && n0 + tid_e < p.N) ? __ldg(p.bias + n0 + tid_e) : 0.f;
      named_bar_sync(1, NUM_EPI_THREADS);
      if (warp == 4) TRACE(5);
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      if (warp == 4) TRACE(6);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + cgrp * COLS_PER_WARP;
      uint32_t r[NCH][32];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) tmem_ld_32x32(t_row + ch * 32, r[ch]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {                                          // accumulator stage may be overwritten
        if constexpr (PAIR) mbar_arrive_cluster(tempty_remote[as]);
        else mbar_arrive(&tempty_bar[as]);
      }
      if (GLU) {
        // this warp's 64 GEMM columns are 32 value columns followed by their 32 gate columns
        const int gc = cgrp * COLS_PER_WARP;                   // GEMM column (in tile) of the value half
        const int oc = cgrp * 32;                              // output column in the tile
        if (o0 + oc < out_n) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float a = __uint_as_float(r[0][i]) + bias_s[gc + i];
            const float gte = __uint_as_float(r[NCH - 1][i]) + bias_s[gc + 32 + i];
            v[i] = a * sigmoid_tanh(gte);                      // activation.py:16-17
          }
          uint8_t* blk = cb + (oc >> 6) * CBLK_BYTES + row_off;
          const int j0 = (oc & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]); o.y = pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]);
            o.z = pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]); o.w = pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]);
            *reinterpret_cast<uint4*>(blk + (((j0 + j) ^ sw) << 4)) = o;
          }
        }
      }
#pragma unroll
      for (int pass = 0; pass < C_::PASSES; ++pass) {
      if (pass > 0) {                                          // the first pass has left the staging buffer
        if (leader) bulk_wait_read0();
        named_bar_sync(1, NUM_EPI_THREADS);
      }
      // phase A, registers only: bias, activation and scale; the results replace the accumulator bits in r
#pragma unroll
      for (int ch = 0; ch < (GLU ? 0 : NCH); ++ch) {
        if (C_::F32_WIDE && ch != pass) continue;
        const int col_in_tile = cgrp * COLS_PER_WARP + ch * 32;
        if (n0 + col_in_tile >= p.N) continue;
        float v[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[col_in_tile + g * 4]);   // smem broadcast
          v[g * 4 + 0] = __uint_as_float(r[ch][g * 4 + 0]) + b4.x;
          v[g * 4 + 1] = __uint_as_float(r[ch][g * 4 + 1]) + b4.y;
          v[g * 4 + 2] = __uint_as_float(r[ch][g * 4 + 2]) + b4.z;
          v[g * 4 + 3] = __uint_as_float(r[ch][g * 4 + 3]) + b4.w;
        }
        if (EPI == EPI_BIAS_SWISH) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = swish_tanh(v[i]);
        }
        if (EPI == EPI_BIAS_RELU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (EPI == EPI_BIAS_RES) {                             // the residual itself is added by the TMA reduction
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= p.res_scale;
        }
        if (OUT_F32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[ch][i] = __float_as_uint(v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) r[ch][i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
        }
      }
      if (LATE_WAIT) {
        if (leader) {
          if (C_::CBUF_COUNT == 2) bulk_wait_read1(); else bulk_wait_read0();
        }
        named_bar_sync(1, NUM_EPI_THREADS);
      }
      // phase B: registers -> staging blocks (128-byte rows, 16-byte chunks XOR-swizzled by the row)
#pragma unroll
      for (int ch = 0; ch < (GLU ? 0 : NCH); ++ch) {
        if (C_::F32_WIDE && ch != pass) continue;
        const int col_in_tile = cgrp * COLS_PER_WARP + ch * 32;
        if (n0 + col_in_tile >= p.N) continue;
        if (OUT_F32) {
          // 32 fp32 columns = one full 128-byte row of a staging block
          uint8_t* blk = cb + (C_::F32_WIDE ? cgrp : (col_in_tile >> 5)) * CBLK_BYTES + row_off;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(blk + ((j ^ sw) << 4)) = make_uint4(r[ch][j * 4], r[ch][j * 4 + 1], r[ch][j * 4 + 2], r[ch][j * 4 + 3]);
        } else {
          // 32 bf16 columns = half a 128-byte row of staging block (col_in_tile / 64)
          uint8_t* blk = cb + (col_in_tile >> 6) * CBLK_BYTES + row_off;
          const int j0 = (col_in_tile & 63) >> 3;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(blk + (((j0 + j) ^ sw) << 4)) = make_uint4(r[ch][j * 4], r[ch][j * 4 + 1], r[ch][j * 4 + 2], r[ch][j * 4 + 3]);
        }
      }
      fence_proxy_async();                                     // staging writes -> visible to the TMA engine
      named_bar_sync(1, NUM_EPI_THREADS);
      if (leader) {
        for (int blk = 0; blk < C_::NUM_CBLK; ++blk) {
          const int col = C_::F32_WIDE ? o0 + blk * 64 + pass * 32 : o0 + blk * C_::CBLK_COLS;
          if (col >= out_n) continue;
          if (CONV) {
            const int ct = m_tile % p.TCH, cf = (m_tile / p.TCH) % p.F2, cbi = m_tile / (p.TCH * p.F2);
            tma_store_4d(&tmC, cb + blk * CBLK_BYTES, col, cf, ct * BLOCK_M, cbi);
          } else if (EPI == EPI_BIAS_RES) {
            tma_reduce_add_2d(&tmC, cb + blk * CBLK_BYTES, col, m_tile * BLOCK_M);   // C += scale * (acc + bias)
          } else {
            tma_store_2d(&tmC, cb + blk * CBLK_BYTES, col, m_tile * BLOCK_M);
          }
        }
        bulk_commit();
      }
      }
      if (warp == 4) TRACE(7);
    }
    if (leader) bulk_wait0();                                   // all output tiles have left shared memory
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();                  // the peer's shared memory and barriers stay valid until both are done
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
}  // namespace
int encode_tmap(CUtensorMap* tm, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                const cuuint64_t* strides, const cuuint32_t* box, const char* what);
namespace {
std::once_flag g_once;
int g_init_status = -1;

using ::cfm::encode_tmap;
int encode_2d_bf16(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                   uint32_t box_inner, uint32_t box_outer, const char* what) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, what);
}

int encode_conv_5d(CUtensorMap* tm, const void* base, const ConvGeom& g) {
  // conv1 output [B][T1p][F1][C] viewed as (c, f1, parity, pair, b); T1p is even.
  const uint64_t C = g.C, F1 = g.F1;
  cuuint64_t dims[5] = {C, F1, 2, static_cast<cuuint64_t>(g.T1p / 2), static_cast<cuuint64_t>(g.B)};
  cuuint64_t strides[4] = {C * 2, F1 * C * 2, 2 * F1 * C * 2, static_cast<cuuint64_t>(g.T1p) * F1 * C * 2};
  cuuint32_t box[5] = {BLOCK_K, 1, 1, BLOCK_M, 1};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, "conv A");
}

template <int BLOCK_N, int EPI, bool CONV, bool SPLITK = false, bool PAIR = false>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p, int num_sms,
           cudaStream_t stream) {
  using C_ = Cfg<BLOCK_N, EPI, PAIR>;
  auto kern = gemm_tc_kernel<BLOCK_N, EPI, CONV, false, SPLITK, false, PAIR>;
  static std::atomic<uint64_t> attr_done{0};        // one per template instantiation
  CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(kern), C_::SMEM_BYTES));
  if (PAIR) {
    const int tiles = p.num_m_pairs * p.num_n_tiles;
    const int pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
    CFM_CUDA_OK(launch_pdl(kern, dim3(static_cast<unsigned>(2 * pairs)), dim3(NUM_THREADS), C_::SMEM_BYTES, stream, 2, tmA,
                           tmB, tmC, p));
    return 0;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles * p.k_splits;
  const int grid = tiles < num_sms ? tiles : num_sms;
  CFM_CUDA_OK(launch_pdl(kern, dim3(static_cast<unsigned>(grid)), dim3(NUM_THREADS), C_::SMEM_BYTES, stream, 1, tmA, tmB, tmC, p));
  return 0;
}

template <int BLOCK_N, bool CONV, bool PAIR>
int launch_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p, Epilogue epi,
               int num_sms, cudaStream_t stream) {
  switch (epi) {
    case EPI_BIAS: return launch<BLOCK_N, EPI_BIAS, CONV, false, PAIR>(tmA, tmB, tmC, p, num_sms, stream);
    case EPI_BIAS_SWISH: return launch<BLOCK_N, EPI_BIAS_SWISH, CONV, false, PAIR>(tmA, tmB, tmC, p, num_sms, stream);
    case EPI_BIAS_RELU: return launch<BLOCK_N, EPI_BIAS_RELU, CONV, false, PAIR>(tmA, tmB, tmC, p, num_sms, stream);
    case EPI_BIAS_GLU:
      if (BLOCK_N == 256 && !CONV) return launch<256, EPI_BIAS_GLU, false, false, PAIR>(tmA, tmB, tmC, p, num_sms, stream);
      break;
    default: break;
  }
  set_error("unsupported epilogue for this tile shape");
  return 1;
}
int launch_epi_f32(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                   Epilogue epi, bool pair, bool wide, int num_sms, cudaStream_t stream) {
  if (epi == EPI_BIAS_RES && p.k_splits > 1) return launch<128, EPI_BIAS_RES, false, true>(tmA, tmB, tmC, p, num_sms, stream);
  if (pair && wide) {
    if (epi == EPI_BIAS_RES) return launch<256, EPI_BIAS_RES, false, false, true>(tmA, tmB, tmC, p, num_sms, stream);
    return launch<256, EPI_BIAS_F32, false, false, true>(tmA, tmB, tmC, p, num_sms, stream);
  }
  if (pair) {
    if (epi == EPI_BIAS_RES) return launch<128, EPI_BIAS_RES, false, false, true>(tmA, tmB, tmC, p, num_sms, stream);
    return launch<128, EPI_BIAS_F32, false, false, true>(tmA, tmB, tmC, p, num_sms, stream);
  }
  if (epi == EPI_BIAS_RES) return launch<128, EPI_BIAS_RES, false>(tmA, tmB, tmC, p, num_sms, stream);
  return launch<128, EPI_BIAS_F32, false>(tmA, tmB, tmC, p, num_sms, stream);
}

}  // namespace

int encode_tmap(CUtensorMap* tm, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
           const cuuint64_t* strides, const cuuint32_t* box, const char* what) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(tm, dt, rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error(std::string("cuTensorMapEncodeTiled(") + what + ") failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 3;
  }
  return 0;
}


int tc_init() {
  std::call_once(g_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      set_error(std::string("cuTensorMapEncodeTiled is not available from the driver: ") + cudaGetErrorString(e));
      g_init_status = 3;
      return;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    g_init_status = 0;
  });
  if (g_init_status != 0) set_error("tensor-map encoder unavailable (no CUDA driver?)");
  return g_init_status;
}

unsigned long long* g_gemm_trace = nullptr;      // debug hook (cfm_debug_set_gemm_trace)
std::atomic<int> g_pdl{1};                       // programmatic dependent launch of the forward kernels (cfm_set_pdl)

int gemm_bf16_tn(const bf16* A, int lda, const bf16* B, int ldb, float* C, int ldc, int M, int N, int K, int num_sms,
                 cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM");
  CFM_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldc % 4 == 0, "TN GEMM needs 16-byte aligned rows");
  CFM_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
              "operands must be 16-byte aligned");
  constexpr int BN = 128;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = nullptr; p.res_scale = 1.f; p.trace = nullptr;
  p.num_kb = static_cast<int>(ceil_div(K, BLOCK_K));
  p.k_splits = 1;
  p.kb_per_split = p.num_kb;
  p.num_m_tiles = static_cast<int>(ceil_div(M, BLOCK_M));
  p.num_n_tiles = static_cast<int>(ceil_div(N, BN));
  const int mn_tiles = p.num_m_tiles * p.num_n_tiles;
  if (mn_tiles * 2 <= num_sms && p.num_kb >= 16) {
    int splits = num_sms / mn_tiles;
    if (splits > p.num_kb / 8) splits = p.num_kb / 8;
    if (splits > 1) {
      p.kb_per_split = static_cast<int>(ceil_div(p.num_kb, splits));
      p.k_splits = static_cast<int>(ceil_div(p.num_kb, p.kb_per_split));
    }
  }
  CUtensorMap tmA, tmB, tmC;
  CFM_TRY(encode_2d_bf16(&tmA, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), 64, BLOCK_K, "TN A"));
  CFM_TRY(encode_2d_bf16(&tmB, B, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldb), 64, BLOCK_K, "TN B"));
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldc) * 4};
    cuuint32_t box[2] = {32, BLOCK_M};
    CFM_TRY(encode_tmap(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dims, strides, box, "TN C"));
  }
  using C_ = Cfg<BN, EPI_BIAS_RES>;
  const int tiles = mn_tiles * p.k_splits;
  const int grid = tiles < num_sms ? tiles : num_sms;
  if (p.k_splits > 1) {
    auto kern = gemm_tc_kernel<BN, EPI_BIAS_RES, false, false, true, true>;
    static std::atomic<uint64_t> done{0};
    CFM_TRY(smem_attr_once(done, reinterpret_cast<const void*>(kern), C_::SMEM_BYTES));
    CFM_CUDA_OK(launch_pdl(kern, dim3(static_cast<unsigned>(grid)), dim3(NUM_THREADS), C_::SMEM_BYTES, stream, 1, tmA, tmB, tmC, p));
  } else {
    auto kern = gemm_tc_kernel<BN, EPI_BIAS_RES, false, false, false, true>;
    static std::atomic<uint64_t> done{0};
    CFM_TRY(smem_attr_once(done, reinterpret_cast<const void*>(kern), C_::SMEM_BYTES));
    CFM_CUDA_OK(launch_pdl(kern, dim3(static_cast<unsigned>(grid)), dim3(NUM_THREADS), C_::SMEM_BYTES, stream, 1, tmA, tmB, tmC, p));
  }
  return 0;
}

int gemm_tf32_tc(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N, int K,
                 int num_sms, cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM");
  CFM_REQUIRE(lda % 4 == 0 && ldw % 4 == 0 && ldc % 4 == 0 && N % 4 == 0, "TF32 GEMM needs 16-byte aligned rows");
  CFM_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C) |
                reinterpret_cast<uintptr_t>(bias)) & 15) == 0, "operands must be 16-byte aligned");
  constexpr int BK = BLOCK_K / 2, BN = 128;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.res_scale = 1.f; p.trace = nullptr;
  p.num_kb = static_cast<int>(ceil_div(K, BK));
  p.k_splits = 1;
  p.kb_per_split = p.num_kb;
  p.num_m_tiles = static_cast<int>(ceil_div(M, BLOCK_M));
  p.num_n_tiles = static_cast<int>(ceil_div(N, BN));
  CUtensorMap tmA, tmB, tmC;
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(lda) * 4};
    cuuint32_t box[2] = {BK, BLOCK_M};
    CFM_TRY(encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, A, dims, strides, box, "tf32 A"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(N)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldw) * 4};
    cuuint32_t box[2] = {BK, BN};
    CFM_TRY(encode_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, W, dims, strides, box, "tf32 W"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldc) * 4};
    cuuint32_t box[2] = {32, BLOCK_M};
    CFM_TRY(encode_tmap(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dims, strides, box, "tf32 C"));
  }
  using C_ = Cfg<BN, EPI_BIAS_F32>;
  auto kern = gemm_tc_kernel<BN, EPI_BIAS_F32, false, true>;
  static std::atomic<uint64_t> attr_done{0};
  CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(kern), C_::SMEM_BYTES));
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  CFM_CUDA_OK(launch_pdl(kern, dim3(static_cast<unsigned>(tiles < num_sms ? tiles : num_sms)), dim3(NUM_THREADS), C_::SMEM_BYTES,
                         stream, 1, tmA, tmB, tmC, p));
  return 0;
}

int gemm_bf16_tc(const bf16* A, int lda, const bf16* W, const float* bias, void* C, int ldc, int M, int N, int K,
                 Epilogue epi, const float* res, float res_scale, const ConvGeom* conv, int num_sms,
                 cudaStream_t stream, bool split_k) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM");
  CFM_REQUIRE(K % 8 == 0 && N % 8 == 0 && ldc % 8 == 0, "tensor-core GEMM needs K, N, ldc multiples of 8");
  CFM_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
              "operands must be 16-byte aligned");
  const bool out_f32 = epi_out_f32(epi);
  // the residual is read through the output tensor map: the update is in place (x += scale * (A W^T + b))
  CFM_REQUIRE(epi != EPI_BIAS_RES || res == static_cast<const float*>(C), "residual epilogue updates C in place (res must alias C)");
  CFM_REQUIRE(conv == nullptr || !out_f32, "conv mode has a bf16 output");
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.res_scale = res_scale; p.trace = g_gemm_trace;
  p.num_kb = static_cast<int>(ceil_div(K, BLOCK_K));
  p.k_splits = 1;
  p.kb_per_split = p.num_kb;
  CUtensorMap tmA, tmB, tmC;
  if (conv != nullptr) {
    CFM_REQUIRE(conv->C % BLOCK_K == 0, "implicit-GEMM conv needs channels to be a multiple of 64");
    CFM_REQUIRE(conv->T1p % 2 == 0 && K == 9 * conv->C && N == conv->C && ldc == conv->C, "bad conv geometry");
    CFM_REQUIRE(M == conv->B * conv->T2 * conv->F2, "conv GEMM row count mismatch");
    p.T2 = conv->T2; p.F2 = conv->F2;
    p.TCH = static_cast<int>(ceil_div(conv->T2, BLOCK_M));
    p.kb_per_tap = conv->C / BLOCK_K;
    p.num_m_tiles = conv->B * conv->F2 * p.TCH;
    CFM_TRY(encode_conv_5d(&tmA, A, *conv));
    // output [B][T2][F2][C] as (c, f2, t2, b), one (128 t2) x (64 c) block per store
    const uint64_t Cc = conv->C, F2 = conv->F2, T2 = conv->T2;
    cuuint64_t dims[4] = {Cc, F2, T2, static_cast<cuuint64_t>(conv->B)};
    cuuint64_t strides[3] = {Cc * 2, F2 * Cc * 2, T2 * F2 * Cc * 2};
    cuuint32_t box[4] = {64, 1, BLOCK_M, 1};
    CFM_TRY(encode_tmap(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, C, dims, strides, box, "conv C"));
  } else {
    CFM_REQUIRE(lda % 8 == 0, "lda must be a multiple of 8");
    p.num_m_tiles = static_cast<int>(ceil_div(M, BLOCK_M));
    CFM_TRY(encode_2d_bf16(&tmA, A, static_cast<uint64_t>(K), static_cast<uint64_t>(M), static_cast<uint64_t>(lda),
                           BLOCK_K, BLOCK_M, "A"));
    CFM_REQUIRE(epi != EPI_BIAS_GLU || N % 64 == 0, "GLU epilogue needs N to be a multiple of 64");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(epi == EPI_BIAS_GLU ? N / 2 : N), static_cast<cuuint64_t>(M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldc) * (out_f32 ? 4 : 2)};
    cuuint32_t box[2] = {out_f32 ? 32u : 64u, BLOCK_M};
    CFM_TRY(encode_tmap(&tmC, out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, dims,
                   strides, box, "C"));
  }
  // tile width: 256 halves the A re-reads, 128 quantises better over the SMs; pick the cheaper schedule
  auto cost = [&](int bn) {
    const int64_t tiles = static_cast<int64_t>(p.num_m_tiles) * ceil_div(N, bn);
    return ceil_div(tiles, num_sms) * bn;
  };
  // CTA pairs (cta_group::2) for the long contractions: the single-CTA main loop is bound by the shared-memory port
  // (see the kernel's header); GEMMs with K < 512 are bound by their epilogue and stay as they are
  const bool pair_ok = !(epi == EPI_BIAS_RES && split_k) && p.num_kb >= 8 && p.num_m_tiles >= 2 && num_sms >= 2;
  const bool use256 = epi == EPI_BIAS_GLU || ((!out_f32 || pair_ok) && (N > 128) && (cost(256) <= cost(128)));
  const int bn = use256 ? 256 : 128;
  p.num_n_tiles = static_cast<int>(ceil_div(N, bn));
  p.num_m_pairs = (p.num_m_tiles + 1) / 2;
  if (epi == EPI_BIAS_RES && split_k) {
    // few output tiles and a long contraction (the weight-gradient GEMMs of the training step): split K over the idle SMs;
    // every split reduce-adds its partial product into C, the bias enters with split 0
    const int mn_tiles = p.num_m_tiles * p.num_n_tiles;
    if (mn_tiles * 2 <= num_sms && p.num_kb >= 16) {
      int splits = num_sms / mn_tiles;
      if (splits > p.num_kb / 8) splits = p.num_kb / 8;
      if (splits > 1) {
        p.kb_per_split = static_cast<int>(ceil_div(p.num_kb, splits));
        p.k_splits = static_cast<int>(ceil_div(p.num_kb, p.kb_per_split));
      }
    }
  }
  const bool pair = pair_ok;
  CFM_TRY(encode_2d_bf16(&tmB, W, static_cast<uint64_t>(K), static_cast<uint64_t>(N), static_cast<uint64_t>(K), BLOCK_K,
                         static_cast<uint32_t>(pair ? bn / 2 : bn), "W"));
  if (out_f32) return launch_epi_f32(tmA, tmB, tmC, p, epi, pair, use256, num_sms, stream);
  if (conv != nullptr) {
    if (pair)
      return use256 ? launch_epi<256, true, true>(tmA, tmB, tmC, p, epi, num_sms, stream)
                    : launch_epi<128, true, true>(tmA, tmB, tmC, p, epi, num_sms, stream);
    return use256 ? launch_epi<256, true, false>(tmA, tmB, tmC, p, epi, num_sms, stream)
                  : launch_epi<128, true, false>(tmA, tmB, tmC, p, epi, num_sms, stream);
  }
  if (pair)
    return use256 ? launch_epi<256, false, true>(tmA, tmB, tmC, p, epi, num_sms, stream)
                  : launch_epi<128, false, true>(tmA, tmB, tmC, p, epi, num_sms, stream);
  return use256 ? launch_epi<256, false, false>(tmA, tmB, tmC, p, epi, num_sms, stream)
                : launch_epi<128, false, false>(tmA, tmB, tmC, p, epi, num_sms, stream);
}

}  // namespace cfm
