// Fused FeedForwardModule half-step for d_model = 256 on tcgen05 tensor cores, on CTA PAIRS (cta_group::2):
//
//     x[M,256] += scale * ( swish( A W1^T + b1 ) W2^T + b2 ),  A = LN(x) as bf16      (ffn.py:15-23 + block.py:19,25)
//
// A pair of CTAs (cluster of 2 = the two SMs of a TPC) owns 256 rows, each CTA its own 128: A operand, hidden tile,
// accumulators and epilogues are per CTA, but every MMA is issued by the pair's leader for both SMs
// (tcgen05.mma.cta_group::2, M = 256) with the WEIGHT operand split between them: a CTA loads only half the rows of
// each weight tile, so the weight bytes written to and read from each SM's shared memory halve and the main loop is no
// longer bound by the 128 B/clk shared-memory port (per 128 hidden units: 192 KB against 2048 clk of MMA; the
// single-CTA form moved 320 KB).  Nothing but A, the weights and the updated x touches global memory:
//   prologue   A arrives already normalised (the kernel that produced x also emitted LN(x) as bf16: the block-final
//              dual LayerNorm, the fused conv module's epilogue, or a LayerNorm launch) and is TMA-loaded straight into
//              the 128B-swizzled K-major layout the MMA descriptors expect.  The kernel never READS x, so x may be
//              updated in place by any number of CTAs sharing a row tile;
//   main loop  the 1024 hidden units are processed in 8 chunks of 128:
//                GEMM1(c): acc1[c&1] (TMEM, 128 cols)  = LN(x) . W1[c]^T          (K = 256)
//                epilogue: hidden = bf16(swish(acc1 + b1)) written IN PLACE over acc1[c&1] in tensor memory
//                GEMM2(c): acc2 (TMEM, 256 cols)      += hidden . W2[:, c]^T      (K = 128, A operand from TMEM)
//              software-pipelined (GEMM1(c+1) is issued before GEMM2(c)) so the tensor pipe keeps working while the 16
//              epilogue warps run the Swish; W1/W2 tiles stream through a TMA/mbarrier ring, one stage per GEMM.
//              Measured (tools/ffn_trace.py): the MMAs of a chunk now take their 2048 clk, and the loop is paced by the
//              chain GEMM1 -> Swish epilogue (~1 us: tcgen05.ld, 32 MUFU per thread at 16/clk/SM, tcgen05.st, arrive on
//              the leader's barrier) -> GEMM2 over the two acc1 buffers: ~1.45 us per chunk.  Deeper pipelines need
//              more tensor memory than the SM has (4 x 64-wide buffers were tried: GEMM1 at N = 64 re-reads the A
//              operand from shared memory twice as often and the port binds again, 1.56 us per chunk);
//   epilogue   scale*(acc2 + b2) is staged in shared memory and leaves through a TMA reduction store, i.e. the residual
//              add is executed by the L2 (x is never re-read).
// Barriers the issuing warp waits on live in the leader; the peer's TMA loads and epilogue warps signal them through
// cluster addresses, and MMA completions are multicast to both CTAs.
// TMEM: 2 x 128 (acc1 double buffer) + 256 (acc2) = 512 columns.
#include "ops.h"

namespace cfm {
namespace {

constexpr int D = 256, HID = 4 * D, BM = 128, HC = 128, NCHUNK = HID / HC;   // items are ranges of 128-wide chunks
constexpr int KB = 64;                                     // K elements per 128-byte swizzled row
constexpr int XS_BYTES = (D / KB) * BM * 128;              // 64 KB: LN(x) as 4 k-blocks
constexpr int OUT_BYTES = 64 * 1024;                       // staging of half an output tile (128 rows x 128 fp32)
constexpr int STAGE_BYTES = 32 * 1024, STAGES = 3;         // one stage = this CTA's half of one chunk's W1 or W2 tile
constexpr int NBUF = 2;                                    // acc1 buffers
constexpr int NUM_EPI_WARPS = 16, NUM_EPI_THREADS = NUM_EPI_WARPS * 32, NUM_THREADS = 128 + NUM_EPI_THREADS;
constexpr int SMEM_BYTES = 1024 + XS_BYTES + OUT_BYTES + STAGES * STAGE_BYTES + 512;
static_assert(SMEM_BYTES <= 227 * 1024, "fused FFN shared memory budget");
constexpr int OUT_BLK_BYTES = BM * 128;                    // staging block: 128 rows x 32 fp32
static_assert(4 * OUT_BLK_BYTES <= OUT_BYTES, "output staging");

struct FfnParams {
  const float* b1;
  const float* b2;
  float scale;
  int M;
  int num_tiles;                 // 256-row pair tiles
  unsigned long long* trace;     // debug: timeline of pair 0's leader (ns)
};
__device__ __forceinline__ unsigned long long ffn_time_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Work items of one CTA pair.  With T row tiles on G pairs, every pair first takes T / G whole tiles; each of the T % G
// left-over tiles is then shared by G / (T % G) pairs, each taking a contiguous range of its 8 HIDDEN chunks, so the
// last wave is as full as the first (C2: 94 tiles of 256 rows on 74 pairs, 20 tiles shared by 3-4 pairs each).  Partial
// items add their share of the output with the same TMA reduction store (the L2 sums them); b2 is added by the item
// that owns chunk 0.
struct Item { int tile, cb, ce; };
struct ItemIter {
  int c, G, R, L, r;
  bool tail_done;
  __device__ ItemIter(int cta, int grid, int tiles) : c(cta), G(grid), R(tiles / grid), L(tiles % grid), r(0), tail_done(false) {}
  __device__ bool next(Item& it) {
    if (r < R) {
      it = Item{r * G + c, 0, NCHUNK};
      ++r;
      return true;
    }
    if (tail_done || L == 0) return false;
    tail_done = true;
    // left-over tile lt is shared by the pairs [lt*G/L, (lt+1)*G/L): one partial item per pair, no straddling
    const int lt = static_cast<int>(static_cast<long long>(c) * L / G);
    const int first = (lt * G + L - 1) / L, last = ((lt + 1) * G + L - 1) / L;     // pairs of this tile: [first, last)
    const int k = last - first, j = c - first;
    const int cb = j * NCHUNK / k, ce = (j + 1) * NCHUNK / k;
    if (cb >= ce) return false;
    it = Item{R * G + lt, cb, ce};
    return true;
  }
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmA, FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* xs = smem;
  uint8_t* stage_out = xs + XS_BYTES;
  uint8_t* ring = stage_out + OUT_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES);   // leader's: bytes of both CTAs
  uint64_t* empty_bar = full_bar + STAGES;                                           // own: multicast MMA completion
  uint64_t* xs_full = empty_bar + STAGES;                                            // leader's
  uint64_t* acc1_full = xs_full + 1;      // [NBUF] own
  uint64_t* hs_full = acc1_full + NBUF;   // [NBUF] leader's: the epilogue warps of BOTH CTAs
  uint64_t* acc2_full = hs_full + NBUF;   // own
  uint64_t* acc2_empty = acc2_full + 1;   // leader's: the epilogue warps of both CTAs have read acc2 of the previous item
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                   // 0 = leader
  const int pair = static_cast<int>(blockIdx.x >> 1), npairs = static_cast<int>(gridDim.x >> 1);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmA);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(xs_full, 1);
    for (int s = 0; s < NBUF; ++s) {
      mbar_init(&acc1_full[s], 1);
      mbar_init(&hs_full[s], 2 * NUM_EPI_WARPS);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 2 * NUM_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();                                        // both CTAs' barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_whole_sm(tmem_slot);
  const uint32_t tm_acc2 = tmem_base + NBUF * HC;
  pdl_wait();                                                // x / a exist once the previous kernel has completed

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer (each CTA: its half of every tile)
    // consumption order per item: G1(cb) G1(cb+1) G2(cb) G1(cb+2) G2(cb+1) ... G2(ce-1); one ring stage per GEMM
    uint32_t it = 0;
    auto acquire = [&]() -> uint8_t* {
      const int s = it % STAGES;
      mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
      return ring + s * STAGE_BYTES;
    };
    auto load_g1 = [&](int c) {                                // W1 rows c*128 + rank*64 .. +63, all 4 k-blocks
      uint8_t* dst = acquire();
      if (lane == 0) {
        const int s = it % STAGES;
        const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
        if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
#pragma unroll
        for (int kb = 0; kb < D / KB; ++kb)
          tma_load_2d_pair(dst + kb * 8192, &tmW1, bar, kb * KB, c * HC + static_cast<int>(rank) * (HC / 2));
      }
      __syncwarp();
      ++it;
    };
    auto load_g2 = [&](int c) {                                // W2 rows rank*128 .. +127, hidden columns c*128 .. +127
      uint8_t* dst = acquire();
      if (lane == 0) {
        const int s = it % STAGES;
        const uint32_t bar = mapa_u32(smem_u32(&full_bar[s]), 0);
        if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
#pragma unroll
        for (int s2 = 0; s2 < HC / KB; ++s2)
          tma_load_2d_pair(dst + s2 * 16384, &tmW2, bar, c * HC + s2 * KB, static_cast<int>(rank) * (D / 2));
      }
      __syncwarp();
      ++it;
    };
    ItemIter items(pair, npairs, p.num_tiles);
    Item w;
    while (items.next(w)) {
      load_g1(w.cb);
      for (int c = w.cb; c < w.ce; ++c) {
        if (c + 1 < w.ce) load_g1(c + 1);
        load_g2(c);
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (the leader's, for both SMs)
    constexpr uint32_t idesc1 = umma_idesc_bf16(2 * BM, HC);    // 256 x 128
    constexpr uint32_t idesc2 = umma_idesc_bf16(2 * BM, D);     // 256 x 256
    uint32_t it = 0, g1 = 0, g2 = 0;                           // ring slot counter, chunks issued to GEMM1 / GEMM2 so far
    unsigned long long w_full = 0, w_hs = 0, w_xs = 0, w_acc2 = 0, t_begin = ffn_time_ns();
    const bool tr = p.trace != nullptr && blockIdx.x == 0;
    auto gemm1 = [&]() {
      // acc1[b] last held the hidden tile of the chunk two back: its GEMM2 was issued before this GEMM1 and the tensor
      // pipe executes in issue order, and that GEMM2 was itself issued after the Swish warps had finished with the
      // buffer (hs_full) - no further handshake is needed
      const int b = g1 & (NBUF - 1), s = it % STAGES;
      unsigned long long t0 = tr ? ffn_time_ns() : 0;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      if (tr) w_full += ffn_time_ns() - t0;
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sb = umma_desc_lo(smem_u32(ring + s * STAGE_BYTES));
#pragma unroll
        for (int kb = 0; kb < D / KB; ++kb) {
          const uint32_t sa = umma_desc_lo(smem_u32(xs + kb * (BM * 128)));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16_lo_pair(tmem_base + b * HC, sa + k * 2, sb + kb * (8192 >> 4) + k * 2, idesc1, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit_pair(&empty_bar[s]);
        tc_commit_pair(&acc1_full[b]);
      }
      __syncwarp();
      ++it;
      ++g1;
    };
    auto gemm2 = [&](bool first, bool last, uint32_t n) {
      const int b = g2 & (NBUF - 1), s = it % STAGES;
      unsigned long long t0 = tr ? ffn_time_ns() : 0;
      if (first && n > 0) {                                 // the previous item's output has left the accumulator
        mbar_wait(acc2_empty, (n - 1) & 1);
        tc_fence_after();
      }
      if (tr) w_acc2 += ffn_time_ns() - t0;
      t0 = tr ? ffn_time_ns() : 0;
      mbar_wait(&hs_full[b], (g2 / NBUF) & 1);
      if (tr) w_hs += ffn_time_ns() - t0;
      t0 = tr ? ffn_time_ns() : 0;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      if (tr) w_full += ffn_time_ns() - t0;
      tc_fence_after();
      if (elect_one()) {
        // hidden units 16*kg .. +15 of the chunk (kg = 4*s2 + k): bf16 pairs in columns 32*(kg/2) + 8*(kg%2) .. +7 of acc1[b]
        const uint32_t sb = umma_desc_lo(smem_u32(ring + s * STAGE_BYTES));
#pragma unroll
        for (int s2 = 0; s2 < HC / KB; ++s2) {
          const uint32_t ta = tmem_base + b * HC + s2 * 64;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16_ts_pair(tm_acc2, ta + 32 * (k >> 1) + 8 * (k & 1), sb + s2 * (16384 >> 4) + k * 2, idesc2,
                               (!first || (s2 | k) != 0) ? 1u : 0u);
        }
        tc_commit_pair(&empty_bar[s]);
        if (last) tc_commit_pair(acc2_full);
      }
      __syncwarp();
      if (tr && lane == 0 && g2 < 24) p.trace[16 + g2] = ffn_time_ns() - t_begin;
      ++it;
      ++g2;
    };
    ItemIter items(pair, npairs, p.num_tiles);
    Item w;
    for (uint32_t n = 0; items.next(w); ++n) {
      unsigned long long t0 = tr ? ffn_time_ns() : 0;
      mbar_wait(xs_full, n & 1);                               // LN(x) of this item staged in both CTAs
      if (tr) w_xs += ffn_time_ns() - t0;
      tc_fence_after();
      gemm1();
      for (int c = w.cb; c < w.ce; ++c) {
        if (c + 1 < w.ce) gemm1();
        gemm2(c == w.cb, c == w.ce - 1, n);
      }
    }
    if (tr && lane == 0) {
      p.trace[0] = ffn_time_ns() - t_begin;
      p.trace[1] = w_xs;
      p.trace[2] = w_full;
      p.trace[3] = w_hs;
      p.trace[4] = w_acc2;
      p.trace[8] = t_begin;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ A-operand requests + epilogues (16 warps)
    const int ew = warp - 4;
    const int q = warp & 3, cgrp = ew >> 2;                     // TMEM lane quadrant (hardware rule: warp % 4); 32-column block
    const int tid_e = threadIdx.x - 128;
    const bool requester = tid_e == 0;
    const bool trw = p.trace != nullptr && blockIdx.x == 0 && tid_e == 0;
    ItemIter items(pair, npairs, p.num_tiles);
    Item w;
    uint32_t gc = 0;                                            // chunks processed so far (acc1 double buffer)
    const int row = q * 32 + lane;
    const int sw = row & 7;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p.b1 + 2 * tid_e));                 // the biases into this SM's L1
    if (tid_e < D / 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.b2 + 2 * tid_e));
    // barriers of the MMA-issuing CTA that this CTA's epilogue warps signal
    uint32_t hs_full_bar[NBUF];
#pragma unroll
    for (int s = 0; s < NBUF; ++s) hs_full_bar[s] = mapa_u32(smem_u32(&hs_full[s]), 0);
    const uint32_t acc2_empty_bar = mapa_u32(smem_u32(acc2_empty), 0);
    const uint32_t xs_full_bar = mapa_u32(smem_u32(xs_full), 0);
    // A operand of an item = four [128 rows x 64] k-blocks straight into the swizzled MMA layout (rows past M arrive
    // as zeros), each CTA its own rows, all bytes completing on the leader's barrier.  It is requested as soon as the
    // previous item's last GEMM1 has finished reading xs, i.e. a chunk and a whole output epilogue ahead of its
    // first use: the next item's first GEMM1s then run under this item's epilogue.
    auto request_xs = [&](int tile) {
      if (rank == 0) mbar_expect_tx(xs_full, 2 * XS_BYTES);
#pragma unroll
      for (int kb = 0; kb < D / KB; ++kb)
        tma_load_2d_pair(xs + kb * (BM * 128), &tmA, xs_full_bar, kb * KB, tile * 2 * BM + static_cast<int>(rank) * BM);
    };
    Item w_next;
    bool have = items.next(w);
    if (have && requester) request_xs(w.tile);
    for (uint32_t n = 0; have; ++n) {
      const bool have_next = items.next(w_next);
      const int m0 = w.tile * 2 * BM + static_cast<int>(rank) * BM;
      // ---- per-chunk epilogue: hidden = bf16(swish(acc1 + b1)) over the accumulator, the TMEM-resident A operand of GEMM2
      for (int c = w.cb; c < w.ce; ++c, ++gc) {
        const int b = gc & (NBUF - 1);
        const float4* b1v = reinterpret_cast<const float4*>(p.b1 + c * HC + cgrp * 32);
        float4 bc[8];                                            // halved biases, ready before the wait
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          bc[g] = __ldg(b1v + g);
          bc[g].x *= 0.5f; bc[g].y *= 0.5f; bc[g].z *= 0.5f; bc[g].w *= 0.5f;
        }
        mbar_wait(&acc1_full[b], (gc / NBUF) & 1);
        tc_fence_after();
        if (c == w.ce - 1 && have_next && requester) request_xs(w_next.tile);   // every GEMM1 of this item is done
        uint32_t r[32];
        const uint32_t t_acc1 = tmem_base + lane_addr + b * HC + cgrp * 32;
        tmem_ld_32x32(t_acc1, r);
        tmem_ld_wait();
        // hidden units cgrp*32 .. +31 of the chunk as bf16 pairs over the first 16 of the 32 accumulator columns this thread
        // has just read (only this thread touches them); swish(x) = h + h tanh(h), h = x/2: one MUFU op per element
        uint32_t pk[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b4 = bc[g];
          const float h0 = fmaf(__uint_as_float(r[g * 4 + 0]), 0.5f, b4.x), h1 = fmaf(__uint_as_float(r[g * 4 + 1]), 0.5f, b4.y);
          const float h2 = fmaf(__uint_as_float(r[g * 4 + 2]), 0.5f, b4.z), h3 = fmaf(__uint_as_float(r[g * 4 + 3]), 0.5f, b4.w);
          float t0, t1, t2, t3;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t2) : "f"(h2));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t3) : "f"(h3));
          pk[2 * g] = pack_bf16x2(fmaf(h0, t0, h0), fmaf(h1, t1, h1));
          pk[2 * g + 1] = pack_bf16x2(fmaf(h2, t2, h2), fmaf(h3, t3, h3));
        }
        tmem_st_32x16(t_acc1, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(hs_full_bar[b]);
        if (trw && gc < 24) p.trace[48 + gc] = ffn_time_ns();
      }
      // ---- final epilogue: x += scale * (acc2 + b2) through a TMA reduction store
      mbar_wait(acc2_full, n & 1);
      tc_fence_after();
      if (trw && n < 4) p.trace[80 + n] = ffn_time_ns();
      {
        // the tile is staged in two halves of 128 columns, and acc2 is handed back to the MMA warp as soon as the second
        // half is in registers.  The second half waits for the TMA engine to have read the first out of the staging
        // buffer - except in a CTA's last item, where the A-operand buffer is dead and takes it
        if (tid_e == 0) bulk_wait_read0();                         // the previous item has left the staging
        named_bar_sync(1, NUM_EPI_THREADS);
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          const int col = hh * 128 + cgrp * 32;
          uint8_t* stage = (hh == 1 && !have_next) ? xs : stage_out;
          uint32_t r[32];
          tmem_ld_32x32(tm_acc2 + lane_addr + col, r);
          tmem_ld_wait();
          if (hh == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc2_empty_bar);
          }
          uint8_t* blk = stage + cgrp * OUT_BLK_BYTES + row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (w.cb == 0) b4 = __ldg(reinterpret_cast<const float4*>(p.b2 + col) + j);     // bias once per row tile
            float4 o;
            o.x = p.scale * (__uint_as_float(r[j * 4 + 0]) + b4.x);
            o.y = p.scale * (__uint_as_float(r[j * 4 + 1]) + b4.y);
            o.z = p.scale * (__uint_as_float(r[j * 4 + 2]) + b4.z);
            o.w = p.scale * (__uint_as_float(r[j * 4 + 3]) + b4.w);
            *reinterpret_cast<float4*>(blk + ((j ^ sw) << 4)) = o;
          }
          fence_proxy_async();
          named_bar_sync(1, NUM_EPI_THREADS);
          if (tid_e == 0) {
            tma_reduce_add_3d(&tmX, stage, 0, m0, hh * 4);            // 4 blocks of [128 rows x 32 columns]
            bulk_commit();
            if (hh == 0 && have_next) bulk_wait_read0();             // the first half has left the staging
          }
          if (hh == 0 && have_next) named_bar_sync(1, NUM_EPI_THREADS);
        }
      }
      if (trw && n < 4) p.trace[84 + n] = ffn_time_ns();
      w = w_next;
      have = have_next;
    }
    if (tid_e == 0) bulk_wait_read0();                   // the last half tile has left shared memory
  }
  tc_fence_before();
  cluster_sync_all();                                        // the peer's shared memory and barriers stay valid until both are done
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace

extern unsigned long long* g_gemm_trace;

bool ffn_fused_supported(int d) { return d == D; }

int ffn_fused_bf16(const bf16* a_pre, float* x, const bf16* W1, const float* b1, const bf16* W2, const float* b2, float scale,
                   int M, int d, int num_sms, cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(d == D, "fused FFN kernel is specialised for d_model = 256");
  CFM_REQUIRE(M > 0, "empty FFN");
  CFM_REQUIRE(num_sms >= 2, "the fused FFN kernel runs on CTA pairs");
  CFM_REQUIRE(a_pre != nullptr && x != nullptr, "fused FFN needs the normalised input and the residual buffer");
  CFM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(a_pre) | reinterpret_cast<uintptr_t>(W1) |
                reinterpret_cast<uintptr_t>(W2) | reinterpret_cast<uintptr_t>(b1) | reinterpret_cast<uintptr_t>(b2)) & 15) == 0,
              "fused FFN operands must be 16-byte aligned");
  CUtensorMap tmW1, tmW2, tmX, tmA;
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * 2};
    cuuint32_t box[2] = {KB, BM};
    CFM_TRY(encode_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a_pre, dims, strides, box, "ffn normalised input"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(HID)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * 2};
    cuuint32_t box[2] = {KB, HC / 2};
    CFM_TRY(encode_tmap(&tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W1, dims, strides, box, "ffn W1"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(HID), static_cast<cuuint64_t>(D)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(HID) * 2};
    cuuint32_t box[2] = {KB, D / 2};
    CFM_TRY(encode_tmap(&tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W2, dims, strides, box, "ffn W2"));
  }
  {  // x as [8 column blocks][M rows][32 columns]: one reduction store moves four staged [128 x 32] blocks
    cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(M), static_cast<cuuint64_t>(D / 32)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(D) * 4, 32 * 4};
    cuuint32_t box[3] = {32, BM, 4};
    CFM_TRY(encode_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, x, dims, strides, box, "ffn x"));
  }
  static std::atomic<uint64_t> attr_done{0};
  CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(ffn_fused_kernel), SMEM_BYTES));
  // Persistent CTA pairs, left-over row tiles shared along the hidden dimension; a share's prologue is one 64 KB TMA load
  // per CTA.  Sharing is safe because the kernel only ever ADDS to x.
  const int tiles = static_cast<int>(ceil_div(M, 2 * BM));
  const int pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
  FfnParams p{b1, b2, scale, M, tiles, g_gemm_trace};
  CFM_CUDA_OK(launch_pdl(ffn_fused_kernel, dim3(static_cast<unsigned>(2 * pairs)), dim3(NUM_THREADS), SMEM_BYTES, stream, 2, tmW1,
                         tmW2, tmX, tmA, p));
  return 0;
}

}  // namespace cfm
