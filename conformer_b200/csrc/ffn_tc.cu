// Fused FeedForwardModule half-step for d_model = 256 on tcgen05 tensor cores:
//
//     x[M,256] += scale * ( swish( A W1^T + b1 ) W2^T + b2 ),  A = LN(x) as bf16      (ffn.py:15-23 + block.py:19,25)
//
// One CTA owns 128 rows.  Nothing but A, the weights and the updated x touches global memory:
//   prologue   A arrives already normalised (the kernel that produced x also emitted LN(x) as bf16: the block-final
//              dual LayerNorm, the fused conv module's epilogue, or a LayerNorm launch) and is TMA-loaded straight into
//              the 128B-swizzled K-major layout the MMA descriptors expect.  The kernel never READS x, so x may be
//              updated in place by any number of CTAs sharing a row tile;
//   main loop  the 1024 hidden units are processed in 8 chunks of 128:
//                GEMM1(c): acc1[c&1] (TMEM, 128 cols)  = LN(x) . W1[c]^T          (K = 256)
//                epilogue: hidden = bf16(swish(acc1 + b1)) written IN PLACE over acc1[c&1] in tensor memory
//                GEMM2(c): acc2 (TMEM, 256 cols)      += hidden . W2[:, c]^T      (K = 128, A operand from TMEM)
//              software-pipelined (GEMM1(c+1) is issued before GEMM2(c)) so the tensor pipe keeps working
//              while the 16 epilogue warps run the Swish; W1/W2 tiles stream through a TMA/mbarrier ring.
//              The loop is bound by the shared-memory port (128 B/clk): per chunk the MMAs read 128 KB (GEMM1, both
//              operands) + 64 KB (GEMM2, W2 only) and TMA writes 128 KB of weights = 2560 clk against 2048 clk of
//              tensor time; with the hidden tile in shared memory it was 3072 clk (measured 1.56 us per chunk);
//   epilogue   scale*(acc2 + b2) is staged (aliasing the now-dead operand buffers) and leaves through a TMA
//              reduction store, i.e. the residual add is executed by the L2 (x is never re-read).
// TMEM: 2 x 128 (acc1 double buffer) + 256 (acc2) = 512 columns.
#include "ops.h"

namespace cfm {
namespace {

constexpr int D = 256, HID = 4 * D, BM = 128, HC = 128, NCHUNK = HID / HC;   // NCHUNK = 8
constexpr int KB = 64;                                     // K elements per 128-byte swizzled row
constexpr int XS_BYTES = (D / KB) * BM * 128;              // 64 KB: LN(x) as 4 k-blocks
constexpr int HS_BYTES = (HC / KB) * BM * 128;             // 32 KB per hidden buffer
constexpr int STAGE_BYTES = 32 * 1024, STAGES = 3;
constexpr int NUM_EPI_WARPS = 16, NUM_EPI_THREADS = NUM_EPI_WARPS * 32, NUM_THREADS = 128 + NUM_EPI_THREADS;
constexpr int SMEM_BYTES = 1024 + XS_BYTES + 2 * HS_BYTES + STAGES * STAGE_BYTES + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "fused FFN shared memory budget");
constexpr int OUT_BLK_BYTES = BM * 128;                    // staging block: 128 rows x 32 fp32
static_assert((D / 32) * OUT_BLK_BYTES <= XS_BYTES + 2 * HS_BYTES, "output staging must fit the operand buffers");

struct FfnParams {
  const float* b1;
  const float* b2;
  float scale;
  int M;
  int num_tiles;
  unsigned long long* trace;     // debug: wait-time totals of CTA 0's MMA warp (ns)
};
__device__ __forceinline__ unsigned long long ffn_time_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Work items of one CTA.  With T row tiles on G CTAs, every CTA first takes T / G whole tiles; each of the T % G
// left-over tiles is then shared by G / (T % G) CTAs, each taking a contiguous range of its 8 HIDDEN chunks, so the
// last wave is as full as the first (187 tiles on 148 SMs: 39 tiles shared by 3-4 CTAs each).  Partial items add
// their share of the output with the same TMA reduction store (the L2 sums them); b2 is added by the item that
// owns chunk 0.
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
    // left-over tile lt is shared by the CTAs [lt*G/L, (lt+1)*G/L): one partial item per CTA, no straddling
    const int lt = static_cast<int>(static_cast<long long>(c) * L / G);
    const int first = (lt * G + L - 1) / L, last = ((lt + 1) * G + L - 1) / L;     // CTAs of this tile: [first, last)
    const int k = last - first, j = c - first;
    const int cb = j * NCHUNK / k, ce = (j + 1) * NCHUNK / k;
    if (cb >= ce) return false;
    it = Item{R * G + lt, cb, ce};
    return true;
  }
};

// The A operand LN(x) arrives already normalised (bf16 [M,256], tensor map tmA) and is TMA-loaded into the MMA layout.
__global__ void __launch_bounds__(NUM_THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmA, FfnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* xs = smem;
  uint8_t* hs = xs + XS_BYTES;                               // (with xs: staging of the output tile only; the hidden tile lives in TMEM)
  uint8_t* ring = hs + 2 * HS_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* xs_full = empty_bar + STAGES;
  uint64_t* acc1_full = xs_full + 1;      // [2]
  uint64_t* hs_full = acc1_full + 2;      // [2]
  uint64_t* acc2_full = hs_full + 2;
  uint64_t* acc2_empty = acc2_full + 1;   // the epilogue warps have read acc2 of the previous item out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

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
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc1_full[s], 1);
      mbar_init(&hs_full[s], NUM_EPI_WARPS);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, NUM_EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_whole_sm(tmem_slot);
  const uint32_t tm_acc2 = tmem_base + 2 * HC;

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer
    // consumption order per item: G1(cb) G1(cb+1) G2(cb) G1(cb+2) G2(cb+1) ... G2(ce-1); each GEMM uses two ring stages
    uint32_t it = 0;
    auto load_g1 = [&](int c) {
      for (int s2 = 0; s2 < 2; ++s2, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        if (lane == 0) {
          uint8_t* dst = ring + s * STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_2d(dst, &tmW1, &full_bar[s], (2 * s2) * KB, c * HC);               // W1[c*128.., k-block 2*s2]
          tma_load_2d(dst + 16384, &tmW1, &full_bar[s], (2 * s2 + 1) * KB, c * HC);   // k-block 2*s2+1
        }
        __syncwarp();
      }
    };
    auto load_g2 = [&](int c) {
      for (int s2 = 0; s2 < 2; ++s2, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_2d(ring + s * STAGE_BYTES, &tmW2, &full_bar[s], c * HC + s2 * KB, 0);  // W2[:, c*128 + s2*64 ..]
        }
        __syncwarp();
      }
    };
    ItemIter items(blockIdx.x, gridDim.x, p.num_tiles);
    Item w;
    while (items.next(w)) {
      load_g1(w.cb);
      for (int c = w.cb; c < w.ce; ++c) {
        if (c + 1 < w.ce) load_g1(c + 1);
        load_g2(c);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc1 = umma_idesc_bf16(BM, HC);     // 128 x 128
    constexpr uint32_t idesc2 = umma_idesc_bf16(BM, D);      // 128 x 256
    uint32_t it = 0, gc1 = 0, gc2 = 0;                        // ring slot counter, chunks issued to GEMM1 / GEMM2 so far
    unsigned long long w_acc1 = 0, w_full = 0, w_hs = 0, w_xs = 0, t_begin = ffn_time_ns();
    const bool tr = p.trace != nullptr && blockIdx.x == 0;
    auto gemm1 = [&]() {
      // acc1[b] last held the hidden tile of the chunk two back: its GEMM2 was issued before this GEMM1 and the tensor
      // pipe executes in issue order, and that GEMM2 was itself issued after the Swish warps had finished with the
      // buffer (hs_full) - no further handshake is needed
      const int b = gc1 & 1;
      unsigned long long t0 = 0;
      for (int s2 = 0; s2 < 2; ++s2, ++it) {
        const int s = it % STAGES;
        t0 = tr ? ffn_time_ns() : 0;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        if (tr) w_full += ffn_time_ns() - t0;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sb = umma_desc_lo(smem_u32(ring + s * STAGE_BYTES));
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t sa = umma_desc_lo(smem_u32(xs + (2 * s2 + kk) * (BM * 128)));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_f16_lo(tmem_base + b * HC, sa + k * 2, sb + kk * (16384 >> 4) + k * 2, idesc1, (s2 | kk | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[s]);
          if (s2 == 1) tc_commit(&acc1_full[b]);
        }
        __syncwarp();
      }
      ++gc1;
    };
    auto gemm2 = [&](bool first, bool last, uint32_t n) {
      const int b = gc2 & 1;
      if (first && n > 0) {                                 // the previous item's output has left the accumulator
        mbar_wait(acc2_empty, (n - 1) & 1);
        tc_fence_after();
      }
      unsigned long long t0 = tr ? ffn_time_ns() : 0;
      mbar_wait(&hs_full[b], (gc2 >> 1) & 1);
      if (tr) w_hs += ffn_time_ns() - t0;
      tc_fence_after();
      for (int s2 = 0; s2 < 2; ++s2, ++it) {
        const int s = it % STAGES;
        t0 = tr ? ffn_time_ns() : 0;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        if (tr) w_full += ffn_time_ns() - t0;
        tc_fence_after();
        if (elect_one()) {
          // hidden units 16*kg .. +15 of the chunk (kg = 4*s2 + k): bf16 pairs in columns 32*(kg/2) + 8*(kg%2) .. +7 of acc1[b]
          const uint32_t ta = tmem_base + b * HC + s2 * 64;
          const uint32_t sb = umma_desc_lo(smem_u32(ring + s * STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16_ts(tm_acc2, ta + 32 * (k >> 1) + 8 * (k & 1), sb + k * 2, idesc2, (!first || (s2 | k) != 0) ? 1u : 0u);
          tc_commit(&empty_bar[s]);
          if (s2 == 1 && last) tc_commit(acc2_full);
        }
        __syncwarp();
      }
      ++gc2;
    };
    ItemIter items(blockIdx.x, gridDim.x, p.num_tiles);
    Item w;
    for (uint32_t n = 0; items.next(w); ++n) {
      unsigned long long t0 = tr ? ffn_time_ns() : 0;
      mbar_wait(xs_full, n & 1);                               // LN(x) of this item staged (and acc2 of the last one drained)
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
      p.trace[4] = w_acc1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ A-operand requests + epilogues (16 warps)
    const int ew = warp - 4;
    const int q = warp & 3, cgrp = ew >> 2;
    const int tid_e = threadIdx.x - 128;
    ItemIter items(blockIdx.x, gridDim.x, p.num_tiles);
    Item w;
    uint32_t gc = 0;                                            // chunks processed so far (acc1 / hs double buffers)
    const int row = q * 32 + lane;
    const int sw = row & 7;
    const float4* b1v = reinterpret_cast<const float4*>(p.b1 + cgrp * 32);
    // A operand of an item = four [128 rows x 64] k-blocks straight into the swizzled MMA layout (rows past M arrive
    // as zeros).  It is requested as soon as the previous item's last GEMM1 has finished reading xs, i.e. a chunk and a
    // whole output epilogue ahead of its first use: the next item's first GEMM1s then run under this item's epilogue.
    auto request_xs = [&](int tile) {
      mbar_expect_tx(xs_full, XS_BYTES);
#pragma unroll
      for (int kb = 0; kb < D / KB; ++kb) tma_load_2d(xs + kb * (BM * 128), &tmA, xs_full, kb * KB, tile * BM);
    };
    Item w_next;
    bool have = items.next(w);
    if (have && tid_e == 0) request_xs(w.tile);
    for (uint32_t n = 0; have; ++n) {
    const bool have_next = items.next(w_next);
    const int m0 = w.tile * BM;
    // ---- per-chunk epilogue: hidden = bf16(swish(acc1 + b1)) into the MMA A layout
    for (int c = w.cb; c < w.ce; ++c, ++gc) {
      const int b = gc & 1;
      float4 bc[8];                                            // issued before the wait: L2 latency overlaps it
#pragma unroll
      for (int g = 0; g < 8; ++g) bc[g] = __ldg(b1v + c * (HC / 4) + g);
      mbar_wait(&acc1_full[b], (gc >> 1) & 1);
      tc_fence_after();
      if (c == w.ce - 1 && have_next && tid_e == 0) request_xs(w_next.tile);   // every GEMM1 of this item is done
      uint32_t r[32];
      const uint32_t t_acc1 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * HC + cgrp * 32;
      tmem_ld_32x32(t_acc1, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 b4 = bc[g];
        v[g * 4 + 0] = swish_tanh(__uint_as_float(r[g * 4 + 0]) + b4.x);
        v[g * 4 + 1] = swish_tanh(__uint_as_float(r[g * 4 + 1]) + b4.y);
        v[g * 4 + 2] = swish_tanh(__uint_as_float(r[g * 4 + 2]) + b4.z);
        v[g * 4 + 3] = swish_tanh(__uint_as_float(r[g * 4 + 3]) + b4.w);
      }
      // hidden units cgrp*32 .. +31 of the chunk, as bf16 pairs, over the first 16 of the 32 accumulator columns this
      // thread has just read (only this thread touches them): the TMEM-resident A operand of GEMM2
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) pk[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
      tmem_st_32x16(t_acc1, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hs_full[b]);
    }
    // ---- final epilogue: x += scale * (acc2 + b2) through a TMA reduction store
    mbar_wait(acc2_full, n & 1);
    tc_fence_after();
    {
      // xs already belongs to the next item: the tile is staged in two halves of 128 columns through the 64 KB that used
      // to hold the hidden tile, and acc2 is handed back to the MMA warp as soon as the second half is in registers
      uint8_t* stage_out = hs;
      if (tid_e == 0) bulk_wait_read0();                         // second half of the previous item has left the staging
      named_bar_sync(1, NUM_EPI_THREADS);
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int col = hh * 128 + cgrp * 32;
        uint32_t r[32];
        tmem_ld_32x32(tm_acc2 + (static_cast<uint32_t>(q * 32) << 16) + col, r);
        tmem_ld_wait();
        if (hh == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc2_empty);
        }
        uint8_t* blk = stage_out + cgrp * OUT_BLK_BYTES + row * 128;
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
          for (int blk4 = 0; blk4 < 4; ++blk4)
            tma_reduce_add_2d(&tmX, stage_out + blk4 * OUT_BLK_BYTES, hh * 128 + blk4 * 32, m0);
          bulk_commit();
          if (hh == 0) bulk_wait_read0();                        // the first half has left the staging
        }
        if (hh == 0) named_bar_sync(1, NUM_EPI_THREADS);
      }
    }
    w = w_next;
    have = have_next;
    }
    if (tid_e == 0) bulk_wait_read0();                   // the last half tile has left shared memory
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

bool ffn_fused_supported(int d) { return d == D; }

int ffn_fused_bf16(const bf16* a_pre, float* x, const bf16* W1, const float* b1, const bf16* W2, const float* b2, float scale,
                   int M, int d, int num_sms, cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(d == D, "fused FFN kernel is specialised for d_model = 256");
  CFM_REQUIRE(M > 0, "empty FFN");
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
    cuuint32_t box[2] = {KB, HC};
    CFM_TRY(encode_tmap(&tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W1, dims, strides, box, "ffn W1"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(HID), static_cast<cuuint64_t>(D)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(HID) * 2};
    cuuint32_t box[2] = {KB, D};
    CFM_TRY(encode_tmap(&tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W2, dims, strides, box, "ffn W2"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * 4};
    cuuint32_t box[2] = {32, BM};
    CFM_TRY(encode_tmap(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, "ffn x"));
  }
  static std::atomic<uint64_t> attr_done{0};
  CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(ffn_fused_kernel), SMEM_BYTES));
  const int tiles = static_cast<int>(ceil_div(M, BM));
  FfnParams p{b1, b2, scale, M, tiles, g_gemm_trace};
  // Persistent CTAs, left-over row tiles shared along the hidden dimension (187 tiles on 148 SMs: the last wave is as full
  // as the first); a share's prologue is one 64 KB TMA load.  Sharing is safe because the kernel only ever ADDS to x.
  const int grid = tiles < num_sms ? tiles : num_sms;
  ffn_fused_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmW1, tmW2, tmX, tmA, p);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
