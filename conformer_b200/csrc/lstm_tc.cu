// LSTM recurrence of the CTC head on tcgen05 tensor cores (bf16 compute mode; decoder.py:16-21).
//
// The recurrence is sequential in time: per step  gates[4H, B] = W_hh[4H, H] . h_{t-1}[B, H]^T + xin_t  and then the cell
// update.  The CUDA-core kernel (decoder.cu, kept for the fp32 parity mode) spends 29 us per step at the head's size
// (B = 64, H = 640), cuDNN's LSTM 12-15 us, this kernel 5.5 us.  A CTA owns 16 hidden units = 64 gate rows and keeps its
// [64 x H] bf16 slice of W_hh in SHARED MEMORY for the whole sequence (80 KB at H = 640: loaded once by TMA), next to a
// buffer for the WHOLE h_{t-1} (80 KB at B = 64).  A step is
//   grid barrier (h_{t-1} of all CTAs is in L2; polled with relaxed loads, one acquire fence afterwards)
//   -> two bulk copies of h_{t-1} (the cells write h_t to global memory as the 128B-swizzled shared-memory image of the
//      MMA's B operand, so no tensor map is involved; the MMAs of the first half run under the second copy)
//   -> tcgen05.mma  D[gate rows, B] (TMEM, fp32) = W_hh tile . h_{t-1}^T,  K = H in 16-wide steps, N = B <= 64 (the
//      instruction is M = 128: accumulator lanes 64-127 read whatever follows the 64 weight rows and are ignored)
//   -> tcgen05.ld: thread = gate row, 64 batch columns; + the input projection xin_t (loaded at the top of the step,
//      in flight during the copies and MMAs; its columns are packed in the same row order, so a warp reads 128 contiguous
//      bytes per batch row)
//   -> the four gates of a unit sit in four lanes of one warp (row order below): a 4 x 4 transpose over those lanes (two
//      shuffle stages run over all cells at once, branch-free selects) hands every lane the i, f, g, o values of ITS cells
//      (unit, batch columns b = 4 j + gate index); cell state c and h stay in registers for the whole sequence
//   -> h_t (bf16) and the layer output (swish(h) for the last layer) are staged in shared memory as [batch row][16 units]
//      and leave as 16-byte stores: h_t into the ping-pong image the next step's copies read; one GPU-scope fence, arrive
//      on the grid barrier.
// Measured step (tools/lstm_trace.py): copies 1.8 us, MMAs done at 2.8, cells 0.8, publish 0.9.  Three things each cost
// more than the arithmetic before they were found: selects written as ?: on a lane-dependent index compile to divergent
// branches (a cell then took 0.67 us instead of 0.16); polling the barrier with ld.acquire invalidates the SM's L1 on
// every poll; loads the compiler is free to sink end up behind the accumulator wait.
// xin and the layer output are TIME-MAJOR ([T, B, .]) so that a step touches one contiguous region.
// Row order: unit u, gate g (nn.LSTM order i, f, g, o)  ->  packed row (u / 16) * 64 + ((u % 16) / 8) * 32 + g * 8 + u % 8,
// i.e. TMEM lane quadrant q (0 or 1) holds units 8q .. 8q+7 of the CTA, lane = g * 8 + u % 8 inside it.
// Packing semantics as in decoder.cu: a sample's state stops at its length, frames beyond it are zero.
// Batches above 64 run as independent slices of 64 (sequences do not interact).
#include <cooperative_groups.h>

#include "ops.h"

namespace cfm {
namespace {

constexpr int LT_ROWS = 64, LT_UNITS = 16, LT_KB = 64, LT_NMAX = 64;
constexpr int LT_THREADS = 4 * 32;                         // warps 0-1: cells; 2: TMA producer; 3: MMA issuer

struct LstmTcParams {
  const float* xin;        // TIME-MAJOR [T, Btot, 4H], columns in the packed row order, b_ih + b_hh included; this slice's b = 0
  const int64_t* lengths;  // [B] or null
  bf16* hbuf;              // [2][H/64][NPAD][64] ping-pong hidden state, stored as the 128B-swizzled shared-memory image of
                           // the MMA's B operand (zero-initialised; rows >= B stay zero)
  unsigned int* bar;       // grid barrier counter (zero-initialised)
  void* out;               // TIME-MAJOR [T, Btot, H]; this slice's b = 0
  int B, T, H, NPAD, last_layer;
  int Btot;                // batch rows per time step in xin / out (row stride between time steps)
  unsigned long long* trace;   // debug
};

// Polling with ld.acquire would invalidate the SM's L1 on every poll (CCTL.IVALL) and hold up the cell warps' stores -
// measured 0.67 us per cell iteration.  Poll relaxed, acquire once.
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long lt_time() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define LTRACE(slot) do { if (p.trace != nullptr && cta == 0 && t == 100 && lane == 0) p.trace[96 + (slot)] = lt_time(); } while (0)
// branch-free select (the lanes of a warp differ in the selector: written as ?: the compiler emits divergent branches, and
// a cell then costs 0.67 us instead of 0.05)
__device__ __forceinline__ float selp(bool c, float a, float b) {
  float r;
  asm("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\tselp.f32 %0, %1, %2, q;\n\t}\n" : "=f"(r) : "f"(a), "f"(b), "r"(static_cast<int>(c)));
  return r;
}

template <typename TOut>
__global__ void __launch_bounds__(LT_THREADS, 1)
lstm_tc_kernel(const __grid_constant__ CUtensorMap tmW, LstmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  const int KB = p.H / LT_KB;
  uint8_t* w_s = smem;                                        // KB blocks of [64 rows x 64 k] bf16 (8 KB each)
  uint8_t* h_s = w_s + KB * (LT_ROWS * 128);                  // KB blocks of [NPAD rows x 64 k] (<= 8 KB each): all of h_{t-1}
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(h_s + KB * (LT_NMAX * 128) + LT_ROWS * 128);   // (+ slack read by lanes 64-127)
  uint64_t* w_bar = full_bar + 16;
  uint64_t* d_full = w_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);
  bf16* st_h = reinterpret_cast<bf16*>(reinterpret_cast<uint8_t*>(full_bar) + 256);        // [64][16] staging of h_t
  TOut* st_o = reinterpret_cast<TOut*>(reinterpret_cast<uint8_t*>(st_h) + LT_NMAX * LT_UNITS * 2);   // [64][16] staging of the output

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int T = p.T, H = p.H, B = p.B, NPAD = p.NPAD;

  if (warp == 2 && lane == 0) {
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < KB; ++s) mbar_init(&full_bar[s], 1);
    mbar_init(w_bar, 1);
    mbar_init(d_full, 1);
    fence_mbar_init();
  }
  if (warp == 3) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 2) {
    // ------------------------------------------------------------------ producer: W_hh slice once, then h_{t-1} per step
    if (lane == 0) {
      mbar_expect_tx(w_bar, static_cast<uint32_t>(KB) * LT_ROWS * 128);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(w_s + kb * (LT_ROWS * 128), &tmW, w_bar, kb * LT_KB, cta * LT_ROWS);
      for (int t = 0; t < T; ++t) {
        if (t > 0) {                                          // every CTA has published h_{t-1} (and, this CTA's cells having
          const unsigned int target = static_cast<unsigned int>(t) * static_cast<unsigned int>(ncta);   // run, its MMAs of
          while (ld_relaxed_gpu(p.bar) < target) {            // step t-1 have finished reading h_s)
            __nanosleep(32);
          }
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          LTRACE(0);
          asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy stores of the peers -> TMA reads
        }
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.hbuf) + static_cast<size_t>(t & 1) * KB * NPAD * 128;
        // the whole image in two copies (k-blocks [0, KB/2) and [KB/2, KB)): the MMAs of the first half run under the
        // second copy
        const int k1 = (KB + 1) / 2;
        mbar_expect_tx(&full_bar[0], static_cast<uint32_t>(k1) * NPAD * 128);
        bulk_load(h_s, src, static_cast<uint32_t>(k1) * NPAD * 128, &full_bar[0]);
        if (KB > k1) {
          mbar_expect_tx(&full_bar[1], static_cast<uint32_t>(KB - k1) * NPAD * 128);
          bulk_load(h_s + static_cast<size_t>(k1) * NPAD * 128, src + static_cast<size_t>(k1) * NPAD * 128,
                    static_cast<uint32_t>(KB - k1) * NPAD * 128, &full_bar[1]);
        }
        LTRACE(1);
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(NPAD >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    mbar_wait(w_bar, 0);
    for (int t = 0; t < T; ++t) {
      const int k1 = (KB + 1) / 2;
      for (int kb = 0; kb < KB; ++kb) {
        if (kb == 0 || kb == k1) {
          mbar_wait(&full_bar[kb == 0 ? 0 : 1], t & 1);
          tc_fence_after();
        }
        if (kb == 0) LTRACE(2);
        if (kb == KB - 1) LTRACE(3);
        if (elect_one()) {
          const uint32_t sa = umma_desc_lo(smem_u32(w_s + kb * (LT_ROWS * 128)));
          const uint32_t sb = umma_desc_lo(smem_u32(h_s + kb * (NPAD * 128)));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_f16_lo(tmem_base, sa + k * 2, sb + k * 2, idesc, (kb | k) != 0 ? 1u : 0u);
          if (kb == KB - 1) tc_commit(d_full);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ cells: thread = packed gate row of this CTA
    const int q = warp, g_me = lane >> 3, u8 = lane & 7;
    const int unit = cta * LT_UNITS + q * 8 + u8;
    const int row = q * 32 + lane;                              // = TMEM lane
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);   // quadrants 0 and 1 hold the 64 real rows
    const int ncell = NPAD >> 2;                                // this lane's cells: batch columns b = 4 j + g_me
    float c_reg[LT_NMAX / 4], h_reg[LT_NMAX / 4];
    int len[LT_NMAX / 4];
#pragma unroll
    for (int j = 0; j < LT_NMAX / 4; ++j) {
      c_reg[j] = h_reg[j] = 0.f;
      const int b = 4 * j + g_me;
      len[j] = (j < ncell && b < B) ? (p.lengths != nullptr ? static_cast<int>(min(p.lengths[b], static_cast<int64_t>(T))) : T) : 0;
    }
    TOut* out = static_cast<TOut*>(p.out);
    const float* xcol = p.xin + static_cast<int64_t>(cta) * LT_ROWS + row;
    for (int t = 0; t < T; ++t) {
      // input projection of this step for this gate row, all batch columns (independent of the recurrence: in flight
      // while the MMAs run)
      // (volatile: left to itself the compiler sinks these loads to their uses after the wait and every cell then pays a
      // DRAM latency - measured 10 us per step)
      float xv[LT_NMAX];
#pragma unroll
      for (int b = 0; b < LT_NMAX; ++b) {
        xv[b] = 0.f;
        if (b < B) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(xv[b]) : "l"(xcol + (static_cast<int64_t>(t) * p.Btot + b) * 4 * H));
      }
      if (warp == 0) LTRACE(4);
      mbar_wait(d_full, t & 1);
      tc_fence_after();
      if (warp == 0) LTRACE(5);
      uint32_t r[LT_NMAX];
      {
        uint32_t lo[32], hi[32];
        tmem_ld_32x32(t_row, lo);
        if (NPAD > 32) tmem_ld_32x32(t_row + 32, hi);
        tmem_ld_wait();
#pragma unroll
        for (int b = 0; b < 32; ++b) {
          r[b] = lo[b];
          r[32 + b] = NPAD > 32 ? hi[b] : 0u;
        }
      }
      tc_fence_before();
      // ---- gates: 4 x 4 transpose over the four lanes that hold the gates of one unit (lane bits 3 and 4 = gate index);
      //      afterwards this lane holds i, f, g, o of batch column 4 j + g_me.  Stage by stage over all cells, so that the
      //      shuffles of different cells overlap.
      const bool b0 = (g_me & 1) != 0, b1 = (g_me & 2) != 0;
      float ga[LT_NMAX / 4][4];
#pragma unroll
      for (int j = 0; j < LT_NMAX / 4; ++j) {                     // stage A (partner: gate ^ 1): keep columns b0 and b0 + 2
        const float a0 = __uint_as_float(r[4 * j + 0]) + xv[4 * j + 0], a1 = __uint_as_float(r[4 * j + 1]) + xv[4 * j + 1];
        const float a2 = __uint_as_float(r[4 * j + 2]) + xv[4 * j + 2], a3 = __uint_as_float(r[4 * j + 3]) + xv[4 * j + 3];
        const float rA0 = __shfl_xor_sync(0xffffffffu, selp(b0, a0, a1), 8), rA1 = __shfl_xor_sync(0xffffffffu, selp(b0, a2, a3), 8);
        const float kA0 = selp(b0, a1, a0), kA1 = selp(b0, a3, a2);
        ga[j][0] = selp(b0, rA0, kA0); ga[j][1] = selp(b0, kA0, rA0);    // column b0: gates 2 b1, 2 b1 + 1
        ga[j][2] = selp(b0, rA1, kA1); ga[j][3] = selp(b0, kA1, rA1);    // column b0 + 2
      }
#pragma unroll
      for (int j = 0; j < LT_NMAX / 4; ++j) {                     // stage B (partner: gate ^ 2): keep column b0 + 2 b1
        const float lo_e = ga[j][0], lo_o = ga[j][1], hi_e = ga[j][2], hi_o = ga[j][3];
        const float rB_e = __shfl_xor_sync(0xffffffffu, selp(b1, lo_e, hi_e), 16), rB_o = __shfl_xor_sync(0xffffffffu, selp(b1, lo_o, hi_o), 16);
        const float kB_e = selp(b1, hi_e, lo_e), kB_o = selp(b1, hi_o, lo_o);
        ga[j][0] = selp(b1, rB_e, kB_e); ga[j][1] = selp(b1, rB_o, kB_o); ga[j][2] = selp(b1, kB_e, rB_e); ga[j][3] = selp(b1, kB_o, rB_o);
      }
      // ---- cells; h_t and the layer output are staged in shared memory as [batch row][16 units] and leave as 16-byte stores
#pragma unroll
      for (int j = 0; j < LT_NMAX / 4; ++j) {
        const int b = 4 * j + g_me;
        const bool live = t < len[j];
        float tg, tc;
        asm("tanh.approx.f32 %0, %1;" : "=f"(tg) : "f"(ga[j][2]));
        const float c = sigmoid_tanh(ga[j][1]) * c_reg[j] + sigmoid_tanh(ga[j][0]) * tg;
        asm("tanh.approx.f32 %0, %1;" : "=f"(tc) : "f"(c));
        const float hn = selp(live, sigmoid_tanh(ga[j][3]) * tc, h_reg[j]);
        c_reg[j] = selp(live, c, c_reg[j]);
        h_reg[j] = hn;
        st_h[b * LT_UNITS + q * 8 + u8] = __float2bfloat16_rn(hn);
        st_o[b * LT_UNITS + q * 8 + u8] = from_f32<TOut>(selp(live, p.last_layer ? swish_tanh(hn) : hn, 0.f));
      }
      if (warp == 0) LTRACE(6);
      named_bar_sync(1, 64);
      {
        // thread = batch row: its 16 units are two 16-byte chunks of the swizzled image (element (b, k) of k-block k / 64 at
        // b * 128 + ((k % 64 / 8) ^ (b & 7)) * 16 + k % 8 * 2) and 16 contiguous outputs
        const int bb = threadIdx.x;
        if (bb < B) {
          const int k0 = cta * LT_UNITS;
          uint8_t* img = reinterpret_cast<uint8_t*>(p.hbuf) + (static_cast<size_t>((t + 1) & 1) * KB + (k0 >> 6)) * NPAD * 128 + bb * 128;
          const uint4* src = reinterpret_cast<const uint4*>(st_h + bb * LT_UNITS);
          const int kc = (k0 & 63) >> 3;
          *reinterpret_cast<uint4*>(img + (((kc) ^ (bb & 7)) << 4)) = src[0];
          *reinterpret_cast<uint4*>(img + (((kc + 1) ^ (bb & 7)) << 4)) = src[1];
          const uint4* so = reinterpret_cast<const uint4*>(st_o + bb * LT_UNITS);
          uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(t) * p.Btot + bb) * H + k0);
#pragma unroll
          for (int i = 0; i < static_cast<int>(LT_UNITS * sizeof(TOut) / 16); ++i) dst[i] = so[i];
        }
      }
      named_bar_sync(1, 64);                                    // every thread's stores are ordered before thread 0 ...
      if (threadIdx.x == 0) {
        __threadfence();                                        // ... whose cumulative GPU-scope fence publishes them
        atomicAdd(p.bar, 1u);
      }
      if (warp == 0) LTRACE(7);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// packed[R(u, g)][:] = src[g * H + u][:],  R as in the header comment
template <typename TDst>
__global__ void lstm_pack_rows_kernel(const float* __restrict__ src, TDst* __restrict__ dst, int H, int K) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(4) * H * K) return;
  const int k = static_cast<int>(idx % K), R = static_cast<int>(idx / K);
  const int tile = R >> 6, l = R & 63, qq = l >> 5, g = (l >> 3) & 3, u = tile * LT_UNITS + qq * 8 + (l & 7);
  dst[idx] = from_f32<TDst>(src[(static_cast<int64_t>(g) * H + u) * K + k]);
}

// [B, T, d] fp32 -> time-major [T, B, d] bf16 (input of the first layer's projection)
__global__ void lstm_to_time_major_kernel(const float* __restrict__ x, bf16* __restrict__ y, int B, int T, int d) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i >= static_cast<int64_t>(B) * T * d) return;
  const int c = static_cast<int>(i % d);
  const int64_t r = i / d;                                      // output row t * B + b
  const int b = static_cast<int>(r % B), t = static_cast<int>(r / B);
  const float4 v = *reinterpret_cast<const float4*>(x + (static_cast<int64_t>(b) * T + t) * d + c);
  *reinterpret_cast<uint2*>(y + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}
// time-major logits [T, B, Vp] -> [B, T, V]
__global__ void lstm_logits_from_time_major_kernel(const float* __restrict__ lg, float* __restrict__ y, int B, int T, int V, int Vp) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(B) * T * V) return;
  const int v = static_cast<int>(i % V);
  const int64_t r = i / V;                                      // output row b * T + t
  const int t = static_cast<int>(r % T), b = static_cast<int>(r / T);
  y[i] = lg[(static_cast<int64_t>(t) * B + b) * Vp + v];
}

}  // namespace

int lstm_to_time_major(const float* x, bf16* y, int B, int T, int d, cudaStream_t stream) {
  CFM_REQUIRE(d % 4 == 0, "time-major cast needs a feature size that is a multiple of 4");
  const int64_t n = static_cast<int64_t>(B) * T * d / 4;
  lstm_to_time_major_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(x, y, B, T, d);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
int lstm_logits_from_time_major(const float* lg, float* y, int B, int T, int V, int Vp, cudaStream_t stream) {
  const int64_t n = static_cast<int64_t>(B) * T * V;
  lstm_logits_from_time_major_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(lg, y, B, T, V, Vp);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

extern unsigned long long* g_gemm_trace;

bool lstm_tc_supported(int H) {
  return H % 64 == 0 && H / 64 <= 16 && static_cast<size_t>(H / 64) * (LT_ROWS + LT_NMAX) * 128 + LT_ROWS * 128 + 2048 + LT_NMAX * LT_UNITS * 6 <= 227 * 1024;
}

int lstm_pack_rows_f32(const float* src, float* dst, int H, int K, cudaStream_t stream) {
  const int64_t n = static_cast<int64_t>(4) * H * K;
  lstm_pack_rows_kernel<float><<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(src, dst, H, K);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
int lstm_pack_rows_bf16(const float* src, bf16* dst, int H, int K, cudaStream_t stream) {
  const int64_t n = static_cast<int64_t>(4) * H * K;
  lstm_pack_rows_kernel<bf16><<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(src, dst, H, K);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

size_t lstm_tc_scratch_bytes(int H) { return static_cast<size_t>(2) * LT_NMAX * H * 2 + 256; }

template <typename TOut>
int lstm_layer_forward_tc(const float* xin_tm, const bf16* w_hh_packed, const int64_t* lengths, void* scratch, TOut* out_tm, int B, int T,
                          int H, bool last_layer, int num_sms, cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(B > 0 && T > 0 && lstm_tc_supported(H), "tensor-core LSTM: hidden size must be a multiple of 64 (up to 1024)");
  CFM_REQUIRE(H / LT_UNITS <= num_sms, "tensor-core LSTM keeps 16 hidden units per SM");
  const int KB = H / LT_KB;
  const int smem = 1024 + KB * (LT_ROWS + LT_NMAX) * 128 + LT_ROWS * 128 + 256 + LT_NMAX * LT_UNITS * 6;
  auto kern = lstm_tc_kernel<TOut>;
  static std::atomic<uint64_t> attr_done{0};
  CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(kern), smem));
  CUtensorMap tmW;
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(4) * H};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(H) * 2};
    cuuint32_t box[2] = {LT_KB, LT_ROWS};
    CFM_TRY(encode_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_hh_packed, dims, strides, box, "lstm W_hh"));
  }
  bf16* hbuf = static_cast<bf16*>(scratch);
  unsigned int* bar = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(scratch) + static_cast<size_t>(2) * LT_NMAX * H * 2);
  for (int b0 = 0; b0 < B; b0 += LT_NMAX) {                     // independent slices of up to 64 sequences
    const int nb = B - b0 < LT_NMAX ? B - b0 : LT_NMAX;
    const int npad = (nb + 15) & ~15;
    CFM_CUDA_OK(cudaMemsetAsync(scratch, 0, lstm_tc_scratch_bytes(H), stream));
    LstmTcParams p{xin_tm + static_cast<int64_t>(b0) * 4 * H, lengths != nullptr ? lengths + b0 : nullptr, hbuf, bar,
                   out_tm + static_cast<int64_t>(b0) * H, nb, T, H, npad, last_layer ? 1 : 0, B, g_gemm_trace};
    void* args[] = {const_cast<CUtensorMap*>(&tmW), &p};
    CFM_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(H / LT_UNITS), dim3(LT_THREADS), args, smem, stream));
  }
  return 0;
}
template int lstm_layer_forward_tc<bf16>(const float*, const bf16*, const int64_t*, void*, bf16*, int, int, int, bool, int, cudaStream_t);
template int lstm_layer_forward_tc<float>(const float*, const bf16*, const int64_t*, void*, float*, int, int, int, bool, int, cudaStream_t);

}  // namespace cfm
