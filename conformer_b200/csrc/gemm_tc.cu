// bf16 GEMM on the 5th-generation tensor cores: C[M,N] = epi(A[M,K] * W[N,K]^T), fp32 accumulate.
//
//   * operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a multi-stage shared-memory
//     ring guarded by full/empty mbarriers;
//   * tcgen05.mma (kind::f16, M=128, N=BLOCK_N, K=16) issued by one thread, accumulators in TMEM,
//     double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * 8 epilogue warps read the accumulator with tcgen05.ld (one thread = one output row) and apply
//     bias / Swish / ReLU / (residual + scale) before storing;
//   * persistent: grid = min(#tiles, #SMs), static round-robin tile schedule, N-tiles innermost so the
//     CTAs running together share the same rows of A in L2.
//
// CONV mode turns the second front-end convolution (Conv2d(d,d,3,stride 2), convolution.py:39,46)
// into an implicit GEMM: the A tile for tap (kh,kw) is a strided box of conv1's channels-last output
// fetched by a 5-D tensor map (the stride-2 time axis is split into (pair, parity)), so no im2col
// buffer exists.
#include <mutex>

#include "ops.h"

namespace cfm {
namespace {

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;

template <int BLOCK_N> struct Cfg {
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;                 // two accumulator stages
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

struct GemmParams {
  int M, N, K;
  int num_m_tiles, num_n_tiles, num_kb;
  const float* bias;
  const float* res;
  float res_scale;
  void* C;
  int ldc;
  // conv mode
  int T2, F2, TCH, kb_per_tap;
};

template <int BLOCK_N, int EPI, bool CONV>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  using C_ = Cfg<BLOCK_N>;
  constexpr int STAGES = C_::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * C_::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, C_::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
      int cb = 0, cf = 0, ct = 0;
      if (CONV) {
        ct = m_tile % p.TCH;
        cf = (m_tile / p.TCH) % p.F2;
        cb = m_tile / (p.TCH * p.F2);
      }
      for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (lane == 0) {
          uint8_t* sa = smem + s * C_::STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], C_::STAGE_BYTES);
          if (CONV) {
            const int tap = kb / p.kb_per_tap, c0 = (kb - tap * p.kb_per_tap) * BLOCK_K;
            const int kw = tap / 3, kh = tap - kw * 3;
            // dims: (c, f1, parity, pair, b); source time index 2*t2 + kw = 2*(t2 + (kw>>1)) + (kw&1)
            tma_load_5d(sa, &tmA, &full_bar[s], c0, 2 * cf + kh, kw & 1, ct * BLOCK_M + (kw >> 1), cb);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[s], kb * BLOCK_K, m_tile * BLOCK_M);
          }
          tma_load_2d(sb, &tmB, &full_bar[s], kb * BLOCK_K, n_tile * BLOCK_N);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
    uint32_t it = 0, acc_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_it) {
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      mbar_wait(&tempty_bar[as], aph ^ 1);                     // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + s * C_::STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = umma_desc_k_sw128(sa + k * UMMA_K * 2);
            const uint64_t db = umma_desc_k_sw128(sb + k * UMMA_K * 2);
            tc_mma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[s]);                            // smem slot free once these MMAs retire
          if (kb == p.num_kb - 1) tc_commit(&tfull_bar[as]);   // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int q = warp & 3;                                    // TMEM lane quadrant this warp may read
    const int col_half = ew >> 2;
    const int r_in_tile = q * 32 + lane;
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_it) {
      const int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
      const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
      int64_t out_row;
      bool row_ok;
      if (CONV) {
        const int ct = m_tile % p.TCH, cf = (m_tile / p.TCH) % p.F2, cb = m_tile / (p.TCH * p.F2);
        const int t2 = ct * BLOCK_M + r_in_tile;
        row_ok = t2 < p.T2;
        out_row = (static_cast<int64_t>(cb) * p.T2 + t2) * p.F2 + cf;
      } else {
        out_row = static_cast<int64_t>(m_tile) * BLOCK_M + r_in_tile;
        row_ok = out_row < p.M;
      }
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N;
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / 64; ++ch) {
        const int col_in_tile = col_half * (BLOCK_N / 2) + ch * 32;
        const int col0 = n_tile * BLOCK_N + col_in_tile;
        uint32_t r[32];
        tmem_ld_32x32(t_row + col_in_tile, r);
        tmem_ld_wait();
        if (row_ok && col0 < p.N) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {                         // groups of 8 columns
            const int col = col0 + g * 8;
            if (col + 8 <= p.N) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
              if (p.bias != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (EPI == EPI_BIAS_SWISH) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = swish_fast(v[i]);
              }
              if (EPI == EPI_BIAS_RELU) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
              }
              if (EPI == EPI_BIAS_F32) {
                float* cr = reinterpret_cast<float*>(p.C) + out_row * p.ldc + col;
                *reinterpret_cast<float4*>(cr) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(cr + 4) = make_float4(v[4], v[5], v[6], v[7]);
              } else if (EPI == EPI_BIAS_RES) {
                const float* rr = p.res + out_row * p.ldc + col;
                const float4 r0 = *reinterpret_cast<const float4*>(rr);
                const float4 r1 = *reinterpret_cast<const float4*>(rr + 4);
                float4 o0, o1;
                o0.x = fmaf(p.res_scale, v[0], r0.x); o0.y = fmaf(p.res_scale, v[1], r0.y);
                o0.z = fmaf(p.res_scale, v[2], r0.z); o0.w = fmaf(p.res_scale, v[3], r0.w);
                o1.x = fmaf(p.res_scale, v[4], r1.x); o1.y = fmaf(p.res_scale, v[5], r1.y);
                o1.z = fmaf(p.res_scale, v[6], r1.z); o1.w = fmaf(p.res_scale, v[7], r1.w);
                float* cr = reinterpret_cast<float*>(p.C) + out_row * p.ldc + col;
                *reinterpret_cast<float4*>(cr) = o0;
                *reinterpret_cast<float4*>(cr + 4) = o1;
              } else {
                uint4 o;
                o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                bf16* cr = reinterpret_cast<bf16*>(p.C) + out_row * p.ldc + col;
                *reinterpret_cast<uint4*>(cr) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C_::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_once;
int g_init_status = -1;

int encode_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems, uint32_t box_inner,
              uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 3;
  }
  return 0;
}

int encode_conv_5d(CUtensorMap* tm, const void* base, const ConvGeom& g) {
  // conv1 output [B][T1p][F1][C] viewed as (c, f1, parity, pair, b); T1p is even.
  const uint64_t C = g.C, F1 = g.F1;
  cuuint64_t dims[5] = {C, F1, 2, static_cast<cuuint64_t>(g.T1p / 2), static_cast<cuuint64_t>(g.B)};
  cuuint64_t strides[4] = {C * 2, F1 * C * 2, 2 * F1 * C * 2, static_cast<cuuint64_t>(g.T1p) * F1 * C * 2};
  cuuint32_t box[5] = {BLOCK_K, 1, 1, BLOCK_M, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(5d) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 3;
  }
  return 0;
}

template <int BLOCK_N, int EPI, bool CONV>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int num_sms, cudaStream_t stream) {
  auto kern = gemm_tc_kernel<BLOCK_N, EPI, CONV>;
  static bool attr_done = false;
  if (!attr_done) {
    CFM_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BLOCK_N>::SMEM_BYTES));
    attr_done = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, NUM_THREADS, Cfg<BLOCK_N>::SMEM_BYTES, stream>>>(tmA, tmB, p);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int BLOCK_N, bool CONV>
int launch_epi(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, Epilogue epi, int num_sms,
               cudaStream_t stream) {
  switch (epi) {
    case EPI_BIAS: return launch<BLOCK_N, EPI_BIAS, CONV>(tmA, tmB, p, num_sms, stream);
    case EPI_BIAS_SWISH: return launch<BLOCK_N, EPI_BIAS_SWISH, CONV>(tmA, tmB, p, num_sms, stream);
    case EPI_BIAS_RELU: return launch<BLOCK_N, EPI_BIAS_RELU, CONV>(tmA, tmB, p, num_sms, stream);
    case EPI_BIAS_RES: return launch<BLOCK_N, EPI_BIAS_RES, CONV>(tmA, tmB, p, num_sms, stream);
    case EPI_BIAS_F32: return launch<BLOCK_N, EPI_BIAS_F32, CONV>(tmA, tmB, p, num_sms, stream);
  }
  set_error("unknown epilogue");
  return 1;
}

}  // namespace

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

int gemm_bf16_tc(const bf16* A, int lda, const bf16* W, const float* bias, void* C, int ldc, int M, int N, int K,
                 Epilogue epi, const float* res, float res_scale, const ConvGeom* conv, int num_sms,
                 cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM");
  CFM_REQUIRE(K % 8 == 0 && N % 8 == 0 && ldc % 8 == 0, "tensor-core GEMM needs K, N, ldc multiples of 8");
  CFM_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(C) & 15) == 0,
              "operands must be 16-byte aligned");
  CFM_REQUIRE(epi != EPI_BIAS_RES || res != nullptr, "residual epilogue without residual");
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.res = res; p.res_scale = res_scale; p.C = C; p.ldc = ldc;
  p.num_kb = static_cast<int>(ceil_div(K, BLOCK_K));
  CUtensorMap tmA, tmB;
  if (conv != nullptr) {
    CFM_REQUIRE(conv->C % BLOCK_K == 0, "implicit-GEMM conv needs channels to be a multiple of 64");
    CFM_REQUIRE(conv->T1p % 2 == 0 && K == 9 * conv->C, "bad conv geometry");
    CFM_REQUIRE(M == conv->B * conv->T2 * conv->F2, "conv GEMM row count mismatch");
    p.T2 = conv->T2; p.F2 = conv->F2;
    p.TCH = static_cast<int>(ceil_div(conv->T2, BLOCK_M));
    p.kb_per_tap = conv->C / BLOCK_K;
    p.num_m_tiles = conv->B * conv->F2 * p.TCH;
    CFM_TRY(encode_conv_5d(&tmA, A, *conv));
  } else {
    CFM_REQUIRE(lda % 8 == 0, "lda must be a multiple of 8");
    p.num_m_tiles = static_cast<int>(ceil_div(M, BLOCK_M));
    CFM_TRY(encode_2d(&tmA, A, static_cast<uint64_t>(K), static_cast<uint64_t>(M), static_cast<uint64_t>(lda), BLOCK_K,
                      BLOCK_M));
  }
  // tile width: 256 halves the A re-reads, 128 quantises better over the SMs; pick the cheaper schedule
  auto cost = [&](int bn) {
    const int64_t tiles = static_cast<int64_t>(p.num_m_tiles) * ceil_div(N, bn);
    return ceil_div(tiles, num_sms) * bn;
  };
  const bool use256 = (N > 128) && (cost(256) <= cost(128));
  const int bn = use256 ? 256 : 128;
  p.num_n_tiles = static_cast<int>(ceil_div(N, bn));
  CFM_TRY(encode_2d(&tmB, W, static_cast<uint64_t>(K), static_cast<uint64_t>(N), static_cast<uint64_t>(K), BLOCK_K,
                    static_cast<uint32_t>(bn)));
  if (conv != nullptr)
    return use256 ? launch_epi<256, true>(tmA, tmB, p, epi, num_sms, stream)
                  : launch_epi<128, true>(tmA, tmB, p, epi, num_sms, stream);
  return use256 ? launch_epi<256, false>(tmA, tmB, p, epi, num_sms, stream)
                : launch_epi<128, false>(tmA, tmB, p, epi, num_sms, stream);
}

}  // namespace cfm
