// First subsampling convolution Conv2d(1, C, 3, stride 2) + ReLU (convolution.py:37,44-45) for C = 256 on the
// tensor cores, written channels-last as [B][T1p][F1][C] bf16 (the layout the conv2 implicit GEMM consumes).
//
// The op writes 958 MB at the C2 shape and is therefore HBM-write bound; on CUDA cores it was bound by the fp32 FMA pipe
// instead (9 MACs per output, 4.3 G MACs).  Here it is a skinny GEMM: row m = (b, t1, f1), K = the 3x3 patch (9 values,
// zero-padded to one K=16 MMA step), N = 256 channels.  Persistent CTAs, one per SM:
//   warps 0,3   build the A tile [128 rows x 16] from the fp32 mel input (bf16, 128B-swizzled rows; the unused K columns
//               of the stage were zeroed once);
//   warp 1      one tcgen05.mma (128 x 256 x 16) per tile into a double-buffered TMEM accumulator;
//   warps 4-19  epilogue: + bias, ReLU, bf16, swizzled staging (double buffered) and TMA store of the 64 KB tile.
#include "ops.h"

namespace cfm {
namespace {

constexpr int C = 256, BM = 128;
constexpr int W_BYTES = C * 128;                    // [256 x 64] bf16, only k < 9 non-zero
constexpr int A_BYTES = BM * 128, A_STAGES = 3;
constexpr int CBLK_BYTES = BM * 128;                // staging block: 128 rows x 64 bf16
constexpr int CBUF_BYTES = (C / 64) * CBLK_BYTES;   // 64 KB
constexpr int NUM_PROD_WARPS = 2, NUM_EPI_WARPS = 16, NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = 128 + NUM_EPI_THREADS;
constexpr int SMEM_BYTES = 1024 + W_BYTES + A_STAGES * A_BYTES + 2 * CBUF_BYTES + C * 4 + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "conv1 shared memory budget");

struct Conv1Params {
  const float* mel;     // [B][n_mel][Tin]
  const float* bias;    // [C]
  int n_mel, Tin, T1, T1p, F1;
  int64_t M;            // B * T1p * F1
  int num_tiles;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv1_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut, Conv1Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align_1024(smem_raw);
  uint8_t* w_s = smem;
  uint8_t* a_s = w_s + W_BYTES;
  uint8_t* cbuf = a_s + A_STAGES * A_BYTES;
  float* bias_s = reinterpret_cast<float*>(cbuf + 2 * CBUF_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_s + C);
  uint64_t* empty_bar = full_bar + A_STAGES;
  uint64_t* tfull_bar = empty_bar + A_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* w_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(&full_bar[s], NUM_PROD_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], NUM_EPI_WARPS);
    }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  // zero the A stages once: the producers only ever rewrite the two 16-byte chunks that hold k = 0..15
  for (int i = threadIdx.x; i < A_STAGES * A_BYTES / 16; i += NUM_THREADS)
    reinterpret_cast<uint4*>(a_s)[i] = make_uint4(0u, 0u, 0u, 0u);
  const int slab = blockIdx.y;                                  // 256 output channels per slab (Conformer-L: two)
  if (threadIdx.x < C) bias_s[threadIdx.x] = p.bias[slab * C + threadIdx.x];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_whole_sm(tmem_slot);
  pdl_wait();

  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------------ A producers: 64 threads, 2 rows each per tile
    const int pt = (warp == 0 ? 0 : 32) + lane;
    if (pt == 0) {
      mbar_expect_tx(w_bar, W_BYTES);
      tma_load_2d(w_s, &tmW, w_bar, 0, slab * C);
    }
    const int64_t rows_per_b = static_cast<int64_t>(p.T1p) * p.F1;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int s = it % A_STAGES;
      float v[2][9];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = pt + rr * 64;
        const int64_t m = static_cast<int64_t>(tile) * BM + r;
#pragma unroll
        for (int i = 0; i < 9; ++i) v[rr][i] = 0.f;
        if (m < p.M) {
          const int b = static_cast<int>(m / rows_per_b);
          const int rem = static_cast<int>(m - b * rows_per_b);
          const int t1 = rem / p.F1, f1 = rem - t1 * p.F1;
          if (t1 < p.T1) {
            const float* src = p.mel + (static_cast<int64_t>(b) * p.n_mel + 2 * f1) * p.Tin + 2 * t1;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) v[rr][kh * 3 + kw] = __ldg(src + static_cast<int64_t>(kh) * p.Tin + kw);
          }
        }
      }
      mbar_wait(&empty_bar[s], ((it / A_STAGES) & 1) ^ 1);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = pt + rr * 64;
        uint8_t* row = a_s + s * A_BYTES + r * 128;
        const uint4 c0 = make_uint4(pack_bf16x2(v[rr][0], v[rr][1]), pack_bf16x2(v[rr][2], v[rr][3]),
                                    pack_bf16x2(v[rr][4], v[rr][5]), pack_bf16x2(v[rr][6], v[rr][7]));
        const uint4 c1 = make_uint4(pack_bf16x2(v[rr][8], 0.f), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(row + ((0 ^ (r & 7)) << 4)) = c0;
        *reinterpret_cast<uint4*>(row + ((1 ^ (r & 7)) << 4)) = c1;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: one K=16 step per tile
    constexpr uint32_t idesc = umma_idesc_bf16(BM, C);
    mbar_wait(w_bar, 0);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int s = it % A_STAGES;
      const uint32_t as = it & 1;
      mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
      mbar_wait(&full_bar[s], (it / A_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        tc_mma_f16_lo(tmem_base + as * C, umma_desc_lo(smem_u32(a_s + s * A_BYTES)), umma_desc_lo(smem_u32(w_s)), idesc, 0u);
        tc_commit(&empty_bar[s]);
        tc_commit(&tfull_bar[as]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: 16 warps, thread = row x 64 channels
    const int q = warp & 3, cgrp = (warp - 4) >> 2;
    const int row = q * 32 + lane, sw = row & 7;
    const int tid_e = threadIdx.x - 128;
    const float4* bsm = reinterpret_cast<const float4*>(bias_s + cgrp * 64);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1;
      uint8_t* cb = cbuf + as * CBUF_BYTES;
      if (tid_e == 0) bulk_wait_read1();                       // the store that used this staging buffer two tiles ago
      named_bar_sync(1, NUM_EPI_THREADS);
      mbar_wait(&tfull_bar[as], (it >> 1) & 1);
      tc_fence_after();
      uint8_t* dst = cb + cgrp * CBLK_BYTES + row * 128;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * C + cgrp * 64 + ch * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          const float4 ba = bsm[ch * 8 + j * 2], bb = bsm[ch * 8 + j * 2 + 1];
          o.x = pack_bf16x2(fmaxf(__uint_as_float(r[j * 8 + 0]) + ba.x, 0.f), fmaxf(__uint_as_float(r[j * 8 + 1]) + ba.y, 0.f));
          o.y = pack_bf16x2(fmaxf(__uint_as_float(r[j * 8 + 2]) + ba.z, 0.f), fmaxf(__uint_as_float(r[j * 8 + 3]) + ba.w, 0.f));
          o.z = pack_bf16x2(fmaxf(__uint_as_float(r[j * 8 + 4]) + bb.x, 0.f), fmaxf(__uint_as_float(r[j * 8 + 5]) + bb.y, 0.f));
          o.w = pack_bf16x2(fmaxf(__uint_as_float(r[j * 8 + 6]) + bb.z, 0.f), fmaxf(__uint_as_float(r[j * 8 + 7]) + bb.w, 0.f));
          *reinterpret_cast<uint4*>(dst + (((ch * 4 + j) ^ sw) << 4)) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      fence_proxy_async();
      named_bar_sync(2, NUM_EPI_THREADS);
      if (tid_e == 0) {
#pragma unroll
        for (int blk = 0; blk < C / 64; ++blk) tma_store_2d(&tmOut, cb + blk * CBLK_BYTES, slab * C + blk * 64, tile * BM);
        bulk_commit();
      }
    }
    if (tid_e == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void pack_conv1_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cn) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;          // out [C][64]: k = kh*3+kw < 9, zero beyond
  if (i >= Cn * 64) return;
  const int c = i >> 6, k = i & 63;
  out[i] = __float2bfloat16_rn(k < 9 ? w[c * 9 + k] : 0.f);
}

}  // namespace

bool conv1_tc_supported(int Cn) { return Cn >= C && Cn % C == 0; }

int pack_conv1_weight(const float* w, bf16* out, int Cn, cudaStream_t stream) {
  pack_conv1_weight_kernel<<<(Cn * 64 + 255) / 256, 256, 0, stream>>>(w, out, Cn);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

int subsample_conv1_tc(const float* mel, const bf16* w_packed, const float* bias, bf16* out, int B, int n_mel, int Tin, int T1,
                       int T1p, int F1, int Cn, int num_sms, cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(Cn >= C && Cn % C == 0, "tensor-core conv1 works on slabs of 256 channels");
  const int slabs = Cn / C;
  CFM_REQUIRE(B > 0 && T1 > 0 && F1 > 0, "input too short for the subsampling convs");
  const int64_t M = static_cast<int64_t>(B) * T1p * F1;
  CFM_REQUIRE(M < (1LL << 31) - BM, "conv1 output has too many rows");
  CUtensorMap tmW, tmOut;
  {
    cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(Cn)};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, C};
    CFM_TRY(encode_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w_packed, dims, strides, box, "conv1 W"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(Cn), static_cast<cuuint64_t>(M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(Cn) * 2};
    cuuint32_t box[2] = {64, BM};
    CFM_TRY(encode_tmap(&tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides, box, "conv1 out"));
  }
  static std::atomic<uint64_t> attr_done{0};
  CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(conv1_tc_kernel), SMEM_BYTES));
  const int tiles = static_cast<int>(ceil_div(M, BM));
  Conv1Params p{mel, bias, n_mel, Tin, T1, T1p, F1, M, tiles};
  const int per_slab = num_sms / slabs > 0 ? num_sms / slabs : 1;
  dim3 grid(static_cast<unsigned>(tiles < per_slab ? tiles : per_slab), static_cast<unsigned>(slabs));
  CFM_CUDA_OK(launch_pdl(conv1_tc_kernel, grid, dim3(NUM_THREADS), SMEM_BYTES, stream, 1, tmW, tmOut, p));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
