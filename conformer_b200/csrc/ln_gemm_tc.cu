// LayerNorm-fused projection for d_model = 256 on tcgen05 tensor cores:
//
//     out[M,N'] = epi( LN(x)[M,256] . W[N,256]^T + b )        N a multiple of 256
//
// used for the fused q|k|v projection (attention.py:15,78-80; epi = bias, N' = N = 768) and for
// pointwise_conv_1 + GLU (convolution.py:22-25; epi = GLU, N = 512, N' = 256).
// One CTA owns 128 rows for ALL output columns: 16 warps LayerNorm the rows once (fp32, in registers) and
// write the bf16 A operand directly in the 128B-swizzled MMA layout, where it stays; the weight is streamed
// through a TMA ring in chunks of 256 output columns (full-width tcgen05.mma 128x256x16), accumulators
// ping-pong between the two halves of TMEM so the epilogue of chunk c overlaps the MMAs of chunk c+1, and the
// bf16 result leaves through a swizzled staging buffer and TMA stores.  Compared with LayerNorm kernel + GEMM
// it removes one launch and the write + re-read of the normalised activations.
#include "ops.h"

namespace cfm {
namespace {

constexpr int D = 256, BM = 128, NC = 256;                 // model width, rows per CTA, GEMM columns per chunk
constexpr int KB = 64;
constexpr int KBLK_BYTES = BM * 128;
constexpr int XS_BYTES = (D / KB) * KBLK_BYTES;            // 64 KB
constexpr int STAGE_BYTES = NC * 128, STAGES = 3;          // [256 x 64] bf16 = 32 KB
constexpr int CBLK_BYTES = BM * 128;                       // staging block: 128 rows x 64 bf16
constexpr int CBUF_BYTES = 4 * CBLK_BYTES;                 // 64 KB (GLU uses half)
constexpr int NUM_EPI_WARPS = 16, NUM_EPI_THREADS = NUM_EPI_WARPS * 32, NUM_THREADS = 128 + NUM_EPI_THREADS;
constexpr int SMEM_BYTES = 1024 + XS_BYTES + STAGES * STAGE_BYTES + CBUF_BYTES + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "ln_gemm shared memory budget");

struct LnGemmParams {
  const float* x;
  const float* ln_g;
  const float* ln_b;
  const float* bias;     // [N] (GLU: packed like the weight rows)
  int M, N;
};

template <bool GLU>
__global__ void __launch_bounds__(NUM_THREADS, 1)
ln_gemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmC, LnGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* xs = smem;
  uint8_t* ring = xs + XS_BYTES;
  uint8_t* cbuf = ring + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(cbuf + CBUF_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* xs_full = empty_bar + STAGES;
  uint64_t* acc_full = xs_full + 1;      // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM;
  const int nchunks = p.N / NC;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(xs_full, NUM_EPI_WARPS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], NUM_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_whole_sm(tmem_slot);

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer
    uint32_t it = 0;
    for (int c = 0; c < nchunks; ++c) {
      for (int kb = 0; kb < D / KB; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          tma_load_2d(ring + s * STAGE_BYTES, &tmW, &full_bar[s], kb * KB, c * NC);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(BM, NC);
    mbar_wait(xs_full, 0);
    tc_fence_after();
    uint32_t it = 0;
    for (int c = 0; c < nchunks; ++c) {
      const int b = c & 1;
      mbar_wait(&acc_empty[b], ((c >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < D / KB; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = umma_desc_lo(smem_u32(xs + kb * KBLK_BYTES));
          const uint32_t sb = umma_desc_lo(smem_u32(ring + s * STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_f16_lo(tmem_base + b * NC, sa + k * 2, sb + k * 2, idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[s]);
          if (kb == D / KB - 1) tc_commit(&acc_full[b]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ LN prologue + epilogue (16 warps)
    const int ew = warp - 4;
    const int q = warp & 3, cgrp = ew >> 2;
    const int tid_e = threadIdx.x - 128;
    const bool leader = tid_e == 0;
    {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.ln_g) + 2 * lane);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.ln_g) + 2 * lane + 1);
      const float4 be0 = __ldg(reinterpret_cast<const float4*>(p.ln_b) + 2 * lane);
      const float4 be1 = __ldg(reinterpret_cast<const float4*>(p.ln_b) + 2 * lane + 1);
      float4 a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = m0 + ew * 8 + i;
        if (row < p.M) {
          const float4* xr = reinterpret_cast<const float4*>(p.x + static_cast<int64_t>(row) * D);
          a[i] = xr[2 * lane];
          b[i] = xr[2 * lane + 1];
        } else {
          a[i] = b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = ew * 8 + i;
        float s = (a[i].x + a[i].y) + (a[i].z + a[i].w) + (b[i].x + b[i].y) + (b[i].z + b[i].w);
        const float mean = warp_sum(s) * (1.0f / D);
        a[i].x -= mean; a[i].y -= mean; a[i].z -= mean; a[i].w -= mean;
        b[i].x -= mean; b[i].y -= mean; b[i].z -= mean; b[i].w -= mean;
        float v = (a[i].x * a[i].x + a[i].y * a[i].y) + (a[i].z * a[i].z + a[i].w * a[i].w) +
                  (b[i].x * b[i].x + b[i].y * b[i].y) + (b[i].z * b[i].z + b[i].w * b[i].w);
        const float rstd = rsqrtf(warp_sum(v) * (1.0f / D) + 1e-5f);
        uint4 o;
        o.x = pack_bf16x2(a[i].x * rstd * g0.x + be0.x, a[i].y * rstd * g0.y + be0.y);
        o.y = pack_bf16x2(a[i].z * rstd * g0.z + be0.z, a[i].w * rstd * g0.w + be0.w);
        o.z = pack_bf16x2(b[i].x * rstd * g1.x + be1.x, b[i].y * rstd * g1.y + be1.y);
        o.w = pack_bf16x2(b[i].z * rstd * g1.z + be1.z, b[i].w * rstd * g1.w + be1.w);
        *reinterpret_cast<uint4*>(xs + (lane >> 3) * KBLK_BYTES + r * 128 + (((lane & 7) ^ (r & 7)) << 4)) = o;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(xs_full);
    }
    const int row = q * 32 + lane;
    const uint32_t row_off = row * 128;
    const int sw = row & 7;
    for (int c = 0; c < nchunks; ++c) {
      const int b = c & 1;
      // bias of this warp's 64 GEMM columns (L1-resident after the first touch: 4 warps share each segment)
      const float4* bsrc = reinterpret_cast<const float4*>(p.bias + c * NC + cgrp * 64);
      if (leader) bulk_wait_read0();                           // previous chunk's TMA store has drained the staging
      named_bar_sync(1, NUM_EPI_THREADS);
      mbar_wait(&acc_full[b], (c >> 1) & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + b * NC + cgrp * 64;
      tmem_ld_32x32(t_row, r0);
      tmem_ld_32x32(t_row + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
      if (GLU) {
        // 32 value columns then their 32 gate columns -> 32 outputs at tile column cgrp*32
        float v[32];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 ba = __ldg(bsrc + g), bg = __ldg(bsrc + 8 + g);
          v[g * 4 + 0] = (__uint_as_float(r0[g * 4 + 0]) + ba.x) * sigmoid_tanh(__uint_as_float(r1[g * 4 + 0]) + bg.x);
          v[g * 4 + 1] = (__uint_as_float(r0[g * 4 + 1]) + ba.y) * sigmoid_tanh(__uint_as_float(r1[g * 4 + 1]) + bg.y);
          v[g * 4 + 2] = (__uint_as_float(r0[g * 4 + 2]) + ba.z) * sigmoid_tanh(__uint_as_float(r1[g * 4 + 2]) + bg.z);
          v[g * 4 + 3] = (__uint_as_float(r0[g * 4 + 3]) + ba.w) * sigmoid_tanh(__uint_as_float(r1[g * 4 + 3]) + bg.w);
        }
        uint8_t* blk = cbuf + (cgrp >> 1) * CBLK_BYTES + row_off;
        const int j0 = (cgrp & 1) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = pack_bf16x2(v[j * 8 + 0], v[j * 8 + 1]); o.y = pack_bf16x2(v[j * 8 + 2], v[j * 8 + 3]);
          o.z = pack_bf16x2(v[j * 8 + 4], v[j * 8 + 5]); o.w = pack_bf16x2(v[j * 8 + 6], v[j * 8 + 7]);
          *reinterpret_cast<uint4*>(blk + (((j0 + j) ^ sw) << 4)) = o;
        }
      } else {
        // 64 columns = one full 128-byte row of staging block cgrp
        uint8_t* blk = cbuf + cgrp * CBLK_BYTES + row_off;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t* rr = j < 4 ? &r0[j * 8] : &r1[(j - 4) * 8];
          const float4 ba = __ldg(bsrc + 2 * j), bb = __ldg(bsrc + 2 * j + 1);
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(rr[0]) + ba.x, __uint_as_float(rr[1]) + ba.y);
          o.y = pack_bf16x2(__uint_as_float(rr[2]) + ba.z, __uint_as_float(rr[3]) + ba.w);
          o.z = pack_bf16x2(__uint_as_float(rr[4]) + bb.x, __uint_as_float(rr[5]) + bb.y);
          o.w = pack_bf16x2(__uint_as_float(rr[6]) + bb.z, __uint_as_float(rr[7]) + bb.w);
          *reinterpret_cast<uint4*>(blk + ((j ^ sw) << 4)) = o;
        }
      }
      fence_proxy_async();
      named_bar_sync(1, NUM_EPI_THREADS);
      if (leader && m0 < p.M) {
        constexpr int OUT_PER_CHUNK = GLU ? NC / 2 : NC;
        for (int blk = 0; blk < OUT_PER_CHUNK / 64; ++blk)
          tma_store_2d(&tmC, cbuf + blk * CBLK_BYTES, c * OUT_PER_CHUNK + blk * 64, m0);
        bulk_commit();
      }
    }
    if (leader) bulk_wait_read0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool ln_gemm_supported(int d, int N) { return d == D && N % NC == 0 && N > 0; }

int ln_gemm_bf16(const float* x, const float* ln_g, const float* ln_b, const bf16* W, const float* bias, bf16* out,
                 int ldc, int M, int N, int d, bool glu, cudaStream_t stream) {
  CFM_TRY(tc_init());
  CFM_REQUIRE(ln_gemm_supported(d, N), "ln_gemm is specialised for d_model = 256 and N % 256 == 0");
  CFM_REQUIRE(M > 0 && ldc % 8 == 0, "bad ln_gemm shape");
  CFM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out) |
                reinterpret_cast<uintptr_t>(ln_g) | reinterpret_cast<uintptr_t>(ln_b) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0,
              "ln_gemm operands must be 16-byte aligned");
  CUtensorMap tmW, tmC;
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(N)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * 2};
    cuuint32_t box[2] = {KB, NC};
    CFM_TRY(encode_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, dims, strides, box, "ln_gemm W"));
  }
  {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(glu ? N / 2 : N), static_cast<cuuint64_t>(M)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldc) * 2};
    cuuint32_t box[2] = {64, BM};
    CFM_TRY(encode_tmap(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides, box, "ln_gemm C"));
  }
  static bool attr_done = false;
  if (!attr_done) {
    CFM_CUDA_OK(cudaFuncSetAttribute(ln_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CFM_CUDA_OK(cudaFuncSetAttribute(ln_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_done = true;
  }
  LnGemmParams p{x, ln_g, ln_b, bias, M, N};
  const int grid = static_cast<int>(ceil_div(M, BM));
  if (glu) ln_gemm_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmW, tmC, p);
  else ln_gemm_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmW, tmC, p);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
