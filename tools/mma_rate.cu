// Micro-benchmark (debug tool, not part of the library): clocks per tcgen05.mma (M = 128, K = 16, bf16) as a function of
// N, of where A comes from (shared memory / tensor memory) and of how many distinct accumulators the MMAs rotate over.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I conformer_b200/csrc tools/mma_rate.cu -o tools/mma_rate
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

using namespace cfm;

// the single-instruction .x32 load (the library's tmem_ld_32x32 issues four .x8 loads)
__device__ __forceinline__ void tmem_ld_32x32_raw(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

template <int N, bool A_TMEM, int NACC>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* a_s = smem;                 // [128 x 64] bf16 (16 KB)
  uint8_t* b_s = smem + 16384;         // [256 x 64] bf16 (32 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_whole_sm(slot);
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t sa = umma_desc_lo(smem_u32(a_s)), sb = umma_desc_lo(smem_u32(b_s));
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {            // rep 0 warms up
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t d = tmem + ((i * 4 + k) % NACC) * N;
            if (A_TMEM) tc_mma_f16_ts(d, tmem + 448 + k * 8, sb + k * 2, idesc, 1u);
            else tc_mma_f16_lo(d, sa + k * 2, sb + k * 2, idesc, 1u);
          }
        }
        tc_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, rep & 1);
      t1 = clock64();
    }
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


// The attention kernel's S pattern: batches of 20 MMAs (N = 64, A in TMEM at 20 different column offsets, B = 5 blocks
// of [64 x 64] x 4 K steps), each batch followed by a commit.  Reports clocks per MMA to ISSUE a batch and to COMPLETE it.
template <int NB, bool DYN = false>
__global__ void __launch_bounds__(128, 1) mma_batch_kernel(long long* out, int batches, uint32_t tmem_zero) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* b_s = smem;                 // 5 x [NB x 64] bf16
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 5 * NB * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 6);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 5 * NB * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    for (int i = 2; i < 5; ++i) mbar_init(bar + i, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_whole_sm(slot) + tmem_zero;
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, NB);
    const uint32_t sb = umma_desc_lo(smem_u32(b_s));
    long long t_issue = 0, t_done = 0;
    for (int bt = 0; bt < batches; ++bt) {
      const long long t0 = clock64();
      // DYN: operand bases change from batch to batch (ring stage, accumulator buffer), as in the attention kernel,
      // so every MMA carries its own uniform address arithmetic instead of loop-invariant registers
      const uint32_t opaque_zero = static_cast<uint32_t>(clock64()) & 0u;      // a zero the compiler cannot fold
      const uint32_t sbd = DYN ? sb + opaque_zero : sb;
      const uint32_t tad = DYN ? tmem + 256 + opaque_zero : tmem + 256;
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 5; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16_ts(tmem + (bt & 1) * NB, tad + (kb * 4 + k) * 8, sbd + kb * (NB * 128 >> 4) + k * 2, idesc,
                          (kb | k) != 0 ? 1u : 0u);
        if (DYN) {                     // the attention kernel's commit pattern: 2 after S, 3 after P.V (dummy barriers here)
          tc_commit(bar + 2);
          tc_mma_f16_ts(tmem + 128, tad, sbd, idesc, 1u);
          tc_mma_f16_ts(tmem + 128, tad + 8, sbd + 2, idesc, 1u);
          tc_mma_f16_ts(tmem + 192, tad + 16, sbd + 4, idesc, 1u);
          tc_mma_f16_ts(tmem + 192, tad + 24, sbd + 6, idesc, 1u);
          tc_commit(bar + 3);
          tc_commit(bar + 4);
        }
        tc_commit(bar);
      }
      __syncwarp();
      const long long t1 = clock64();
      mbar_wait(bar, bt & 1);
      const long long t2 = clock64();
      if (bt > 0) { t_issue += t1 - t0; t_done += t2 - t0; }
    }
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { out[0] = t_issue; out[1] = t_done; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int NB, bool DYN = false>
void run_batch(int grid, long long* d_out) {
  const int batches = 65, smem = 1024 + 5 * NB * 128 + 128;
  cudaFuncSetAttribute(mma_batch_kernel<NB, DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_batch_kernel<NB, DYN><<<grid, 128, smem>>>(d_out, batches, 0u);
  cudaError_t e = cudaDeviceSynchronize();
  long long r[2] = {0, 0};
  cudaMemcpy(r, d_out, 16, cudaMemcpyDeviceToHost);
  printf("batch of 20 MMAs N=%d A tmem%s, grid %3d: issue %6.1f clk/MMA, issue+complete %6.1f clk/MMA (floor %d)%s\n", NB,
         DYN ? " (+4 MMAs, per-batch bases, 4 commits)" : "", grid,
         r[0] / (64.0 * (DYN ? 24 : 20)), r[1] / (64.0 * (DYN ? 24 : 20)), NB / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}


// Same batches while 8 other warps read TMEM continuously (tcgen05.ld, 32 columns each, as the softmax warps do), or,
// with SMEM_NOISE, read shared memory instead: is the MMA slowed by TMEM-read or shared-memory-port contention?
template <bool A_TMEM, int NOISE>     // NOISE: 0 none, 1 tcgen05.ld, 2 ld.shared, 3 tcgen05.st
__global__ void __launch_bounds__(384, 1) mma_noise_kernel(long long* out, int batches, uint32_t tmem_zero, const uint8_t* gsrc = nullptr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* b_s = smem;                 // 5 x [64 x 64] bf16
  uint8_t* a_s = smem + 5 * 8192;      // 5 x [128 x 64] bf16
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 5 * 8192 + 5 * 16384);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 6);
  volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (5 * 8192 + 5 * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 5, 1);
    fence_mbar_init();
    *stop = 0;
  }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_whole_sm(slot) + tmem_zero;
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    const uint32_t sb = umma_desc_lo(smem_u32(b_s)), sa = umma_desc_lo(smem_u32(a_s));
    long long t_done = 0;
    for (int bt = 0; bt < batches; ++bt) {
      const long long t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 5; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (A_TMEM) tc_mma_f16_ts(tmem + (bt & 1) * 64, tmem + 256 + (kb * 4 + k) * 8, sb + kb * (8192 >> 4) + k * 2, idesc, (kb | k) != 0 ? 1u : 0u);
            else tc_mma_f16_lo(tmem + (bt & 1) * 64, sa + kb * (16384 >> 4) + k * 2, sb + kb * (8192 >> 4) + k * 2, idesc, (kb | k) != 0 ? 1u : 0u);
          }
        tc_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, bt & 1);
      if (bt > 0) t_done += clock64() - t0;
    }
    if (lane == 0) { *stop = 1; if (blockIdx.x == 0) out[0] = t_done; }
  } else if (warp == 4 && NOISE == 5) {
    // bulk async copies global -> shared (40 KB per round, the attention kernel's K|E tile) into the A region, back to back
    uint64_t* cbar = bar + 5;
    long long n = 0;
    if (lane == 0) {
      while (*stop == 0) {
        mbar_expect_tx(cbar, 5 * 8192);
        for (int i = 0; i < 5; ++i)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(a_s + i * 8192)), "l"(gsrc + ((n * 5 + i) % 64) * 8192), "r"(8192u), "r"(smem_u32(cbar))
                       : "memory");
        while (!mbar_try_wait(cbar, n & 1)) {
        }
        ++n;
      }
      if (blockIdx.x == 0) out[1] = n * 10;     // reported as "iterations" x 4096 B x 8 -> bytes = n * 40960
    }
  } else if (warp >= 4 && NOISE != 0 && NOISE != 5) {
    const uint32_t tl = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 128;
    uint32_t acc = 0;
    long long n = 0;
    while (*stop == 0) {
      uint32_t r[32];
      if (NOISE == 1) {
        tmem_ld_32x32_raw(tl, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; k += 4) acc ^= r[k];
      } else if (NOISE == 4) {                 // arithmetic only: ex2 + fma chains, as the softmax warps run
        float a = __uint_as_float(0x3f800000u + acc), b = 0.5f;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          float e;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a - b));
          b = fmaf(e, 0.25f, b);
          a = fmaf(a, 0.999f, e);
        }
        acc += __float_as_uint(a + b);
      } else if (NOISE == 3) {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = acc + i;
        tmem_st_32x32(tl + 64, r);
        tmem_st_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 v = *reinterpret_cast<const uint4*>(a_s + ((threadIdx.x * 16 + i * 6144) & 65535));
          acc += v.x ^ v.w;
        }
      }
      ++n;
    }
    if (acc == 0x12345678u) out[3] = acc;
    if (warp == 4 && lane == 0 && blockIdx.x == 0) out[1] = n;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <bool A_TMEM, int NOISE>
void run_noise(const char* name, long long* d_out, int grid = 1) {
  const int batches = 65, smem = 1024 + 5 * 8192 + 5 * 16384 + 128;
  static uint8_t* gsrc = nullptr;
  if (gsrc == nullptr) { cudaMalloc(&gsrc, 64 * 8192); cudaMemset(gsrc, 0x3c, 64 * 8192); }
  cudaFuncSetAttribute(mma_noise_kernel<A_TMEM, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_noise_kernel<A_TMEM, NOISE><<<grid, 384, smem>>>(d_out, batches, 0u, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  long long r[2] = {0, 0};
  cudaMemcpy(r, d_out, 16, cudaMemcpyDeviceToHost);
  const double clk = static_cast<double>(r[0]);
  printf("%-46s %6.1f clk/MMA; noise: %.0f iterations/warp = %.1f B/clk %s\n", name, clk / (64.0 * 20), static_cast<double>(r[1]),
         NOISE == 0 ? 0.0 : static_cast<double>(r[1]) * 8 * 4096 / clk, e == cudaSuccess ? "" : cudaGetErrorString(e));
}


// TMEM read / write rate of NW warps with the tensor pipe idle (32 lanes x NC columns per instruction, one wait each)
template <int NC, bool STORE>
__global__ void __launch_bounds__(512, 1) tmem_rw_kernel(long long* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tl = (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (NC == 32) {
      uint32_t r[32];
      if (STORE) {
#pragma unroll
        for (int k = 0; k < 32; ++k) r[k] = acc + k;
        tmem_st_32x32(tl, r);
        tmem_st_wait();
      } else {
        tmem_ld_32x32_raw(tl, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; k += 4) acc ^= r[k];
      }
    } else {
      uint32_t r[8];
      if (NC == 8) {
        tmem_ld_32x8(tl + (i & 3) * 8, r);
        tmem_ld_wait_for(r);
        acc ^= r[0] ^ r[4];
      } else {                      // NC == 33: four .x8 loads, one wait
        uint32_t q[32];
        tmem_ld_32x32(tl, q);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; k += 4) acc ^= q[k];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (acc == 0x12345678u) out[3] = acc;
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (warp == 0) tmem_dealloc(0, 512);
}
template <int NC, bool STORE>
void run_rw(int nwarps, long long* d_out) {
  const int iters = 2000;
  tmem_rw_kernel<NC, STORE><<<1, nwarps * 32, 0>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long clk = 0;
  cudaMemcpy(&clk, d_out, 8, cudaMemcpyDeviceToHost);
  const int cols = NC == 33 ? 32 : NC;
  printf("tensor pipe idle, %2d warps tcgen05.%s %s: %6.1f clk per load+wait, %6.1f B/clk per SM %s\n", nwarps,
         STORE ? "st" : "ld", NC == 33 ? "4 x .x8" : NC == 32 ? ".x32   " : ".x8    ", static_cast<double>(clk) / iters,
         static_cast<double>(iters) * nwarps * cols * 128 / clk,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// P.V form: B operand MN-major ([16 keys x 64] rows per K step, 2048 B apart), N = 64, A in TMEM
__global__ void __launch_bounds__(128, 1) mma_bmn_kernel(long long* out, int iters, uint32_t tmem_zero) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* b_s = smem;                 // V tile [64 keys x 64] bf16 (8 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_whole_sm(slot) + tmem_zero;
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16_bmn(128, 64);
    const uint32_t sb = umma_desc_lo(smem_u32(b_s));
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_f16_ts(tmem + 128, tmem + 256 + k * 8, sb + k * (2048 >> 4), idesc, 1u);
        }
        tc_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, rep & 1);
      t1 = clock64();
    }
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
void run_bmn(long long* d_out) {
  const int iters = 256, smem = 1024 + 8192 + 64;
  mma_bmn_kernel<<<1, 128, smem>>>(d_out, iters, 0u);
  cudaError_t e = cudaDeviceSynchronize();
  long long clk = 0;
  cudaMemcpy(&clk, d_out, 8, cudaMemcpyDeviceToHost);
  printf("N=64  A tmem, B MN-major (P.V form)          : %7.1f clk per MMA (floor 32)%s\n", static_cast<double>(clk) / (iters * 4),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int N, bool A_TMEM, int NACC>
void run(const char* name, int grid, long long* d_out) {
  const int iters = 256, smem = 1024 + 16384 + 32768 + 64;
  cudaFuncSetAttribute(mma_rate_kernel<N, A_TMEM, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate_kernel<N, A_TMEM, NACC><<<grid, 128, smem>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long clk = 0;
  cudaMemcpy(&clk, d_out, 8, cudaMemcpyDeviceToHost);
  printf("%-34s grid %3d: %7.1f clk per MMA (floor %d)%s\n", name, grid, static_cast<double>(clk) / (iters * 4), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 64);
  for (int nw : {1, 4, 8, 16}) run_rw<32, false>(nw, d_out);
  for (int nw : {4, 8, 16}) run_rw<8, false>(nw, d_out);
  for (int nw : {1, 4, 8, 16}) run_rw<33, false>(nw, d_out);
  for (int nw : {4, 8}) run_rw<32, true>(nw, d_out);
  run_noise<true, 0>("A tmem, no noise", d_out);
  run_noise<true, 1>("A tmem, 8 warps tcgen05.ld x32", d_out);
  run_noise<true, 3>("A tmem, 8 warps tcgen05.st x32", d_out);
  run_noise<true, 2>("A tmem, 8 warps ld.shared (8 x 16 B)", d_out);
  run_noise<true, 4>("A tmem, 8 warps ex2 + fma chains", d_out);
  run_noise<true, 5>("A tmem, bulk copies global->smem (1 CTA)", d_out);
  run_noise<true, 5>("A tmem, bulk copies global->smem (148 CTAs)", d_out, 148);
  run_noise<false, 0>("A smem, no noise", d_out);
  run_noise<false, 1>("A smem, 8 warps tcgen05.ld x32", d_out);
  run_noise<false, 2>("A smem, 8 warps ld.shared (8 x 16 B)", d_out);
  run_bmn(d_out);
  run_batch<64, true>(1, d_out);
  run_batch<64>(1, d_out);
  run_batch<64>(148, d_out);
  run_batch<128>(1, d_out);
  for (int grid : {1}) {
    run<64, false, 1>("N=64  A smem, 1 accumulator", grid, d_out);
    run<64, true, 1>("N=64  A tmem, 1 accumulator", grid, d_out);
    run<64, true, 2>("N=64  A tmem, 2 accumulators", grid, d_out);
    run<64, true, 4>("N=64  A tmem, 4 accumulators", grid, d_out);
    run<128, false, 1>("N=128 A smem, 1 accumulator", grid, d_out);
    run<128, true, 1>("N=128 A tmem, 1 accumulator", grid, d_out);
    run<128, true, 2>("N=128 A tmem, 2 accumulators", grid, d_out);
    run<256, false, 1>("N=256 A smem, 1 accumulator", grid, d_out);
    run<256, true, 1>("N=256 A tmem, 1 accumulator", grid, d_out);
  }
  return 0;
}
