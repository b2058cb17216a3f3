// fp32 SIMT GEMM with fused epilogues: the true-fp32 parity path (rtol 1e-4 gate) and the
// path for shapes the tensor-core kernel does not take.  C[M,N] = epi(A[M,K] * W[N,K]^T).
// Both operands are K-contiguous (activations row-major, nn.Linear weights [out,in]).
#include "ops.h"

namespace cfm {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;   // 256 threads, 4x4 outputs each

struct ARows {                       // how GEMM row m, column k map to an address in A
  const float* base;
  int lda;
  ConvGeom g;
  bool conv;
};

__device__ __forceinline__ int64_t a_row_offset(const ARows& a, int m) {
  if (!a.conv) return static_cast<int64_t>(m) * a.lda;
  int f2 = m % a.g.F2;
  int t2 = (m / a.g.F2) % a.g.T2;
  int b = m / (a.g.F2 * a.g.T2);
  return ((static_cast<int64_t>(b) * a.g.T1p + 2 * t2) * a.g.F1 + 2 * f2) * a.g.C;
}
__device__ __forceinline__ int64_t a_col_offset(const ARows& a, int k) {
  if (!a.conv) return k;
  int tap = k / a.g.C, c = k - tap * a.g.C;
  int kw = tap / 3, kh = tap - kw * 3;
  return (static_cast<int64_t>(kw) * a.g.F1 + kh) * a.g.C + c;
}

template <int EPI>
__global__ void __launch_bounds__(256) gemm_f32_kernel(ARows a, const float* __restrict__ W,
                                                       const float* __restrict__ bias, float* __restrict__ C,
                                                       int ldc, int M, int N, int K,
                                                       const float* __restrict__ res, float res_scale) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;        // loader: row 0..63, k offset 0,4,8,12
  const int ty = tid >> 4, tx = tid & 15;             // compute: 16x16 threads

  const int am = m0 + lr;
  const int64_t a_off = (am < M) ? a_row_offset(a, am) : 0;
  const int wn = n0 + lr;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
    const int k = k0 + lk;
    if (am < M && k < K) {                             // K % 4 == 0 is required by the launcher
      const float4 t = *reinterpret_cast<const float4*>(a.base + a_off + a_col_offset(a, k));
      av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
    }
    if (wn < N && k < K) {
      const float4 t = *reinterpret_cast<const float4*>(W + static_cast<int64_t>(wn) * K + k);
      wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lr] = av[i];
      Ws[lk + i][lr] = wv[i];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 af = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float4 wf = *reinterpret_cast<const float4*>(&Ws[kk][tx * TN]);
      const float ar[4] = {af.x, af.y, af.z, af.w};
      const float wr[4] = {wf.x, wf.y, wf.z, wf.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (EPI == EPI_BIAS_SWISH) v = swish_f(v);
      if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
      if (EPI == EPI_BIAS_RES) v = res[static_cast<int64_t>(m) * ldc + n] + res_scale * v;
      C[static_cast<int64_t>(m) * ldc + n] = v;
    }
  }
}

// Larger tile for the shapes that matter (M, N >= 128): 128 x 128 outputs per block, 8 x 8 per thread as four 4 x 4
// quadrants, K in steps of 16 through a double-buffered shared-memory tile with the next step's global loads in flight
// under the FMAs.  Every output is still ONE fmaf chain over k = 0 .. K-1, so results are bit-identical to the small kernel.
constexpr int GBM = 128, GBN = 128, GBK = 16;

template <int EPI>
__global__ void __launch_bounds__(256) gemm_f32_big_kernel(ARows a, const float* __restrict__ W, const float* __restrict__ bias,
                                                           float* __restrict__ C, int ldc, int M, int N, int K,
                                                           const float* __restrict__ res, float res_scale) {
  __shared__ float As[2][GBK][GBM + 4];
  __shared__ float Ws[2][GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
  const int lr = tid >> 1, lk = (tid & 1) * 8;        // loader: row 0..127, k offset 0 or 8 (two float4 each)
  const int ty = tid >> 4, tx = tid & 15;
  const int am = m0 + lr, wn = n0 + lr;
  const int64_t a_off = (am < M) ? a_row_offset(a, am) : 0;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = k0 + lk + 4 * h;
      ra[h] = (am < M && k < K) ? *reinterpret_cast<const float4*>(a.base + a_off + a_col_offset(a, k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[h] = (wn < N && k < K) ? *reinterpret_cast<const float4*>(W + static_cast<int64_t>(wn) * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = lk + 4 * h;
      As[buf][kk + 0][lr] = ra[h].x; As[buf][kk + 1][lr] = ra[h].y; As[buf][kk + 2][lr] = ra[h].z; As[buf][kk + 3][lr] = ra[h].w;
      Ws[buf][kk + 0][lr] = rw[h].x; Ws[buf][kk + 1][lr] = rw[h].y; Ws[buf][kk + 2][lr] = rw[h].z; Ws[buf][kk + 3][lr] = rw[h].w;
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += GBK) {
    const bool more = k0 + GBK < K;
    if (more) gload(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][kk][64 + tx * 4]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float wr[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (EPI == EPI_BIAS_SWISH) v = swish_f(v);
      if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
      if (EPI == EPI_BIAS_RES) v = res[static_cast<int64_t>(m) * ldc + n] + res_scale * v;
      C[static_cast<int64_t>(m) * ldc + n] = v;
    }
  }
}

}  // namespace

int gemm_f32(const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int M, int N,
             int K, Epilogue epi, const float* res, float res_scale, const ConvGeom* conv,
             cudaStream_t stream) {
  CFM_REQUIRE(M > 0 && N > 0 && K > 0, "empty GEMM");
  CFM_REQUIRE(K % 4 == 0 && (conv != nullptr || lda % 4 == 0), "fp32 GEMM needs K and lda multiples of 4");
  CFM_REQUIRE(conv == nullptr || conv->C % 4 == 0, "conv channels must be a multiple of 4");
  CFM_REQUIRE(epi != EPI_BIAS_RES || res != nullptr, "residual epilogue without residual");
  ARows a{A, lda, conv ? *conv : ConvGeom{}, conv != nullptr};
  dim3 grid(static_cast<unsigned>(ceil_div(N, BN)), static_cast<unsigned>(ceil_div(M, BM)));
  CFM_REQUIRE(grid.y <= 65535u * 32u, "M too large");
  if (grid.y > 65535u) {                  // split very tall problems over several launches
    const int rows_per = 65535 * BM;
    CFM_REQUIRE(conv == nullptr, "tall conv GEMM: split the batch on the caller side");
    for (int m = 0; m < M; m += rows_per) {
      int mm = (M - m < rows_per) ? (M - m) : rows_per;
      CFM_TRY(gemm_f32(A + static_cast<int64_t>(m) * lda, lda, W, bias, C + static_cast<int64_t>(m) * ldc, ldc,
                       mm, N, K, epi, res ? res + static_cast<int64_t>(m) * ldc : nullptr, res_scale, nullptr,
                       stream));
    }
    return 0;
  }
  if (ceil_div(M, GBM) * ceil_div(N, GBN) >= 256) {       // at least ~2 waves of 128 x 128 tiles; smaller problems keep 64 x 64
    dim3 gb(static_cast<unsigned>(ceil_div(N, GBN)), static_cast<unsigned>(ceil_div(M, GBM)));
    switch (epi) {
      case EPI_BIAS_SWISH: gemm_f32_big_kernel<EPI_BIAS_SWISH><<<gb, 256, 0, stream>>>(a, W, bias, C, ldc, M, N, K, res, res_scale); break;
      case EPI_BIAS_RELU: gemm_f32_big_kernel<EPI_BIAS_RELU><<<gb, 256, 0, stream>>>(a, W, bias, C, ldc, M, N, K, res, res_scale); break;
      case EPI_BIAS_RES: gemm_f32_big_kernel<EPI_BIAS_RES><<<gb, 256, 0, stream>>>(a, W, bias, C, ldc, M, N, K, res, res_scale); break;
      default: gemm_f32_big_kernel<EPI_BIAS><<<gb, 256, 0, stream>>>(a, W, bias, C, ldc, M, N, K, res, res_scale); break;
    }
    CFM_CUDA_OK(cudaGetLastError());
    return 0;
  }
  switch (epi) {
    case EPI_BIAS: gemm_f32_kernel<EPI_BIAS><<<grid, 256, 0, stream>>>(a, W, bias, C, ldc, M, N, K, res, res_scale); break;
    case EPI_BIAS_SWISH: gemm_f32_kernel<EPI_BIAS_SWISH><<<grid, 256, 0, stream>>>(a, W, bias, C, ldc, M, N, K, res, res_scale); break;
    case EPI_BIAS_RELU: gemm_f32_kernel<EPI_BIAS_RELU><<<grid, 256, 0, stream>>>(a, W, bias, C, ldc, M, N, K, res, res_scale); break;
    case EPI_BIAS_F32: gemm_f32_kernel<EPI_BIAS><<<grid, 256, 0, stream>>>(a, W, bias, C, ldc, M, N, K, res, res_scale); break;
    case EPI_BIAS_RES: gemm_f32_kernel<EPI_BIAS_RES><<<grid, 256, 0, stream>>>(a, W, bias, C, ldc, M, N, K, res, res_scale); break;
  }
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
