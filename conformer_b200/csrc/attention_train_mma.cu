// Attention backward of the training step with the matrix products on the tensor cores (bf16 mode).
//
// Same mathematics and the same flash-style tiling as attention_train.cu (nothing of size T x T is stored; 64 x 64
// (query x key) tiles; the relative shift is an index into a 127-row window of the projected table), but every product of
// a tile is a warp-level bf16 MMA with fp32 accumulation on shared-memory operands instead of CUDA-core dot products:
//
//   query-side kernel, one CTA = 64 queries of one (b, h), loop over key tiles:
//     S  = Qc K^T            [64 x 64]      U = Qp Pw^T  [64 x 128] (all window rows; s_ij = S_ij + U[i][j - i + 63])
//     dP = dO V^T            [64 x 64]
//     element-wise: P = exp(s - lse), ds = P (f dP - D) / sqrt(dh)  ->  ds as a bf16 tile, and "unshifted" into a
//     [64 x 128] tile dsu[i][j - i + 63] (the transpose of what the shift does to U)
//     dQc += ds K,  dQp += dsu Pw,  dPw[128 x 64] += dsu^T Qp    (window gradient: rows < 64 leave through atomics after
//                                                                  each tile, rows >= 64 carry over to the next window)
//   key-side kernel, one CTA = 64 keys, loop over query tiles: the same S / U / dP / element-wise stage, then
//     dV += Pd^T dO,  dK += ds^T Qc
//
// This is the interim tensor-core form: warp-level mma.sync.m16n8k16 with ldmatrix operand loads (the HMMA path).  It
// takes the backward from the fp32 FMA pipe to the tensor pipe; a tcgen05 version with the accumulators in tensor memory
// is the next step.  Profile-guided history of the two kernels at C4 (14.2 -> 9.0 ms per step):
// one thread per column for the du / dv sums instead of contended shared-memory float atomics (compare-and-swap loops),
// all tiles of an iteration loaded in one pass (one memory latency instead of one per tile), ldmatrix + mma.sync instead
// of nvcuda::wmma (whose loads compiled to generic 32-bit loads).  The fp32 gradient-parity mode keeps the CUDA-core
// kernels of attention_train.cu.
#include "train.h"

namespace cfm {
namespace {

constexpr int TQ = 64, TK = 64, DHP = 64, WINR = 128;             // 127 window rows used, 128 stored
constexpr int LDB = DHP + 8;                                       // bf16 operand tiles: 72 elements = 144 B per row
constexpr int LDU = WINR + 8;                                      // bf16 dsu tile: 136 elements
constexpr int LDF = TK + 4;                                        // fp32 score tiles: 68 floats
constexpr int LDUF = WINR + 4;                                     // fp32 U tile: 132 floats
constexpr int THREADS = 256;

struct MmaArgs {
  const bf16* qkv;
  const bf16* pp;
  const float* u;
  const float* v;
  const uint8_t* key_pad;
  const float* lse;
  const float* D;
  const float* dctx;
  int Tn, H, dh;
  float scale;
  Drop dr;
};

// 16 x 16 x 16 tile products as two mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with ldmatrix operand loads.  (The
// nvcuda::wmma loads this file started with compiled to 32-bit GENERIC loads plus MOVM transposes - 528 LD.E for 144 HMMA in
// the query-side kernel, a quarter of its stall samples; one ldmatrix.x4 replaces sixteen of them.)
struct FragA { uint32_t r[4]; };                 // A[16 m][16 k]
struct FragB { uint32_t r[4]; };                 // B[16 k][16 n]: r[0], r[1] = columns 0-7; r[2], r[3] = columns 8-15
struct FragC { float c[8]; };                    // C[16][16]: c[0..3] = columns 0-7 (rows g, g+8), c[4..7] = columns 8-15
__device__ __forceinline__ void frag_zero(FragC& c) {
#pragma unroll
  for (int i = 0; i < 8; ++i) c.c[i] = 0.f;
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
// A row-major: p = &A[m0][k0], pitch ld (elements)
__device__ __forceinline__ void ld_a(FragA& a, const bf16* p, int ld) {
  const int l = threadIdx.x & 31;
  ldsm4(a.r, p + ((l & 7) + 8 * ((l >> 3) & 1)) * ld + 8 * (l >> 4));
}
// A stored transposed: element A[m][k] at p[k * ld + m], p = &smem[k0][m0]
__device__ __forceinline__ void ld_a_t(FragA& a, const bf16* p, int ld) {
  const int l = threadIdx.x & 31, mi = l >> 3;
  ldsm4_t(a.r, p + ((l & 7) + 8 * (mi >> 1)) * ld + 8 * (mi & 1));
}
// B stored [n][k] (k contiguous): p = &smem[n0][k0]
__device__ __forceinline__ void ld_b_nk(FragB& b, const bf16* p, int ld) {
  const int l = threadIdx.x & 31, mi = l >> 3;
  ldsm4(b.r, p + ((l & 7) + 8 * (mi >> 1)) * ld + 8 * (mi & 1));
}
// B stored [k][n] (n contiguous): p = &smem[k0][n0]
__device__ __forceinline__ void ld_b_kn(FragB& b, const bf16* p, int ld) {
  const int l = threadIdx.x & 31, mi = l >> 3;
  ldsm4_t(b.r, p + ((l & 7) + 8 * (mi & 1)) * ld + 8 * (mi >> 1));
}
__device__ __forceinline__ void frag_mma(FragC& c, const FragA& a, const FragB& b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c.c[0]), "+f"(c.c[1]), "+f"(c.c[2]), "+f"(c.c[3])
               : "r"(a.r[0]), "r"(a.r[1]), "r"(a.r[2]), "r"(a.r[3]), "r"(b.r[0]), "r"(b.r[1]));
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c.c[4]), "+f"(c.c[5]), "+f"(c.c[6]), "+f"(c.c[7])
               : "r"(a.r[0]), "r"(a.r[1]), "r"(a.r[2]), "r"(a.r[3]), "r"(b.r[2]), "r"(b.r[3]));
}
// C <-> fp32 row-major shared memory: p = &out[m0][n0]
__device__ __forceinline__ void frag_store(float* p, int ld, const FragC& c) {
  const int l = threadIdx.x & 31, g = l >> 2, t = l & 3;
  *reinterpret_cast<float2*>(p + g * ld + 2 * t) = make_float2(c.c[0], c.c[1]);
  *reinterpret_cast<float2*>(p + (g + 8) * ld + 2 * t) = make_float2(c.c[2], c.c[3]);
  *reinterpret_cast<float2*>(p + g * ld + 8 + 2 * t) = make_float2(c.c[4], c.c[5]);
  *reinterpret_cast<float2*>(p + (g + 8) * ld + 8 + 2 * t) = make_float2(c.c[6], c.c[7]);
}
__device__ __forceinline__ void frag_load(FragC& c, const float* p, int ld) {
  const int l = threadIdx.x & 31, g = l >> 2, t = l & 3;
  const float2 v0 = *reinterpret_cast<const float2*>(p + g * ld + 2 * t), v1 = *reinterpret_cast<const float2*>(p + (g + 8) * ld + 2 * t);
  const float2 v2 = *reinterpret_cast<const float2*>(p + g * ld + 8 + 2 * t), v3 = *reinterpret_cast<const float2*>(p + (g + 8) * ld + 8 + 2 * t);
  c.c[0] = v0.x; c.c[1] = v0.y; c.c[2] = v1.x; c.c[3] = v1.y; c.c[4] = v2.x; c.c[5] = v2.y; c.c[6] = v3.x; c.c[7] = v3.y;
}

// dst[r][c] (bf16, pitch LDB) = src[(row0 + r) * stride + c] (+ add[c]) for row0 + r in [lo, hi), zeros elsewhere / c >= dh
__device__ __forceinline__ void load_tile(bf16* dst, const bf16* src, int64_t stride, int row0, int lo, int hi, int rows, int dh,
                                          const float* add) {
  for (int idx = threadIdx.x; idx < rows * (DHP / 8); idx += THREADS) {
    const int r = idx >> 3, c8 = (idx & 7) * 8, row = row0 + r;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (row >= lo && row < hi && c8 < dh) {
      o = *reinterpret_cast<const uint4*>(src + static_cast<int64_t>(row) * stride + c8);
      if (add != nullptr) {
        uint32_t* w = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          w[e] = pack_bf16x2(__uint_as_float(w[e] << 16) + add[c8 + 2 * e], __uint_as_float(w[e] & 0xffff0000u) + add[c8 + 2 * e + 1]);
      }
    }
    *reinterpret_cast<uint4*>(dst + r * LDB + c8) = o;
  }
}
__device__ __forceinline__ void load_tile_f32(bf16* dst, const float* src, int64_t stride, int row0, int hi, int rows, int dh) {
  for (int idx = threadIdx.x; idx < rows * (DHP / 8); idx += THREADS) {
    const int r = idx >> 3, c8 = (idx & 7) * 8, row = row0 + r;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (row < hi && c8 < dh) {
      const float4 a = *reinterpret_cast<const float4*>(src + static_cast<int64_t>(row) * stride + c8);
      const float4 b = *reinterpret_cast<const float4*>(src + static_cast<int64_t>(row) * stride + c8 + 4);
      o = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
    *reinterpret_cast<uint4*>(dst + r * LDB + c8) = o;
  }
}

// The tiles of one loop iteration in ONE pass: every global load is issued before the first shared-memory store, so the
// iteration pays one memory latency instead of one per tile (the per-tile loaders above made the load wait of each
// tile the top stall of both kernels).
__device__ __forceinline__ uint4 add_bias8(uint4 o, const float* add, int c8) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int e = 0; e < 4; ++e)
    w[e] = pack_bf16x2(__uint_as_float(w[e] << 16) + add[c8 + 2 * e], __uint_as_float(w[e] & 0xffff0000u) + add[c8 + 2 * e + 1]);
  return o;
}
// query side, per key tile: K, V rows j0 .. j0+63 and the 128-row window of the projected table starting at row r0
__device__ __forceinline__ void load_key_tiles(bf16* Kt, bf16* Vt, bf16* Pw, const bf16* kbase, const bf16* vbase, int64_t stride, int j0,
                                               int Tn, const bf16* pp, int64_t ppstride, int r0, int R, int dh) {
  static_assert(TK * (DHP / 8) == 2 * THREADS && WINR * (DHP / 8) == 4 * THREADS, "loader shape");
  uint4 k[2], v[2], w[4];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + it * THREADS, r = idx >> 3, c8 = (idx & 7) * 8, row = j0 + r;
    k[it] = v[it] = make_uint4(0u, 0u, 0u, 0u);
    if (row < Tn && c8 < dh) {
      k[it] = *reinterpret_cast<const uint4*>(kbase + static_cast<int64_t>(row) * stride + c8);
      v[it] = *reinterpret_cast<const uint4*>(vbase + static_cast<int64_t>(row) * stride + c8);
    }
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = threadIdx.x + it * THREADS, r = idx >> 3, c8 = (idx & 7) * 8, row = r0 + r;
    w[it] = make_uint4(0u, 0u, 0u, 0u);
    if (row >= 0 && row < R && c8 < dh) w[it] = *reinterpret_cast<const uint4*>(pp + static_cast<int64_t>(row) * ppstride + c8);
  }
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + it * THREADS, r = idx >> 3, c8 = (idx & 7) * 8;
    *reinterpret_cast<uint4*>(Kt + r * LDB + c8) = k[it];
    *reinterpret_cast<uint4*>(Vt + r * LDB + c8) = v[it];
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = threadIdx.x + it * THREADS, r = idx >> 3, c8 = (idx & 7) * 8;
    *reinterpret_cast<uint4*>(Pw + r * LDB + c8) = w[it];
  }
}
// key side, per query tile: q rows i0 .. i0+63 (+ u -> Qc, + v -> Qp), dctx rows (fp32 -> bf16) and the window of the table
__device__ __forceinline__ void load_query_tiles(bf16* Qc, bf16* Qp, bf16* dO, bf16* Pw, const bf16* qbase, int64_t stride, const float* u,
                                                 const float* vb, const float* dctx, int64_t dstride, int i0, int Tn, const bf16* pp,
                                                 int64_t ppstride, int r0, int R, int dh) {
  uint4 q[2], w[4];
  float4 g[2][2];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + it * THREADS, r = idx >> 3, c8 = (idx & 7) * 8, row = i0 + r;
    q[it] = make_uint4(0u, 0u, 0u, 0u);
    g[it][0] = g[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < Tn && c8 < dh) {
      q[it] = *reinterpret_cast<const uint4*>(qbase + static_cast<int64_t>(row) * stride + c8);
      g[it][0] = *reinterpret_cast<const float4*>(dctx + static_cast<int64_t>(row) * dstride + c8);
      g[it][1] = *reinterpret_cast<const float4*>(dctx + static_cast<int64_t>(row) * dstride + c8 + 4);
    }
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = threadIdx.x + it * THREADS, r = idx >> 3, c8 = (idx & 7) * 8, row = r0 + r;
    w[it] = make_uint4(0u, 0u, 0u, 0u);
    if (row >= 0 && row < R && c8 < dh) w[it] = *reinterpret_cast<const uint4*>(pp + static_cast<int64_t>(row) * ppstride + c8);
  }
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + it * THREADS, r = idx >> 3, c8 = (idx & 7) * 8, row = i0 + r;
    const bool ok = row < Tn && c8 < dh;                          // rows past T stay zero (no bias either), as load_tile does
    *reinterpret_cast<uint4*>(Qc + r * LDB + c8) = ok ? add_bias8(q[it], u, c8) : q[it];
    *reinterpret_cast<uint4*>(Qp + r * LDB + c8) = ok ? add_bias8(q[it], vb, c8) : q[it];
    *reinterpret_cast<uint4*>(dO + r * LDB + c8) = make_uint4(pack_bf16x2(g[it][0].x, g[it][0].y), pack_bf16x2(g[it][0].z, g[it][0].w),
                                                              pack_bf16x2(g[it][1].x, g[it][1].y), pack_bf16x2(g[it][1].z, g[it][1].w));
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int idx = threadIdx.x + it * THREADS, r = idx >> 3, c8 = (idx & 7) * 8;
    *reinterpret_cast<uint4*>(Pw + r * LDB + c8) = w[it];
  }
}

// C[16 rb .. , 16 cb ..] = A[rows 16 rb][K] . B^T, A row-major (pitch lda), B stored [n][k] (pitch ldb): K = 64
__device__ __forceinline__ void mma_abt(const bf16* A, int lda, const bf16* Bm, int ldb, int rb, int cb, float* out, int ldo) {
  FragC c;
  frag_zero(c);
#pragma unroll
  for (int k = 0; k < DHP; k += 16) {
    FragA a;
    FragB b;
    ld_a(a, A + rb * 16 * lda + k, lda);
    ld_b_nk(b, Bm + cb * 16 * ldb + k, ldb);
    frag_mma(c, a, b);
  }
  frag_store(out + rb * 16 * ldo + cb * 16, ldo, c);
}

// the shared first half of both kernels: S, U, dP on the tensor cores, then the per-entry stage.
// Writes Pd (bf16, nullable), ds (bf16) and, if dsu != nullptr, the unshifted ds (bf16 [64][LDU], zeroed first).
__device__ __forceinline__ void scores_and_grads(const MmaArgs& p, const bf16* Qc, const bf16* Qp, const bf16* dO, const bf16* Kt, const bf16* Vt,
                                                 const bf16* Pw, float* Ss, float* Us, float* DPs, const float* lse_s, const float* D_s,
                                                 bf16* Pd_b, bf16* dS_b, bf16* dSu_b, int b, int h, int i0, int j0) {
  const DropK dk = p.dr.load();
  const int warp = threadIdx.x >> 5;
  const int rb = warp >> 1;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int cb = (warp & 1) * 2 + e;
    mma_abt(Qc, LDB, Kt, LDB, rb, cb, Ss, LDF);
    mma_abt(dO, LDB, Vt, LDB, rb, cb, DPs, LDF);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) mma_abt(Qp, LDB, Pw, LDB, rb, (warp & 1) * 4 + e, Us, LDUF);
  if (dSu_b != nullptr)
    for (int idx = threadIdx.x; idx < TQ * (LDU / 8); idx += THREADS) reinterpret_cast<uint4*>(dSu_b)[idx] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  // per-entry stage: thread = (row i, 16 consecutive keys)
  const int il = threadIdx.x >> 2, jl0 = (threadIdx.x & 3) * 16, i = i0 + il, Tn = p.Tn;
  const float lse = lse_s[il], Dv = D_s[il];
  uint32_t pk_p[8], pk_s[8];
#pragma unroll
  for (int e2 = 0; e2 < 8; ++e2) {
    float pd2[2], ds2[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int jl = jl0 + 2 * e2 + q, j = j0 + jl;
      const bool dead = j >= Tn || i >= Tn || (p.key_pad != nullptr && p.key_pad[static_cast<int64_t>(b) * Tn + j] != 0);
      float pd = 0.f, ds = 0.f;
      if (!dead) {
        const float s = p.scale * (Ss[il * LDF + jl] + Us[il * LDUF + jl - il + (TQ - 1)]);
        const float pv = __expf(s - lse);
        const float f = dk.factor((static_cast<int64_t>(b * p.H + h) * Tn + i) * Tn + j);
        pd = pv * f;
        ds = pv * (f * DPs[il * LDF + jl] - Dv) * p.scale;
      }
      pd2[q] = pd;
      ds2[q] = ds;
      if (dSu_b != nullptr) dSu_b[il * LDU + jl - il + (TQ - 1)] = __float2bfloat16_rn(ds);
    }
    pk_p[e2] = pack_bf16x2(pd2[0], pd2[1]);
    pk_s[e2] = pack_bf16x2(ds2[0], ds2[1]);
  }
  if (Pd_b != nullptr) {
    *reinterpret_cast<uint4*>(Pd_b + il * LDB + jl0) = make_uint4(pk_p[0], pk_p[1], pk_p[2], pk_p[3]);
    *reinterpret_cast<uint4*>(Pd_b + il * LDB + jl0 + 8) = make_uint4(pk_p[4], pk_p[5], pk_p[6], pk_p[7]);
  }
  *reinterpret_cast<uint4*>(dS_b + il * LDB + jl0) = make_uint4(pk_s[0], pk_s[1], pk_s[2], pk_s[3]);
  *reinterpret_cast<uint4*>(dS_b + il * LDB + jl0 + 8) = make_uint4(pk_s[4], pk_s[5], pk_s[6], pk_s[7]);
  __syncthreads();
}

struct SmemQ {
  static constexpr int TILES_B16 = (5 * TQ + WINR) * LDB + TQ * LDB + TQ * LDU;              // Qc Qp dO K V | Pw | ds | dsu
  static constexpr int F32 = 2 * TQ * LDF + TQ * LDUF + WINR * LDF + 2 * TQ + 2 * DHP;       // S dP | U | dPw accumulator | lse D | red
  static constexpr int BYTES = TILES_B16 * 2 + F32 * 4;
};
struct SmemKV {
  static constexpr int TILES_B16 = (5 * TQ + WINR) * LDB + 2 * TQ * LDB;                      // ... | Pd | ds
  static constexpr int F32 = 2 * TQ * LDF + TQ * LDUF + 2 * TQ;
  static constexpr int BYTES = TILES_B16 * 2 + F32 * 4;
};

// ------------------------------------------------------------------------------------------------ query side
__global__ void __launch_bounds__(THREADS, 1) attn_bwd_q_mma_kernel(MmaArgs p, bf16* __restrict__ dqkv, float* __restrict__ dpp,
                                                                    float* __restrict__ du, float* __restrict__ dv) {
  pdl_wait();
  extern __shared__ __align__(128) uint8_t sm_raw[];
  float* Ss = reinterpret_cast<float*>(sm_raw);
  float* DPs = Ss + TQ * LDF;
  float* Us = DPs + TQ * LDF;
  float* Wacc = Us + TQ * LDUF;                                   // [128][LDF] window gradient of this query tile
  float* lse_s = Wacc + WINR * LDF;
  float* D_s = lse_s + TQ;
  float* red = D_s + TQ;
  bf16* Qc = reinterpret_cast<bf16*>(red + 2 * DHP);
  bf16* Qp = Qc + TQ * LDB;
  bf16* dO = Qp + TQ * LDB;
  bf16* Kt = dO + TQ * LDB;
  bf16* Vt = Kt + TK * LDB;
  bf16* Pw = Vt + TK * LDB;
  bf16* dS_b = Pw + WINR * LDB;
  bf16* dSu_b = dS_b + TQ * LDB;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TQ;
  const int Tn = p.Tn, dh = p.dh, d = p.H * dh;
  const int warp = threadIdx.x >> 5;
  const bf16* qkv = p.qkv + static_cast<int64_t>(b) * Tn * 3 * d + h * dh;
  const bf16* pp = p.pp + h * dh;
  load_tile(Qc, qkv, 3 * d, i0, 0, Tn, TQ, dh, p.u + h * dh);
  load_tile(Qp, qkv, 3 * d, i0, 0, Tn, TQ, dh, p.v + h * dh);
  load_tile_f32(dO, p.dctx + static_cast<int64_t>(b) * Tn * d + h * dh, d, i0, Tn, TQ, dh);
  for (int i = threadIdx.x; i < TQ; i += THREADS) {
    const bool ok = i0 + i < Tn;
    lse_s[i] = ok ? p.lse[(static_cast<int64_t>(b) * p.H + h) * Tn + i0 + i] : 0.f;
    D_s[i] = ok ? p.D[(static_cast<int64_t>(b) * p.H + h) * Tn + i0 + i] : 0.f;
  }
  for (int i = threadIdx.x; i < WINR * LDF; i += THREADS) Wacc[i] = 0.f;
  // dQ accumulators: warp -> row block rb = warp / 2, column blocks (warp & 1) * 2 + {0, 1}; content and positional parts
  FragC dqc[2], dqp[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    frag_zero(dqc[e]);
    frag_zero(dqp[e]);
  }
  const int rb = warp >> 1;
  for (int j0 = 0; j0 < Tn; j0 += TK) {
    __syncthreads();
    const int r0 = (Tn - 1) + j0 - i0 - (TQ - 1);
    load_key_tiles(Kt, Vt, Pw, qkv + d, qkv + 2 * d, 3 * d, j0, Tn, pp, d, r0, 2 * Tn - 1, dh);
    __syncthreads();
    scores_and_grads(p, Qc, Qp, dO, Kt, Vt, Pw, Ss, Us, DPs, lse_s, D_s, nullptr, dS_b, dSu_b, b, h, i0, j0);
    // dQc += ds K   (A = ds [i][j], B[k = j][n = d] = K[j][d]);   dQp += dsu Pw   (K = 128 window rows)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int cb = (warp & 1) * 2 + e;
#pragma unroll
      for (int k = 0; k < TK; k += 16) {
        FragA a;
        FragB bm;
        ld_a(a, dS_b + rb * 16 * LDB + k, LDB);
        ld_b_kn(bm, Kt + k * LDB + cb * 16, LDB);
        frag_mma(dqc[e], a, bm);
      }
#pragma unroll
      for (int k = 0; k < WINR; k += 16) {
        FragA a;
        FragB bm;
        ld_a(a, dSu_b + rb * 16 * LDU + k, LDU);
        ld_b_kn(bm, Pw + k * LDB + cb * 16, LDB);
        frag_mma(dqp[e], a, bm);
      }
    }
    // dPw[w][d] += sum_i dsu[i][w] Qp[i][d]: warp = 16 window rows, all four column blocks; accumulator lives in shared memory
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      FragC c;
      frag_load(c, Wacc + warp * 16 * LDF + cb * 16, LDF);
#pragma unroll
      for (int k = 0; k < TQ; k += 16) {
        FragA a;                                                   // A[m = w][k = i] stored at dsu[i][w]
        FragB bm;
        ld_a_t(a, dSu_b + k * LDU + warp * 16, LDU);
        ld_b_kn(bm, Qp + k * LDB + cb * 16, LDB);
        frag_mma(c, a, bm);
      }
      frag_store(Wacc + warp * 16 * LDF + cb * 16, LDF, c);
    }
    __syncthreads();
    // window rows < 64 are complete: add them to the table gradient; rows >= 64 become rows - 64 of the next window
    const bool last = j0 + TK >= Tn;
    for (int idx = threadIdx.x; idx < (last ? WINR : TK) * (DHP / 4); idx += THREADS) {
      const int w = idx / (DHP / 4), c4 = (idx - w * (DHP / 4)) * 4, r = r0 + w;
      if (r >= 0 && r < 2 * Tn - 1 && c4 < dh) {                   // one 16-byte vector reduction per 4 channels
        float4* dst = reinterpret_cast<float4*>(dpp + static_cast<int64_t>(r) * d + h * dh + c4);
        atomicAdd(dst, *reinterpret_cast<const float4*>(Wacc + w * LDF + c4));
      }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < TK * DHP; idx += THREADS) {
      const int w = idx >> 6, c = idx & 63;
      Wacc[w * LDF + c] = Wacc[(w + TK) * LDF + c];
      Wacc[(w + TK) * LDF + c] = 0.f;
    }
  }
  __syncthreads();
  // dq = dQc + dQp (bf16), du += column sums of dQc, dv += column sums of dQp (valid rows only)
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int cb = (warp & 1) * 2 + e;
    frag_store(Ss + rb * 16 * LDF + cb * 16, LDF, dqc[e]);
    frag_store(DPs + rb * 16 * LDF + cb * 16, LDF, dqp[e]);
  }
  __syncthreads();
  {
    const int il = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 16, i = i0 + il;
    if (i < Tn) {
      bf16* dst = dqkv + (static_cast<int64_t>(b) * Tn + i) * 3 * d + h * dh;
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const float a0 = Ss[il * LDF + c0 + c], a1 = Ss[il * LDF + c0 + c + 1];
        const float b0 = DPs[il * LDF + c0 + c], b1 = DPs[il * LDF + c0 + c + 1];
        if (c0 + c < dh) *reinterpret_cast<uint32_t*>(dst + c0 + c) = pack_bf16x2(a0 + b0, a1 + b1);
      }
    }
  }
  // column sums over the valid rows: one thread per column (64 threads adding into one shared address made the
  // compare-and-swap loop of the shared-memory float atomic spin for a tenth of the kernel)
  if (threadIdx.x < 2 * DHP) {
    const int c = threadIdx.x % DHP;
    const float* src = threadIdx.x < DHP ? Ss : DPs;
    const int nrow = min(TQ, Tn - i0);
    float acc = 0.f;
    for (int r = 0; r < nrow; ++r) acc += src[r * LDF + c];
    if (c < dh) atomicAdd((threadIdx.x < DHP ? du : dv) + h * dh + c, acc);
  }
}

// ------------------------------------------------------------------------------------------------ key side
__global__ void __launch_bounds__(THREADS, 1) attn_bwd_kv_mma_kernel(MmaArgs p, bf16* __restrict__ dqkv) {
  pdl_wait();
  extern __shared__ __align__(128) uint8_t sm_raw[];
  float* Ss = reinterpret_cast<float*>(sm_raw);
  float* DPs = Ss + TQ * LDF;
  float* Us = DPs + TQ * LDF;
  float* lse_s = Us + TQ * LDUF;
  float* D_s = lse_s + TQ;
  bf16* Qc = reinterpret_cast<bf16*>(D_s + TQ);
  bf16* Qp = Qc + TQ * LDB;
  bf16* dO = Qp + TQ * LDB;
  bf16* Kt = dO + TQ * LDB;
  bf16* Vt = Kt + TK * LDB;
  bf16* Pw = Vt + TK * LDB;
  bf16* Pd_b = Pw + WINR * LDB;
  bf16* dS_b = Pd_b + TQ * LDB;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * TK;
  const int Tn = p.Tn, dh = p.dh, d = p.H * dh;
  const int warp = threadIdx.x >> 5, rb = warp >> 1;
  const bf16* qkv = p.qkv + static_cast<int64_t>(b) * Tn * 3 * d + h * dh;
  const bf16* pp = p.pp + h * dh;
  load_tile(Kt, qkv + d, 3 * d, j0, 0, Tn, TK, dh, nullptr);
  load_tile(Vt, qkv + 2 * d, 3 * d, j0, 0, Tn, TK, dh, nullptr);
  FragC dvv[2], dk[2];                                             // rows = keys 16 rb .., column blocks (warp & 1) * 2 + {0, 1}
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    frag_zero(dvv[e]);
    frag_zero(dk[e]);
  }
  for (int i0 = 0; i0 < Tn; i0 += TQ) {
    __syncthreads();
    load_query_tiles(Qc, Qp, dO, Pw, qkv, 3 * d, p.u + h * dh, p.v + h * dh, p.dctx + static_cast<int64_t>(b) * Tn * d + h * dh, d, i0, Tn,
                     pp, d, (Tn - 1) + j0 - i0 - (TQ - 1), 2 * Tn - 1, dh);
    for (int i = threadIdx.x; i < TQ; i += THREADS) {
      const bool ok = i0 + i < Tn;
      lse_s[i] = ok ? p.lse[(static_cast<int64_t>(b) * p.H + h) * Tn + i0 + i] : 0.f;
      D_s[i] = ok ? p.D[(static_cast<int64_t>(b) * p.H + h) * Tn + i0 + i] : 0.f;
    }
    __syncthreads();
    scores_and_grads(p, Qc, Qp, dO, Kt, Vt, Pw, Ss, Us, DPs, lse_s, D_s, Pd_b, dS_b, nullptr, b, h, i0, j0);
    // dV[j][d] += sum_i Pd[i][j] dO[i][d],  dK[j][d] += sum_i ds[i][j] Qc[i][d]    (A[m = j][k = i] stored at [i][j])
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int cb = (warp & 1) * 2 + e;
#pragma unroll
      for (int k = 0; k < TQ; k += 16) {
        FragA a;
        FragB bm;
        ld_a_t(a, Pd_b + k * LDB + rb * 16, LDB);
        ld_b_kn(bm, dO + k * LDB + cb * 16, LDB);
        frag_mma(dvv[e], a, bm);
        ld_a_t(a, dS_b + k * LDB + rb * 16, LDB);
        ld_b_kn(bm, Qc + k * LDB + cb * 16, LDB);
        frag_mma(dk[e], a, bm);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int cb = (warp & 1) * 2 + e;
    frag_store(Ss + rb * 16 * LDF + cb * 16, LDF, dk[e]);
    frag_store(DPs + rb * 16 * LDF + cb * 16, LDF, dvv[e]);
  }
  __syncthreads();
  const int jl = threadIdx.x >> 2, c0 = (threadIdx.x & 3) * 16, j = j0 + jl;
  if (j < Tn) {
    bf16* dst = dqkv + (static_cast<int64_t>(b) * Tn + j) * 3 * d + h * dh;
#pragma unroll
    for (int c = 0; c < 16; c += 2)
      if (c0 + c < dh) {
        *reinterpret_cast<uint32_t*>(dst + d + c0 + c) = pack_bf16x2(Ss[jl * LDF + c0 + c], Ss[jl * LDF + c0 + c + 1]);
        *reinterpret_cast<uint32_t*>(dst + 2 * d + c0 + c) = pack_bf16x2(DPs[jl * LDF + c0 + c], DPs[jl * LDF + c0 + c + 1]);
      }
  }
}

}  // namespace

// D (row term) must already be in `scratch` (attn_train_bwd computes it for both paths)
int attn_train_bwd_mma(const bf16* qkv, const bf16* pp, const float* u, const float* v, const uint8_t* key_pad, const float* lse,
                       const float* dctx, bf16* dqkv, float* dpp, float* du, float* dv, const float* scratch_D, int B, int Tn, int H, int dh,
                       Drop dr, cudaStream_t st) {
  CFM_REQUIRE(dh % 8 == 0 && dh <= DHP, "tensor-core attention backward needs a head size that is a multiple of 8, <= 64");
  static std::atomic<uint64_t> done_q{0}, done_kv{0};
  CFM_TRY(smem_attr_once(done_q, reinterpret_cast<const void*>(attn_bwd_q_mma_kernel), SmemQ::BYTES));
  CFM_TRY(smem_attr_once(done_kv, reinterpret_cast<const void*>(attn_bwd_kv_mma_kernel), SmemKV::BYTES));
  MmaArgs a{qkv, pp, u, v, key_pad, lse, scratch_D, dctx, Tn, H, dh, 1.0f / sqrtf(static_cast<float>(dh)), dr};
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, TQ)), H, B);
  CFM_CUDA_OK(launch_pdl(attn_bwd_q_mma_kernel, dim3(grid), dim3(THREADS), SmemQ::BYTES, st, 1, a, dqkv, dpp, du, dv));
  CFM_CUDA_OK(cudaGetLastError());
  CFM_CUDA_OK(launch_pdl(attn_bwd_kv_mma_kernel, dim3(grid), dim3(THREADS), SmemKV::BYTES, st, 1, a, dqkv));
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
