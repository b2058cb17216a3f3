// Relative-position multi-head self-attention, flash style, bf16 tensor-core MMAs (head size 64).
//
//   score[i,j] = ((q_i+u).k_j + (q_i+v).p[T-1+j-i]) * (1/sqrt(dh))     (attention.py:47-56)
//
// One CTA = 64 query rows of one (batch, head); 4 warps x 16 rows.  Keys are walked in tiles of 64,
// double-buffered with cp.async.  For a (16 x 64) warp tile the positional rows it needs are ONE
// contiguous window of 16+64-1 rows of the projected table, but row ii reads it at offset 15-ii
// (the Transformer-XL shift, attention.py:94-102).  The warp therefore multiplies (q+v) with the
// whole window (10 n-tiles instead of 8), parks the 16x80 result in a private shared-memory patch
// and reads it back along the diagonal - the [T, 2T-1] positional score tensor and its shifted
// copy never exist.  Online softmax in fp32 (exp2 with the scale folded in), P.V on tensor cores.
// Masked keys get weight exactly 0.
#include <math_constants.h>

#include "ops.h"

namespace cfm {
namespace {

constexpr int AQ = 64, AK = 64, DH = 64, NWARP = 4;
constexpr int WIN = AQ + AK;                 // window rows staged per key tile (127 used)
constexpr int SKEW_LD = 88;                  // floats per row of the per-warp skew patch (80 used)
constexpr int TILE_B = AK * DH * 2;          // 8 KB
constexpr int WIN_B = WIN * DH * 2;          // 16 KB
constexpr int STAGE_B = 2 * TILE_B + WIN_B;  // K, V, window
constexpr int SMEM_B = 2 * STAGE_B + NWARP * 16 * SKEW_LD * 4;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int bytes = valid ? 16 : 0;          // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` in a [rows][128 B] tile, XOR-swizzled
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(NWARP * 32, 2)
rel_attention_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ pos, int ldp,
                         const float* __restrict__ ub, const float* __restrict__ vb,
                         const uint8_t* __restrict__ key_pad, bf16* __restrict__ ctx, int Tn, int H,
                         float scale_log2e) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, c = lane & 3;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * AQ;
  const int d = H * DH;
  const int64_t row_stride = 3 * static_cast<int64_t>(d);
  const bf16* base = qkv + static_cast<int64_t>(b) * Tn * row_stride + h * DH;
  const bf16* pbase = pos + h * DH;
  const uint32_t smem_base = smem_u32(smem);
  float* skew = reinterpret_cast<float*>(smem + 2 * STAGE_B) + warp * 16 * SKEW_LD;
  const int ntiles = (Tn + AK - 1) / AK;

  auto load_stage = [&](int stage, int jt) {
    const int j0 = jt * AK;
    const uint32_t sK = smem_base + stage * STAGE_B, sV = sK + TILE_B, sP = sV + TILE_B;
    for (int idx = tid; idx < AK * 8; idx += NWARP * 32) {          // K and V: 64 rows x 8 chunks
      const int r = idx >> 3, ch = idx & 7;
      const bool ok = (j0 + r) < Tn;
      const bf16* src = base + static_cast<int64_t>(ok ? j0 + r : 0) * row_stride + ch * 8;
      cp_async16(sK + swz(r, ch), src + d, ok);
      cp_async16(sV + swz(r, ch), src + 2 * d, ok);
    }
    const int r0 = Tn - 1 + j0 - i0 - (AQ - 1);                     // table row of window entry 0
    for (int idx = tid; idx < WIN * 8; idx += NWARP * 32) {
      const int w = idx >> 3, ch = idx & 7;
      const int r = r0 + w;
      const bool ok = (r >= 0) && (r <= 2 * Tn - 2);
      cp_async16(sP + swz(w, ch), pbase + static_cast<int64_t>(ok ? r : 0) * ldp + ch * 8, ok);
    }
  };

  load_stage(0, 0);
  cp_async_commit();

  // ---- Q fragments (A operand, 16 rows x 64) with the two biases folded in, bf16 like the reference's autocast
  uint32_t qu[4][4], qv[4][4];
  {
    const int ra = i0 + warp * 16 + g, rb = ra + 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int col = ks * 16 + half * 8 + 2 * c;
        const float u0 = ub[h * DH + col], u1 = ub[h * DH + col + 1];
        const float v0 = vb[h * DH + col], v1 = vb[h * DH + col + 1];
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
        if (ra < Tn) {
          const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(base + ra * row_stride + col);
          a0 = __low2float(t); a1 = __high2float(t);
        }
        if (rb < Tn) {
          const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(base + rb * row_stride + col);
          b0 = __low2float(t); b1 = __high2float(t);
        }
        // the softmax scale (1/sqrt(dh), in log2 units) is folded into the query operands
        qu[ks][half * 2 + 0] = pack_bf16x2((a0 + u0) * scale_log2e, (a1 + u1) * scale_log2e);
        qu[ks][half * 2 + 1] = pack_bf16x2((b0 + u0) * scale_log2e, (b1 + u1) * scale_log2e);
        qv[ks][half * 2 + 0] = pack_bf16x2((a0 + v0) * scale_log2e, (a1 + v1) * scale_log2e);
        qv[ks][half * 2 + 1] = pack_bf16x2((b0 + v0) * scale_log2e, (b1 + v1) * scale_log2e);
      }
    }
  }

  float o[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m_run[2] = {-CUDART_INF_F, -CUDART_INF_F}, l_run[2] = {0.f, 0.f};
  const int wbase = (AQ - 16) - 16 * warp;        // first window row this warp's 16 rows can touch

  for (int jt = 0; jt < ntiles; ++jt) {
    const int stage = jt & 1;
    if (jt + 1 < ntiles) {
      load_stage(stage ^ 1, jt + 1);              // buffer was released by the barrier at the end of jt-1
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t sK = smem_base + stage * STAGE_B, sV = sK + TILE_B, sP = sV + TILE_B;
    const int j0 = jt * AK;

    // ---- content scores: S = (q+u) K^T   (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sK + swz(n * 8 + (lane & 7), kp * 4 + (lane >> 3)), b0, b1, b2, b3);
        mma_bf16(s[n], qu[kp * 2], b0, b1);
        mma_bf16(s[n], qu[kp * 2 + 1], b2, b3);
      }
    }
    // ---- positional scores against the warp's window (16 x 80), parked in the skew patch
#pragma unroll
    for (int n = 0; n < 10; ++n) {
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sP + swz(wbase + n * 8 + (lane & 7), kp * 4 + (lane >> 3)), b0, b1, b2, b3);
        mma_bf16(ps, qv[kp * 2], b0, b1);
        mma_bf16(ps, qv[kp * 2 + 1], b2, b3);
      }
      *reinterpret_cast<float2*>(&skew[g * SKEW_LD + n * 8 + 2 * c]) = make_float2(ps[0], ps[1]);
      *reinterpret_cast<float2*>(&skew[(g + 8) * SKEW_LD + n * 8 + 2 * c]) = make_float2(ps[2], ps[3]);
    }
    __syncwarp();
    // ---- usable-key bitmask of this tile (warp-uniform; all ones on the common fully-valid tile)
    uint32_t vm[2] = {0xffffffffu, 0xffffffffu};
    bool need_mask = false;
    if (j0 + AK > Tn || key_pad != nullptr) {
      const int ja = j0 + lane, jb = ja + 32;
      const bool va = ja < Tn && (key_pad == nullptr || key_pad[static_cast<int64_t>(b) * Tn + ja] == 0);
      const bool vb = jb < Tn && (key_pad == nullptr || key_pad[static_cast<int64_t>(b) * Tn + jb] == 0);
      vm[0] = __ballot_sync(0xffffffffu, va);
      vm[1] = __ballot_sync(0xffffffffu, vb);
      need_mask = (vm[0] & vm[1]) != 0xffffffffu;
    }
    // ---- add the shifted positional term, mask, online softmax (scores are already in log2 units)
    float tmax[2] = {-CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int rl = g + (e >> 1) * 8;                      // local row 0..15
        const int jl = n * 8 + 2 * c + (e & 1);               // key within the tile
        float v = s[n][e] + skew[rl * SKEW_LD + jl - rl + 15];
        if (need_mask && ((vm[n >> 2] >> (jl & 31)) & 1u) == 0) v = -CUDART_INF_F;
        s[n][e] = v;
        tmax[e >> 1] = fmaxf(tmax[e >> 1], v);
      }
    }
    __syncwarp();                                             // skew patch is rewritten next tile
    float alpha[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 1));
      tmax[r] = fmaxf(tmax[r], __shfl_xor_sync(0xffffffffu, tmax[r], 2));
      const float m_new = fmaxf(m_run[r], tmax[r]);
      m_use[r] = (m_new == -CUDART_INF_F) ? 0.f : m_new;
      alpha[r] = ex2_approx(m_run[r] - m_use[r]);
      m_run[r] = m_new;
    }
    float psum[2] = {0.f, 0.f};
    uint32_t pa[4][4];                                        // P as A fragments for the 4 key k-steps
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float p0 = ex2_approx(s[n][0] - m_use[0]), p1 = ex2_approx(s[n][1] - m_use[0]);
      const float p2 = ex2_approx(s[n][2] - m_use[1]), p3 = ex2_approx(s[n][3] - m_use[1]);
      psum[0] += p0 + p1;
      psum[1] += p2 + p3;
      pa[n >> 1][(n & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      psum[r] += __shfl_xor_sync(0xffffffffu, psum[r], 1);
      psum[r] += __shfl_xor_sync(0xffffffffu, psum[r], 2);
      l_run[r] = l_run[r] * alpha[r] + psum[r];
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      o[n][0] *= alpha[0]; o[n][1] *= alpha[0];
      o[n][2] *= alpha[1]; o[n][3] *= alpha[1];
    }
    // ---- O += P V   (k = keys, n = head dim; V is [key][dh] so B comes through ldmatrix.trans)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sV + swz(ks * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), np * 2 + (lane >> 4)), b0, b1, b2, b3);
        mma_bf16(o[np * 2], pa[ks], b0, b1);
        mma_bf16(o[np * 2 + 1], pa[ks], b2, b3);
      }
    }
    __syncthreads();                                          // everyone done with this stage's buffers
  }

  // ---- normalise and store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = i0 + warp * 16 + g + r * 8;
    if (row >= Tn) continue;
    const float inv = (l_run[r] > 0.f) ? 1.f / l_run[r] : 0.f;
    bf16* orow = ctx + (static_cast<int64_t>(b) * Tn + row) * d + h * DH;
#pragma unroll
    for (int n = 0; n < 8; ++n)
      *reinterpret_cast<uint32_t*>(orow + n * 8 + 2 * c) = pack_bf16x2(o[n][r * 2] * inv, o[n][r * 2 + 1] * inv);
  }
}

}  // namespace

int rel_attention_bf16_mma(const bf16* qkv, const bf16* pos, int ldp, const float* u, const float* v,
                           const uint8_t* key_pad, bf16* ctx, int B, int Tn, int H, int dh, cudaStream_t stream) {
  CFM_REQUIRE(dh == DH, "tensor-core attention is specialised for head size 64");
  CFM_REQUIRE(B > 0 && Tn > 0 && H > 0 && H <= 65535 && B <= 65535, "bad attention shape");
  CFM_REQUIRE(ldp % 8 == 0 && (H * dh) % 8 == 0, "attention operands must be 16-byte aligned per row");
  static bool attr_set = false;
  if (!attr_set) {
    CFM_CUDA_OK(cudaFuncSetAttribute(rel_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_B));
    attr_set = true;
  }
  dim3 grid(static_cast<unsigned>(ceil_div(Tn, AQ)), static_cast<unsigned>(H), static_cast<unsigned>(B));
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(dh));
  rel_attention_mma_kernel<<<grid, NWARP * 32, SMEM_B, stream>>>(qkv, pos, ldp, u, v, key_pad, ctx, Tn, H, scale_log2e);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
