// CTC head of the reference on the device: the step right AFTER the encoder path.
//
//   Decoder.forward (model/modules/decoder.py:15-27):
//     pack_padded_sequence -> LSTM(d_model -> hidden, n_layers) -> pad_packed_sequence -> Swish -> BatchNorm1d (eval)
//     -> Linear(hidden -> vocab)
//
// Packing semantics kept: a sample's state stops advancing at its length and the frames at or beyond it are ZERO
// before the Swish (pad_packed_sequence pads with zeros), so the logits there are the constant Linear(BatchNorm(0)).
//
// Mapping.  The input projection of every layer (x W_ih^T + b_ih + b_hh for all frames at once) and the vocabulary
// projection are dense GEMMs (tcgen05 in bf16 mode, CUDA cores in fp32 mode; eval-mode BatchNorm is folded into the
// vocabulary projection at pack time).  The recurrence is sequential in time, so it is ONE persistent cooperative
// kernel per layer: a CTA owns 5 hidden units = 20 gate rows of W_hh, held in REGISTERS for all T steps (warp = 2 rows,
// lane = a contiguous k-slice); per step it stages h_{t-1} (fp32, from L2) in shared memory 32 batch rows at a time,
// forms the 20 x 32 gate pre-activations (5 LDS.128 + 40 FMA per batch row and thread, then a butterfly that reduces
// all 64 partial sums of a thread in 62 shuffles), applies the gate non-linearities to its 5 x 32 cells, writes h_t and
// the layer output, and meets the other CTAs at a grid barrier.  All recurrent arithmetic is fp32.
#include <cooperative_groups.h>

#include "ops.h"

namespace cg = cooperative_groups;

namespace cfm {
namespace {

constexpr int LSTM_U = 5, LSTM_ROWS = 4 * LSTM_U, LSTM_WARPS = LSTM_ROWS / 2, LSTM_THREADS = LSTM_WARPS * 32;
constexpr int LSTM_BCH = 32;                               // batch rows staged per pass

struct LstmParams {
  const float* xin;        // [B, T, 4H]: x W_ih^T + b_ih + b_hh, gate order i | f | g | o (nn.LSTM)
  const float* w_hh;       // [4H, H]
  const int64_t* lengths;  // [B] or null
  float* h_state;          // [2][B][H] ping-pong, zero-initialised
  float* c_state;          // [B][H], zero-initialised
  void* out;               // [B, T, H] TOut
  int B, T, H;
  int last_layer;          // 1: out = swish(h) (decoder.py:22); 0: out = h (input of the next LSTM layer)
};

template <int KI, typename TOut>
__global__ void __launch_bounds__(LSTM_THREADS, 1) lstm_layer_kernel(LstmParams p) {
  extern __shared__ float lstm_smem[];
  constexpr int HP = 32 * KI;                              // padded row length: lane owns k in [lane*KI, lane*KI + KI)
  float* hs = lstm_smem;                                   // [LSTM_BCH][HP]
  float* gates_s = hs + LSTM_BCH * HP;                     // [LSTM_BCH][LSTM_ROWS]
  cg::grid_group grid = cg::this_grid();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = p.H, B = p.B, T = p.T;
  const int unit0 = blockIdx.x * LSTM_U;

  // recurrent weights of this warp's two gate rows, this lane's k-slice: registers for the whole sequence
  float w[2][KI];
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int r = warp * 2 + rr, g = r / LSTM_U, unit = unit0 + r % LSTM_U;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      const int k = lane * KI + i;
      w[rr][i] = (unit < H && k < H) ? __ldg(p.w_hh + (static_cast<int64_t>(g) * H + unit) * H + k) : 0.f;
    }
  }
  TOut* out = static_cast<TOut*>(p.out);

  for (int t = 0; t < T; ++t) {
    const float* h_prev = p.h_state + static_cast<int64_t>(t & 1) * B * H;
    float* h_next = p.h_state + static_cast<int64_t>((t + 1) & 1) * B * H;
    for (int b0 = 0; b0 < B; b0 += LSTM_BCH) {
      const int nb = min(LSTM_BCH, B - b0);
      // ---- stage h_{t-1} of these batch rows (zero padded to HP / LSTM_BCH rows)
      for (int i = threadIdx.x; i < LSTM_BCH * HP; i += LSTM_THREADS) {
        const int bl = i / HP, k = i - bl * HP;
        hs[i] = (bl < nb && k < H) ? __ldcg(h_prev + static_cast<int64_t>(b0 + bl) * H + k) : 0.f;
      }
      __syncthreads();
      // ---- partial dot products: v[2*bl + rr] = sum over this lane's k-slice
      float v[2 * LSTM_BCH];
#pragma unroll
      for (int bl = 0; bl < LSTM_BCH; ++bl) {
        const float4* hr = reinterpret_cast<const float4*>(hs + bl * HP + lane * KI);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i4 = 0; i4 < KI / 4; ++i4) {
          const float4 h4 = hr[i4];
          a0 = fmaf(w[0][4 * i4 + 0], h4.x, a0); a1 = fmaf(w[1][4 * i4 + 0], h4.x, a1);
          a0 = fmaf(w[0][4 * i4 + 1], h4.y, a0); a1 = fmaf(w[1][4 * i4 + 1], h4.y, a1);
          a0 = fmaf(w[0][4 * i4 + 2], h4.z, a0); a1 = fmaf(w[1][4 * i4 + 2], h4.z, a1);
          a0 = fmaf(w[0][4 * i4 + 3], h4.w, a0); a1 = fmaf(w[1][4 * i4 + 3], h4.w, a1);
        }
        v[2 * bl] = a0;
        v[2 * bl + 1] = a1;
      }
      // ---- butterfly: after the five exchanges lane L holds the two complete sums of batch row L
#pragma unroll
      for (int s = 16, n = 2 * LSTM_BCH; s >= 1; s >>= 1, n >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < n / 2; ++j) {
          const float send = upper ? v[j] : v[j + n / 2];
          const float keep = upper ? v[j + n / 2] : v[j];
          v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
      }
      gates_s[lane * LSTM_ROWS + warp * 2] = v[0];
      gates_s[lane * LSTM_ROWS + warp * 2 + 1] = v[1];
      __syncthreads();
      // ---- cells: thread = (batch row, unit)
      if (threadIdx.x < LSTM_BCH * LSTM_U) {
        const int bl = threadIdx.x & 31, ul = threadIdx.x >> 5, b = b0 + bl, unit = unit0 + ul;
        if (bl < nb && unit < H) {
          const float* xr = p.xin + (static_cast<int64_t>(b) * T + t) * 4 * H + unit;
          const float gi = gates_s[bl * LSTM_ROWS + ul] + xr[0];
          const float gf = gates_s[bl * LSTM_ROWS + LSTM_U + ul] + xr[H];
          const float gg = gates_s[bl * LSTM_ROWS + 2 * LSTM_U + ul] + xr[2 * H];
          const float go = gates_s[bl * LSTM_ROWS + 3 * LSTM_U + ul] + xr[3 * H];
          const int64_t si = static_cast<int64_t>(b) * H + unit;
          const bool live = p.lengths == nullptr || t < p.lengths[b];
          float hn = 0.f;
          if (live) {
            const float c = sigmoid_f(gf) * p.c_state[si] + sigmoid_f(gi) * tanhf(gg);
            hn = sigmoid_f(go) * tanhf(c);
            p.c_state[si] = c;
            h_next[si] = hn;
          } else {
            h_next[si] = __ldcg(h_prev + si);                  // the packed sequence has ended: state frozen
          }
          const float o = !live ? 0.f : (p.last_layer ? swish_f(hn) : hn);
          out[(static_cast<int64_t>(b) * T + t) * H + unit] = from_f32<TOut>(o);
        }
      }
      __syncthreads();                                         // hs / gates_s are reused by the next pass
    }
    __threadfence();
    grid.sync();                                               // every h_t element is visible before step t+1
  }
}

template <int KI, typename TOut>
int launch_lstm(LstmParams p, cudaStream_t stream) {
  auto kern = lstm_layer_kernel<KI, TOut>;
  const int smem = (LSTM_BCH * 32 * KI + LSTM_BCH * LSTM_ROWS) * 4;
  static std::atomic<uint64_t> attr_done{0};
  CFM_TRY(smem_attr_once(attr_done, reinterpret_cast<const void*>(kern), smem));
  void* args[] = {&p};
  const int grid = static_cast<int>(ceil_div(p.H, LSTM_U));
  CFM_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(LSTM_THREADS), args, smem, stream));
  return 0;
}

template <typename TOut>
int lstm_dispatch(const LstmParams& p, cudaStream_t stream) {
  const int ki = static_cast<int>(ceil_div(p.H, 32));
  if (ki <= 4) return launch_lstm<4, TOut>(p, stream);
  if (ki <= 8) return launch_lstm<8, TOut>(p, stream);
  if (ki <= 12) return launch_lstm<12, TOut>(p, stream);
  if (ki <= 16) return launch_lstm<16, TOut>(p, stream);
  if (ki <= 20) return launch_lstm<20, TOut>(p, stream);
  return launch_lstm<24, TOut>(p, stream);
}

}  // namespace

template <typename TOut>
int lstm_layer_forward(const float* xin, const float* w_hh, const int64_t* lengths, float* h_state, float* c_state, TOut* out,
                       int B, int T, int H, bool last_layer, int num_sms, cudaStream_t stream) {
  CFM_REQUIRE(B > 0 && T > 0 && H > 0, "empty LSTM");
  CFM_REQUIRE(ceil_div(H, LSTM_U) <= num_sms && H <= 24 * 32,
              "the persistent LSTM kernel keeps 5 hidden units per SM: lstm_hidden_dim up to min(5 * #SMs, 768)");
  CFM_CUDA_OK(cudaMemsetAsync(h_state, 0, static_cast<size_t>(2) * B * H * 4, stream));
  CFM_CUDA_OK(cudaMemsetAsync(c_state, 0, static_cast<size_t>(B) * H * 4, stream));
  LstmParams p{xin, w_hh, lengths, h_state, c_state, out, B, T, H, last_layer ? 1 : 0};
  return lstm_dispatch<TOut>(p, stream);
}
template int lstm_layer_forward<float>(const float*, const float*, const int64_t*, float*, float*, float*, int, int, int, bool, int, cudaStream_t);
template int lstm_layer_forward<bf16>(const float*, const float*, const int64_t*, float*, float*, bf16*, int, int, int, bool, int, cudaStream_t);

// fold eval-mode BatchNorm1d (decoder.py:25) into the vocabulary projection that follows it (decoder.py:27):
//   W'[v,h] = W[v,h] s[h],  b'[v] = b[v] + sum_h W[v,h] (beta[h] - mean[h] s[h]),  s = gamma / sqrt(var + eps)
__global__ void fold_bn_linear_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                                      float eps, float* __restrict__ w_out, float* __restrict__ b_out, int V, int Vp, int H) {
  const int v = blockIdx.x;                                  // one block per (padded) output row
  float acc = 0.f;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float wv = 0.f;
    if (v < V) {
      const float s = gamma[h] / sqrtf(var[h] + eps);
      wv = w[static_cast<int64_t>(v) * H + h];
      acc += wv * (beta[h] - mean[h] * s);
      wv *= s;
    }
    w_out[static_cast<int64_t>(v) * H + h] = wv;
  }
  __shared__ float red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
    b_out[v] = v < V ? b[v] + t : 0.f;
  }
  (void)Vp;
}
int fold_bn_linear(const float* w, const float* b, const float* gamma, const float* beta, const float* mean, const float* var,
                   float eps, float* w_out, float* b_out, int V, int Vp, int H, cudaStream_t stream) {
  fold_bn_linear_kernel<<<Vp, 256, 0, stream>>>(w, b, gamma, beta, mean, var, eps, w_out, b_out, V, Vp, H);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}
__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = a[i] + b[i];
}
int add_vec(const float* a, const float* b, float* o, int n, cudaStream_t stream) {
  add_vec_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a, b, o, n);
  CFM_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace cfm
