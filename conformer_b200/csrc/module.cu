// Module handles: weight registration/packing and the forward orchestration of every reference
// module on the encoder path (Encoder, ConformerBlock, FeedForwardModule, MultiHeadSelfAttentionModule,
// RelativeMultiHeadAttention, ConvolutionModule, ConvolutionSubsampling).
#include "module.h"

#include <cmath>
#include <cstring>

namespace cfm {

// =================================================================================== arena
int Arena::reserve(size_t bytes) {
  release();
  CFM_CUDA_OK(cudaMalloc(&base_, bytes));
  cap_ = bytes;
  used_ = 0;
  return 0;
}
void Arena::release() {
  if (base_) cudaFree(base_);
  base_ = nullptr;
  cap_ = used_ = 0;
}
void* Arena::take(size_t bytes) {
  const size_t a = (used_ + 255) & ~static_cast<size_t>(255);
  if (a + bytes > cap_) return nullptr;
  used_ = a + bytes;
  return static_cast<char*>(base_) + a;
}

// =================================================================================== packing
namespace {

struct Packer {
  Module& m;
  cudaStream_t stream;
  bool bf16_mode;
  int status = 0;

  const float* param(const std::string& name, int64_t numel) {
    auto it = m.params.find(name);
    if (it == m.params.end()) {
      if (!status) set_error("missing parameter '" + name + "' (expected " + std::to_string(numel) + " elements)");
      status = 1;
      return nullptr;
    }
    if (it->second.numel != numel) {
      if (!status)
        set_error("parameter '" + name + "' has " + std::to_string(it->second.numel) + " elements, expected " +
                  std::to_string(numel));
      status = 1;
      return nullptr;
    }
    it->second.used = true;
    return it->second.data;
  }
  float* alloc_f32(int64_t n) {
    float* p = static_cast<float*>(m.arena.take(static_cast<size_t>(n) * 4));
    if (!p && !status) { set_error("weight arena exhausted"); status = 4; }
    return p;
  }
  bf16* alloc_bf16(int64_t n) {
    bf16* p = static_cast<bf16*>(m.arena.take(static_cast<size_t>(n) * 2));
    if (!p && !status) { set_error("weight arena exhausted"); status = 4; }
    return p;
  }
  // copy a vector parameter
  const float* vec(const std::string& name, int64_t n) {
    const float* src = param(name, n);
    float* dst = alloc_f32(n);
    if (status) return nullptr;
    if (cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, stream) != cudaSuccess) {
      set_error("cudaMemcpyAsync failed while packing " + name);
      status = 2;
    }
    return dst;
  }
  void finish_linear(Linear& L, float* w32, float* b, int N, int K) {
    L.w32 = w32; L.b = b; L.N = N; L.K = K; L.w16 = nullptr;
    if (bf16_mode && !status) {
      L.w16 = alloc_bf16(static_cast<int64_t>(N) * K);
      if (!status && cast_f32_to_bf16(w32, L.w16, static_cast<int64_t>(N) * K, stream)) status = 2;
    }
  }
  // nn.Linear / 1x1 Conv1d: weight [N,K(,1)], bias [N]
  void linear(Linear& L, const std::string& pre, int N, int K) {
    float* w = const_cast<float*>(vec(pre + ".weight", static_cast<int64_t>(N) * K));
    float* b = const_cast<float*>(vec(pre + ".bias", N));
    finish_linear(L, w, b, N, K);
  }
  void ffn(FfnW& w, const std::string& pre, int d) {
    w.ln_g = vec(pre + "layer_norm.weight", d);
    w.ln_b = vec(pre + "layer_norm.bias", d);
    linear(w.hidden, pre + "hidden_linear", 4 * d, d);
    linear(w.out, pre + "out_linear", d, 4 * d);
  }
  void rel_attn(AttnW& w, const std::string& pre, int d, int H) {
    w.u = vec(pre + "content_bias", d);
    w.v = vec(pre + "position_bias", d);
    // fused q|k|v projection: rows [0,d) query, [d,2d) key, [2d,3d) value (attention.py:78-80)
    float* wq = alloc_f32(3LL * d * d);
    float* bq = alloc_f32(3LL * d);
    const char* names[3] = {"query_proj", "key_proj", "value_proj"};
    for (int i = 0; i < 3 && !status; ++i) {
      const float* sw = param(pre + names[i] + ".weight", static_cast<int64_t>(d) * d);
      const float* sb = param(pre + names[i] + ".bias", d);
      if (status) break;
      cudaMemcpyAsync(wq + static_cast<int64_t>(i) * d * d, sw, static_cast<size_t>(d) * d * 4, cudaMemcpyDeviceToDevice, stream);
      cudaMemcpyAsync(bq + static_cast<int64_t>(i) * d, sb, static_cast<size_t>(d) * 4, cudaMemcpyDeviceToDevice, stream);
    }
    finish_linear(w.qkv, wq, bq, 3 * d, d);
    linear(w.pos, pre + "pos_proj", d, d);
    linear(w.out, pre + "out_proj", d, d);
    w.wpT = nullptr;
    if (bf16_mode && !status && rel_attention_fact_supported(d, d / H)) {
      w.wpT = alloc_bf16(static_cast<int64_t>(d) * d);
      if (!status && pack_pos_fact(w.pos.w32, w.wpT, d, d / H, stream)) status = 2;
      // second copy of the fused projection with content bias and softmax scale folded into the query rows
      float* wq2 = alloc_f32(3LL * d * d);
      float* bq2 = alloc_f32(3LL * d);
      float* pcst = alloc_f32(static_cast<int64_t>(H) * d);
      if (!status) {
        cudaMemcpyAsync(wq2, wq, static_cast<size_t>(3) * d * d * 4, cudaMemcpyDeviceToDevice, stream);
        cudaMemcpyAsync(bq2, bq, static_cast<size_t>(3) * d * 4, cudaMemcpyDeviceToDevice, stream);
        if (pack_q_fact(wq2, bq2, w.u, d, d / H, stream) || pos_const_fact(w.pos.w32, w.u, w.v, pcst, d, d / H, stream)) status = 2;
      }
      finish_linear(w.qkv_fact, wq2, bq2, 3 * d, d);
      w.pos_const = pcst;
    }
  }
  void mhsa(AttnW& w, const std::string& pre, int d, int H) {
    w.ln_g = vec(pre + "layer_norm.weight", d);
    w.ln_b = vec(pre + "layer_norm.bias", d);
    rel_attn(w, pre + "attention.", d, H);
  }
  void conv(ConvW& w, const std::string& pre, int d, int k) {
    w.ln_g = vec(pre + "layer_norm.weight", d);
    w.ln_b = vec(pre + "layer_norm.bias", d);
    w.glu_fused = bf16_mode && d % 64 == 0 && k == 31;
    if (w.glu_fused) {
      // value/gate rows interleaved per 32 channels so that the GEMM epilogue can apply GLU (activation.py:16-17)
      const float* w1 = param(pre + "pointwise_conv_1.weight", 2LL * d * d);
      const float* b1 = param(pre + "pointwise_conv_1.bias", 2LL * d);
      float* w1p = alloc_f32(2LL * d * d);
      float* b1p = alloc_f32(2LL * d);
      if (!status && (pack_glu_rows(w1, w1p, d, d, stream) || pack_glu_rows(b1, b1p, d, 1, stream))) status = 2;
      finish_linear(w.pw1, w1p, b1p, 2 * d, d);
    } else {
      linear(w.pw1, pre + "pointwise_conv_1", 2 * d, d);
    }
    linear(w.pw2, pre + "pointwise_conv_2", d, d);
    const float* dw = param(pre + "deepwise_conv.weight", static_cast<int64_t>(d) * k);
    const float* db = param(pre + "deepwise_conv.bias", d);
    const float* bw = param(pre + "batch_norm.weight", d);
    const float* bb = param(pre + "batch_norm.bias", d);
    const float* rm = param(pre + "batch_norm.running_mean", d);
    const float* rv = param(pre + "batch_norm.running_var", d);
    float* wo = alloc_f32(static_cast<int64_t>(d) * k);
    float* bo = alloc_f32(d);
    if (!status && fold_dwconv_bn(dw, db, bw, bb, rm, rv, 1e-5f, wo, bo, d, k, stream)) status = 2;
    w.dw_w = wo; w.dw_b = bo; w.ksize = k;
  }
  void block(BlockW& w, const std::string& pre, int d, int H, int k) {
    ffn(w.ffn1, pre + "ffn_1.", d);
    mhsa(w.attn, pre + "attention.", d, H);
    conv(w.conv, pre + "conv.", d, k);
    ffn(w.ffn2, pre + "ffn_2.", d);
    w.ln_g = vec(pre + "layer_norm.weight", d);
    w.ln_b = vec(pre + "layer_norm.bias", d);
  }
  // CTC head (decoder.py:10-13): nn.LSTM parameter names weight_ih_l{k}, weight_hh_l{k}, bias_ih_l{k}, bias_hh_l{k}
  void decoder(DecW& w, int d, int V, int H, int L) {
    w.V = V; w.H = H; w.Vp = (V + 7) & ~7;
    w.layers.resize(L);
    for (int l = 0; l < L && !status; ++l) {
      const int in = l == 0 ? d : H;
      const std::string sfx = "_l" + std::to_string(l);
      float* wih = const_cast<float*>(vec("lstm.weight_ih" + sfx, 4LL * H * in));
      const float* bih = param("lstm.bias_ih" + sfx, 4LL * H);
      const float* bhh = param("lstm.bias_hh" + sfx, 4LL * H);
      float* bsum = alloc_f32(4LL * H);
      if (!status && add_vec(bih, bhh, bsum, 4 * H, stream)) status = 2;
      w.layers[l].w_hh = vec("lstm.weight_hh" + sfx, 4LL * H * H);
      if (bf16_mode && lstm_tc_supported(H)) {
        // tensor-core recurrence: gate rows of W_ih, the bias and W_hh in the row order of lstm_tc.cu
        float* wih_p = alloc_f32(4LL * H * in);
        float* bsum_p = alloc_f32(4LL * H);
        w.layers[l].w_hh16p = alloc_bf16(4LL * H * H);
        if (!status && (lstm_pack_rows_f32(wih, wih_p, H, in, stream) || lstm_pack_rows_f32(bsum, bsum_p, H, 1, stream) ||
                        lstm_pack_rows_bf16(w.layers[l].w_hh, w.layers[l].w_hh16p, H, H, stream)))
          status = 2;
        finish_linear(w.layers[l].ih, wih_p, bsum_p, 4 * H, in);
        continue;
      }
      finish_linear(w.layers[l].ih, wih, bsum, 4 * H, in);
    }
    const float* lw = param("linear.weight", static_cast<int64_t>(V) * H);
    const float* lb = param("linear.bias", V);
    const float* g = param("norm.weight", H);
    const float* be = param("norm.bias", H);
    const float* rm = param("norm.running_mean", H);
    const float* rv = param("norm.running_var", H);
    float* wo = alloc_f32(static_cast<int64_t>(w.Vp) * H);
    float* bo = alloc_f32(w.Vp);
    if (!status && fold_bn_linear(lw, lb, g, be, rm, rv, 1e-5f, wo, bo, V, w.Vp, H, stream)) status = 2;
    finish_linear(w.out, wo, bo, w.Vp, H);
  }
  void front(FrontW& w, const std::string& pre, int d) {
    w.c1_w = vec(pre + "conv_1.weight", static_cast<int64_t>(d) * 9);
    w.c1_b = vec(pre + "conv_1.bias", d);
    w.c1_w16 = nullptr;
    if (bf16_mode && !status && conv1_tc_supported(d)) {
      w.c1_w16 = alloc_bf16(static_cast<int64_t>(d) * 64);
      if (!status && pack_conv1_weight(w.c1_w, w.c1_w16, d, stream)) status = 2;
    }
    const float* w2 = param(pre + "conv_2.weight", static_cast<int64_t>(d) * d * 9);
    float* w2p = alloc_f32(static_cast<int64_t>(d) * d * 9);
    float* b2 = const_cast<float*>(vec(pre + "conv_2.bias", d));
    if (!status && pack_conv2_weight(w2, w2p, d, stream)) status = 2;
    finish_linear(w.conv2, w2p, b2, d, 9 * d);
  }
};

size_t arena_bytes_for(const Module& m) {
  size_t total = 0;
  for (const auto& kv : m.params) total += static_cast<size_t>(kv.second.numel);
  // fp32 copy + bf16 copy + per-tensor 256-byte alignment slack + the fused pos_proj copy
  const size_t d = m.cfg.d_model;
  // (the CTC head keeps a second, row-packed copy of its LSTM matrices for the tensor-core recurrence)
  const size_t per_elem = m.kind == CFM_KIND_DECODER ? 14 : 6;
  return total * per_elem + (m.params.size() * 4 + 64) * 256 + static_cast<size_t>(m.cfg.n_blocks + 1) * d * d * 32 + (1 << 16);
}

}  // namespace

int Module::finalize(int compute_mode, cudaStream_t stream) {
  CFM_REQUIRE(compute_mode == 0 || compute_mode == 1, "compute must be CFM_COMPUTE_F32 or CFM_COMPUTE_BF16");
  const int d = cfg.d_model, H = cfg.n_heads, k = cfg.kernel_size;
  CFM_REQUIRE(d > 0 && d % 4 == 0, "d_model must be a positive multiple of 4");
  CFM_REQUIRE(kind == CFM_KIND_FFN || kind == CFM_KIND_CONV_MODULE || kind == CFM_KIND_SUBSAMPLING || kind == CFM_KIND_DECODER ||
                  (H > 0 && d % H == 0),
              "d_model must be divisible by n_heads");
  int dev = 0;
  CFM_CUDA_OK(cudaGetDevice(&dev));
  CFM_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  bf16_mode = compute_mode == 1;
  if (bf16_mode) {
    CFM_TRY(tc_init());
    CFM_REQUIRE(d % 8 == 0, "bf16 mode needs d_model to be a multiple of 8");
    if (kind == CFM_KIND_ENCODER || kind == CFM_KIND_SUBSAMPLING)
      CFM_REQUIRE(d % 64 == 0, "bf16 mode needs d_model to be a multiple of 64 for the implicit-GEMM front-end; use fp32");
  }
  for (auto& kv : params) kv.second.used = false;
  CFM_TRY(arena.reserve(arena_bytes_for(*this)));
  Packer P{*this, stream, bf16_mode};
  blocks.clear();
  switch (kind) {
    case CFM_KIND_FFN: P.ffn(ffn, "", d); break;
    case CFM_KIND_MHSA: P.mhsa(attn, "", d, H); break;
    case CFM_KIND_REL_ATTENTION: P.rel_attn(attn, "", d, H); break;
    case CFM_KIND_CONV_MODULE: P.conv(conv, "", d, k); break;
    case CFM_KIND_BLOCK: blocks.resize(1); P.block(blocks[0], "", d, H, k); break;
    case CFM_KIND_SUBSAMPLING: P.front(front, "", d); break;
    case CFM_KIND_DECODER:
      CFM_REQUIRE(cfg.vocab_size > 0 && cfg.lstm_hidden_dim > 0 && cfg.n_lstm_layers > 0, "decoder needs vocab_size, lstm_hidden_dim, n_lstm_layers");
      CFM_REQUIRE(!bf16_mode || cfg.lstm_hidden_dim % 8 == 0, "bf16 mode needs lstm_hidden_dim to be a multiple of 8");
      P.decoder(dec, d, cfg.vocab_size, cfg.lstm_hidden_dim, cfg.n_lstm_layers);
      break;
    case CFM_KIND_ENCODER: {
      P.front(front, "downsampling_conv.", d);
      const int F2 = freq_bins();
      const float* lw = P.param("linear.weight", static_cast<int64_t>(d) * d * F2);
      float* lwp = P.alloc_f32(static_cast<int64_t>(d) * d * F2);
      float* lb = const_cast<float*>(P.vec("linear.bias", d));
      if (!P.status && pack_linear_weight(lw, lwp, d, d, F2, stream)) P.status = 2;
      P.finish_linear(lin, lwp, lb, d, d * F2);
      div_term = P.vec("rel_pe.div_term", d / 2);
      blocks.resize(cfg.n_blocks);
      for (int i = 0; i < cfg.n_blocks && !P.status; ++i)
        P.block(blocks[i], "layers." + std::to_string(i) + ".", d, H, k);
      // all layers' pos_proj fused into one [L*d, d] projection of the shared sinusoid table
      if (!P.status) {
        const int L = cfg.n_blocks;
        float* w = P.alloc_f32(static_cast<int64_t>(L) * d * d);
        float* b = P.alloc_f32(static_cast<int64_t>(L) * d);
        for (int i = 0; i < L && !P.status; ++i) {
          cudaMemcpyAsync(w + static_cast<int64_t>(i) * d * d, blocks[i].attn.pos.w32, static_cast<size_t>(d) * d * 4,
                          cudaMemcpyDeviceToDevice, stream);
          cudaMemcpyAsync(b + static_cast<int64_t>(i) * d, blocks[i].attn.pos.b, static_cast<size_t>(d) * 4,
                          cudaMemcpyDeviceToDevice, stream);
        }
        P.finish_linear(pos_all, w, b, L * d, d);
      }
      break;
    }
    default: set_error("unknown module kind"); return 1;
  }
  if (P.status) return P.status;
  if (kind != CFM_KIND_ENCODER && kind != CFM_KIND_FFN && kind != CFM_KIND_CONV_MODULE && kind != CFM_KIND_SUBSAMPLING &&
      kind != CFM_KIND_DECODER) {
    // default sinusoid frequencies for a stand-alone attention/block handle called without `pos`
    std::vector<float> dt(d / 2);
    const float c = static_cast<float>(-(std::log(10000.0) / d));
    for (int i = 0; i < d / 2; ++i) dt[i] = std::exp(static_cast<float>(2 * i) * c);
    float* p = P.alloc_f32(d / 2);
    if (!p) return 4;
    CFM_CUDA_OK(cudaMemcpyAsync(p, dt.data(), dt.size() * 4, cudaMemcpyHostToDevice, stream));
    CFM_CUDA_OK(cudaStreamSynchronize(stream));       // dt is a host temporary
    div_term = p;
  }
  for (const auto& kv : params)
    if (!kv.second.used) {
      set_error("unexpected parameter '" + kv.first + "' for this module kind");
      return 1;
    }
  CFM_CUDA_OK(cudaGetLastError());
  CFM_CUDA_OK(cudaStreamSynchronize(stream));         // the caller may now free/overwrite its tensors
  finalized = true;
  return 0;
}

// =================================================================================== forward
namespace {

template <typename TA> struct Dispatch;
template <> struct Dispatch<float> {
  static int gemm(Ctx& c, const float* A, int lda, const Linear& L, void* C, int ldc, int M, Epilogue e,
                  const float* res, float scale, const ConvGeom* g) {
    ++c.launches;
    Timed t(c, KC_GEMM_F32, 2.0 * M * L.N * L.K, 4.0 * (static_cast<double>(M) * (L.K + L.N) + static_cast<double>(L.N) * L.K));
    return gemm_f32(A, lda, L.w32, L.b, static_cast<float*>(C), ldc, M, L.N, L.K, e, res, scale, g, c.stream);
  }
};
template <> struct Dispatch<bf16> {
  static int gemm(Ctx& c, const bf16* A, int lda, const Linear& L, void* C, int ldc, int M, Epilogue e,
                  const float* res, float scale, const ConvGeom* g) {
    ++c.launches;
    if (e == EPI_BIAS_RES && res != static_cast<const float*>(C)) {
      // the tensor-core kernel updates C in place; seed it with the residual when the caller wants a copy
      CFM_CUDA_OK(cudaMemcpy2DAsync(C, static_cast<size_t>(ldc) * 4, res, static_cast<size_t>(ldc) * 4,
                                    static_cast<size_t>(L.N) * 4, M, cudaMemcpyDeviceToDevice, c.stream));
      res = static_cast<const float*>(C);
    }
    Timed t(c, KC_GEMM_TC, 2.0 * M * L.N * L.K, 2.0 * (static_cast<double>(M) * (L.K + L.N) + static_cast<double>(L.N) * L.K));
    return gemm_bf16_tc(A, lda, L.w16, L.b, C, ldc, M, L.N, L.K, e, res, scale, g, c.num_sms, c.stream);
  }
};

inline int attention_any(const float* qkv, const float* pos, int ldp, const float* u, const float* v, const uint8_t* kp,
                         float* ctx, int B, int Tn, int H, int dh, cudaStream_t st) {
  return rel_attention<float>(qkv, pos, ldp, u, v, kp, ctx, B, Tn, H, dh, st);
}
inline int attention_any(const bf16* qkv, const bf16* pos, int ldp, const float* u, const float* v, const uint8_t* kp,
                         bf16* ctx, int B, int Tn, int H, int dh, cudaStream_t st) {
  if (dh == 64) return rel_attention_bf16_tc(qkv, pos, ldp, u, v, kp, ctx, B, Tn, H, dh, st);
  return rel_attention<bf16>(qkv, pos, ldp, u, v, kp, ctx, B, Tn, H, dh, st);
}

inline int dwconv_fast(const float*, const float*, const float*, float*, int, int, int, int, cudaStream_t) {
  set_error("fused-GLU depthwise path is bf16 only");
  return 1;
}
inline int dwconv_fast(const bf16* g, const float* w, const float* b, bf16* y, int B, int Tn, int d, int k, cudaStream_t st) {
  return dwconv_bn_swish_bf16(g, w, b, y, B, Tn, d, k, st);
}

template <typename TA> TA* ws_take(Ctx& c, int64_t n) { return static_cast<TA*>(c.ws.take(static_cast<size_t>(n) * sizeof(TA))); }

#define KRUN(cls, flops, bytes, call)      \
  do {                                     \
    ++c.launches;                          \
    Timed _t(c, cls, flops, bytes);        \
    CFM_TRY(call);                         \
  } while (0)
#define LN_BYTES(M, d, so) (static_cast<double>(M) * (d) * (4 + (so)))

#define WS_CHECK(ptr) CFM_REQUIRE((ptr) != nullptr || c.ws.measuring(), "workspace too small; query cfm_workspace_bytes")

// out = res + scale * FFN(LN(x))      (ffn.py:15-23, block.py:19,25)
inline bool ffn_fast_ok(const float*, int) { return false; }
inline bool ffn_fast_ok(const bf16*, int d) { return ffn_fused_supported(d); }

// a_pre (bf16 mode, nullable): LN(x) already computed by the kernel that produced x.
template <typename TA>
int ffn_fwd(Ctx& c, const FfnW& w, const float* x, const float* res, float scale, float* out, int64_t M, int d,
            const bf16* a_pre = nullptr) {
  if (ffn_fast_ok(static_cast<const TA*>(nullptr), d)) {
    // fused GEMM -> Swish -> GEMM -> residual kernel on the normalised input.  The kernel only ADDS into `out` (which
    // must hold the residual before it runs) and never reads x, so out may alias x or res.
    const size_t mark = c.ws.mark();
    if (a_pre == nullptr) {
      // no producer normalised x for us (block 0's ffn_1, the stand-alone module, an unfused conv module in front)
      bf16* a = ws_take<bf16>(c, M * d);
      WS_CHECK(a);
      if (!c.ws.measuring())
        KRUN(KC_LAYERNORM, 0, LN_BYTES(M, d, 2), layer_norm<bf16>(x, w.ln_g, w.ln_b, a, M, d, c.stream));
      a_pre = a;
    }
    if (!c.ws.measuring()) {
      if (res == nullptr) CFM_CUDA_OK(cudaMemsetAsync(out, 0, static_cast<size_t>(M) * d * 4, c.stream));
      else if (res != out) CFM_CUDA_OK(cudaMemcpyAsync(out, res, static_cast<size_t>(M) * d * 4, cudaMemcpyDeviceToDevice, c.stream));
      KRUN(KC_FFN_FUSED, 16.0 * M * d * d, 6.0 * M * d + 16.0 * d * d,
           ffn_fused_bf16(a_pre, out, w.hidden.w16, w.hidden.b, w.out.w16, w.out.b, scale, static_cast<int>(M), d, c.num_sms,
                          c.stream));
    }
    c.ws.reset(mark);
    return 0;
  }
  const size_t mark = c.ws.mark();
  TA* a = ws_take<TA>(c, M * d);
  TA* h = ws_take<TA>(c, M * 4 * d);
  WS_CHECK(a); WS_CHECK(h);
  if (!c.ws.measuring()) {
    KRUN(KC_LAYERNORM, 0, LN_BYTES(M, d, sizeof(TA)), layer_norm<TA>(x, w.ln_g, w.ln_b, a, M, d, c.stream));
    CFM_TRY(Dispatch<TA>::gemm(c, a, d, w.hidden, h, 4 * d, static_cast<int>(M), EPI_BIAS_SWISH, nullptr, 0.f, nullptr));
    CFM_TRY(Dispatch<TA>::gemm(c, h, 4 * d, w.out, out, d, static_cast<int>(M), res ? EPI_BIAS_RES : EPI_BIAS_F32, res, scale, nullptr));
  }
  c.ws.reset(mark);
  return 0;
}

// projected positional table for one or all layers: P = PE * Wp^T + bp   (attention.py:81)
template <typename TA>
int pos_project(Ctx& c, const Linear& wp, const float* div_term, const float* pos_f32, int Tn, int d, TA** out) {
  const int R = 2 * Tn - 1;
  TA* pe = ws_take<TA>(c, static_cast<int64_t>(R) * d);
  TA* pp = ws_take<TA>(c, static_cast<int64_t>(R) * wp.N);
  WS_CHECK(pe); WS_CHECK(pp);
  *out = pp;
  if (c.ws.measuring()) return 0;
  if (pos_f32 != nullptr) {
    if (sizeof(TA) == 4) CFM_CUDA_OK(cudaMemcpyAsync(pe, pos_f32, static_cast<size_t>(R) * d * 4, cudaMemcpyDeviceToDevice, c.stream));
    else KRUN(KC_MISC, 0, 6.0 * R * d, cast_f32_to_bf16(pos_f32, reinterpret_cast<bf16*>(pe), static_cast<int64_t>(R) * d, c.stream));
  } else {
    KRUN(KC_MISC, 0, static_cast<double>(R) * d * sizeof(TA), rel_pos_table<TA>(div_term, pe, Tn, d, c.stream));
  }
  return Dispatch<TA>::gemm(c, pe, d, wp, pp, wp.N, R, EPI_BIAS, nullptr, 0.f, nullptr);
}

// sinusoid tables of the factorised positional term (bf16 path, d_model 256): cos|sin of t*w_k as bf16 and fp16
struct FactPos { const bf16* e16 = nullptr; const void* r16 = nullptr; };
// The positional term of a module called WITHOUT a caller-supplied table (the encoder's own rel_pe sinusoids,
// position.py:11-28) is evaluated by the factorised kernel where it exists (bf16, d_model 256, head size 64); a table the
// caller passes in (any values) goes through pos_proj and the relative-shift kernel.  The choice follows from the inputs.
inline bool fact_ok(const float*, const AttnW&, int, int, int) { return false; }
inline bool fact_ok(const bf16*, const AttnW& w, int d, int H, int) {
  return w.wpT != nullptr && rel_attention_fact_supported(d, d / H);
}
template <typename TA>
int fact_tables(Ctx& c, const float* div_term, int Tn, int d, FactPos* fp) {
  bf16* e16 = ws_take<bf16>(c, static_cast<int64_t>(Tn) * d);
  bf16* r16 = ws_take<bf16>(c, static_cast<int64_t>(Tn) * d);       // fp16 payload, same size
  WS_CHECK(e16); WS_CHECK(r16);
  fp->e16 = e16; fp->r16 = r16;
  if (c.ws.measuring()) return 0;
  KRUN(KC_MISC, 0, 4.0 * Tn * d, pos_fact_tables(div_term, e16, r16, Tn, d, c.stream));
  return 0;
}
inline int attention_fact_any(const float*, const AttnW&, const FactPos&, const uint8_t*, float*, int, int, int, int, cudaStream_t) {
  set_error("factorised attention is bf16 only");
  return 1;
}
inline int attention_fact_any(const bf16* qkv, const AttnW& w, const FactPos& fp, const uint8_t* kp, bf16* ctx, int B, int Tn,
                              int H, int dh, cudaStream_t st) {
  return rel_attention_bf16_fact(qkv, w.wpT, fp.e16, fp.r16, w.pos_const, kp, ctx, B, Tn, H, dh, st);
}

// out = res + MHSA(x)      (attention.py:14-18,74-92; block.py:21)
// Positional scores come either from a projected table (pos_proj, any `pos` the caller supplied) or, when fp is
// given, from the factorised sinusoid form (the encoder's own rel_pe table).
template <typename TA>
int mhsa_fwd(Ctx& c, const AttnW& w, bool pre_ln, const float* x, const TA* pos_proj, int ldp, const FactPos* fp,
             const uint8_t* key_pad, const float* res, float* out, int B, int Tn, int d, int H, bool arbitrary_mask = true) {
  const int64_t M = static_cast<int64_t>(B) * Tn;
  const size_t mark = c.ws.mark();
  TA* a = ws_take<TA>(c, M * d);
  TA* qkv = ws_take<TA>(c, M * 3 * d);
  TA* ctx = ws_take<TA>(c, M * d);
  WS_CHECK(a); WS_CHECK(qkv); WS_CHECK(ctx);
  if (!c.ws.measuring()) {
    const TA* ain = a;
    const Linear& wqkv = fp != nullptr ? w.qkv_fact : w.qkv;      // the factorised kernel wants c (q + u) as its queries
    if (pre_ln) {
      KRUN(KC_LAYERNORM, 0, LN_BYTES(M, d, sizeof(TA)), layer_norm<TA>(x, w.ln_g, w.ln_b, a, M, d, c.stream));
    } else if (sizeof(TA) == 4) {
      ain = reinterpret_cast<const TA*>(x);
    } else {
      KRUN(KC_MISC, 0, 6.0 * M * d, cast_f32_to_bf16(x, reinterpret_cast<bf16*>(a), M * d, c.stream));
    }
    CFM_TRY(Dispatch<TA>::gemm(c, ain, d, wqkv, qkv, 3 * d, static_cast<int>(M), EPI_BIAS, nullptr, 0.f, nullptr));
    // algorithmic work: content + positional + PV = 6*T*d FLOPs per token
    if (fp != nullptr) {
      KRUN(KC_ATTENTION, 6.0 * M * Tn * d, static_cast<double>(M) * 4 * d * sizeof(TA),
           attention_fact_any(qkv, w, *fp, key_pad, ctx, B, Tn, H, d / H, c.stream));
    } else {
      KRUN(KC_ATTENTION, 6.0 * M * Tn * d, static_cast<double>(M) * 4 * d * sizeof(TA),
           attention_any(qkv, pos_proj, ldp, w.u, w.v, key_pad, ctx, B, Tn, H, d / H, c.stream));
    }
    // a caller-supplied mask may hide every key of a sample (the encoder's own length masks cannot: lengths' >= 1)
    if (key_pad != nullptr && arbitrary_mask)
      KRUN(KC_MISC, 0, static_cast<double>(M), all_masked_mean_v<TA>(key_pad, qkv, ctx, B, Tn, d, c.stream));
    CFM_TRY(Dispatch<TA>::gemm(c, ctx, d, w.out, out, d, static_cast<int>(M), res ? EPI_BIAS_RES : EPI_BIAS_F32, res, 1.f, nullptr));
  }
  c.ws.reset(mark);
  return 0;
}

// out = res + ConvolutionModule(x)      (convolution.py:21-32; block.py:23)
inline bool conv_fast_ok(const float*, const ConvW&, int) { return false; }
inline bool conv_fast_ok(const bf16*, const ConvW& w, int d) {
  return w.glu_fused && conv_module_fused_supported(d, w.ksize);
}

// next_ln / a_out (fused kernel only, nullable): also emit LN(out; next_ln) as bf16 for the FFN that follows.
template <typename TA>
int conv_fwd(Ctx& c, const ConvW& w, const float* x, const float* res, float* out, int B, int Tn, int d,
             const FfnW* next_ln = nullptr, bf16* a_out = nullptr) {
  const int64_t M = static_cast<int64_t>(B) * Tn;
  if (conv_fast_ok(static_cast<const TA*>(nullptr), w, d) && out != x) {
    // single fused kernel: LN -> pw1 -> GLU -> depthwise+BN+Swish -> pw2 -> +residual
    if (c.ws.measuring()) return 0;
    ++c.launches;
    Timed t(c, KC_CONV_FUSED, 6.0 * M * d * d + 2.0 * w.ksize * M * d, (res ? 12.0 : 8.0) * M * d + 6.0 * d * d);
    return conv_module_fused_bf16(x, res, out, w.ln_g, w.ln_b, w.pw1.w16, w.pw1.b, w.dw_w, w.dw_b, w.pw2.w16, w.pw2.b,
                                  a_out ? next_ln->ln_g : nullptr, a_out ? next_ln->ln_b : nullptr, a_out, B, Tn, d, w.ksize,
                                  c.stream);
  }
  const size_t mark = c.ws.mark();
  TA* a = ws_take<TA>(c, M * d);
  TA* u = ws_take<TA>(c, M * 2 * d);
  TA* g = ws_take<TA>(c, M * d);
  WS_CHECK(a); WS_CHECK(u); WS_CHECK(g);
  if (!c.ws.measuring()) {
    KRUN(KC_LAYERNORM, 0, LN_BYTES(M, d, sizeof(TA)), layer_norm<TA>(x, w.ln_g, w.ln_b, a, M, d, c.stream));
    // algorithmic bytes of the depthwise stage: input [M,d] + output [M,d] + taps
    const double dw_bytes = (2.0 * M * d + (w.ksize + 1.0) * d) * sizeof(TA);
    if (w.glu_fused) {
      CFM_TRY(Dispatch<TA>::gemm(c, a, d, w.pw1, u, d, static_cast<int>(M), EPI_BIAS_GLU, nullptr, 0.f, nullptr));
      KRUN(KC_DWCONV, 2.0 * w.ksize * M * d, dw_bytes, dwconv_fast(u, w.dw_w, w.dw_b, g, B, Tn, d, w.ksize, c.stream));
    } else {
      CFM_TRY(Dispatch<TA>::gemm(c, a, d, w.pw1, u, 2 * d, static_cast<int>(M), EPI_BIAS, nullptr, 0.f, nullptr));
      // (this variant also reads the GLU gate half, which is not counted)
      KRUN(KC_DWCONV, 2.0 * w.ksize * M * d, dw_bytes,
           dwconv_glu_bn_swish<TA>(u, w.dw_w, w.dw_b, g, B, Tn, d, w.ksize, c.stream));
    }
    CFM_TRY(Dispatch<TA>::gemm(c, g, d, w.pw2, out, d, static_cast<int>(M), res ? EPI_BIAS_RES : EPI_BIAS_F32, res, 1.f, nullptr));
  }
  c.ws.reset(mark);
  return 0;
}

// y = ConformerBlock(x)      (block.py:17-29).  xw is an fp32 [M,d] working buffer; y may alias xw.
// bf16 fast path: the producers of an FFN's input also emit its LayerNorm-ed bf16 copy, so the fused FFN kernel starts
// from one TMA load: a1_pre = LN(x; ffn_1) from the previous block's final LayerNorm (nullable); when next_ffn1 is
// given, this block's final LayerNorm writes LN(y; next_ffn1) into a1_next for the block that follows.
template <typename TA>
int block_fwd(Ctx& c, const BlockW& w, const float* x, float* xw, float* y, const TA* pos_proj, int ldp, const FactPos* fp,
              const uint8_t* key_pad, int B, int Tn, int d, int H, const bf16* a1_pre = nullptr,
              const FfnW* next_ffn1 = nullptr, bf16* a1_next = nullptr, bool arbitrary_mask = true) {
  const int64_t M = static_cast<int64_t>(B) * Tn;
  CFM_TRY(ffn_fwd<TA>(c, w.ffn1, x, x, 0.5f, xw, M, d, a1_pre));
  CFM_TRY((mhsa_fwd<TA>(c, w.attn, true, xw, pos_proj, ldp, fp, key_pad, xw, xw, B, Tn, d, H, arbitrary_mask)));
  // the fused conv-module kernel reads a time halo, so it writes a second residual buffer instead of updating in place
  const size_t mark = c.ws.mark();
  float* xc = xw;
  bf16* a2 = nullptr;                                   // LN(conv output; ffn_2) from the fused conv module's epilogue
  if (conv_fast_ok(static_cast<const TA*>(nullptr), w.conv, d)) {
    xc = ws_take<float>(c, M * d);
    WS_CHECK(xc);
    if (ffn_fast_ok(static_cast<const TA*>(nullptr), d)) {
      a2 = ws_take<bf16>(c, M * d);
      WS_CHECK(a2);
    }
  }
  CFM_TRY(conv_fwd<TA>(c, w.conv, xw, xw, xc, B, Tn, d, &w.ffn2, a2));
  CFM_TRY(ffn_fwd<TA>(c, w.ffn2, xc, xc, 0.5f, xc, M, d, a2));
  if (!c.ws.measuring()) {
    if (a1_next != nullptr && next_ffn1 != nullptr) {
      KRUN(KC_LAYERNORM, 0, LN_BYTES(M, d, 4) + 2.0 * M * d,
           layer_norm_dual(xc, w.ln_g, w.ln_b, y, next_ffn1->ln_g, next_ffn1->ln_b, a1_next, M, d, c.stream));
    } else {
      KRUN(KC_LAYERNORM, 0, LN_BYTES(M, d, 4), layer_norm<float>(xc, w.ln_g, w.ln_b, y, M, d, c.stream));
    }
  }
  c.ws.reset(mark);
  return 0;
}

// front-end up to (and excluding) Encoder.linear: h2 [B*T2*F2, C] rows (b,t2,f2), channels-last
template <typename TA>
int front_fwd(Ctx& c, const Module& m, const float* mel, int B, int Tin, TA** h2_out) {
  const int d = m.cfg.d_model, n_mel = m.cfg.n_mel_channels;
  const int T1 = (Tin - 1) / 2, T2 = (T1 - 1) / 2, F1 = (n_mel - 1) / 2, F2 = (F1 - 1) / 2;
  CFM_REQUIRE(Tin >= 7 && T2 >= 1 && F2 >= 1, "input too short for two stride-2 3x3 convolutions");
  const int T1p = T1 + (T1 & 1);
  TA* out1 = ws_take<TA>(c, static_cast<int64_t>(B) * T1p * F1 * d);
  TA* h2 = ws_take<TA>(c, static_cast<int64_t>(B) * T2 * F2 * d);
  WS_CHECK(out1); WS_CHECK(h2);
  *h2_out = h2;
  if (c.ws.measuring()) return 0;
  if (sizeof(TA) == 2 && m.front.c1_w16 != nullptr) {
    KRUN(KC_CONV1, 18.0 * d * F1 * T1 * B, 4.0 * B * n_mel * Tin + static_cast<double>(B) * T1 * F1 * d * sizeof(TA),
         subsample_conv1_tc(mel, m.front.c1_w16, m.front.c1_b, reinterpret_cast<bf16*>(out1), B, n_mel, Tin, T1, T1p, F1, d,
                            c.num_sms, c.stream));
  } else {
    KRUN(KC_CONV1, 18.0 * d * F1 * T1 * B, 4.0 * B * n_mel * Tin + static_cast<double>(B) * T1 * F1 * d * sizeof(TA),
         subsample_conv1<TA>(mel, m.front.c1_w, m.front.c1_b, out1, B, n_mel, Tin, T1, T1p, F1, d, c.stream));
  }
  ConvGeom g{B, T1p, F1, d, T2, F2};
  return Dispatch<TA>::gemm(c, out1, 0, m.front.conv2, h2, d, B * T2 * F2, EPI_BIAS_RELU, nullptr, 0.f, &g);
}

__global__ void permute_front_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int F2, int C) {
  // in [rows][F2][C] -> out [rows][C*F2] with feature c*F2+f (convolution.py:51-52)
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * F2 * C) return;
  const int64_t r = i / (static_cast<int64_t>(F2) * C);
  const int rem = static_cast<int>(i - r * F2 * C);
  const int c = rem / F2, f = rem - c * F2;
  out[i] = in[(r * F2 + f) * C + c];
}
__global__ void permute_front_bf16_kernel(const bf16* __restrict__ in, float* __restrict__ out, int64_t rows, int F2, int C) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * F2 * C) return;
  const int64_t r = i / (static_cast<int64_t>(F2) * C);
  const int rem = static_cast<int>(i - r * F2 * C);
  const int c = rem / F2, f = rem - c * F2;
  out[i] = __bfloat162float(in[(r * F2 + f) * C + c]);
}

}  // namespace

template <typename TA>
int Module::forward_impl(Ctx& c, int op, const float* x, const float* pos, const uint8_t* key_pad,
                         const int64_t* len_in, float* y, int64_t* len_out, int B, int frames) {
  const int d = cfg.d_model, H = cfg.n_heads;
  switch (op) {
    case CFM_KIND_FFN:
      return ffn_fwd<TA>(c, ffn, x, nullptr, 1.f, y, static_cast<int64_t>(B) * frames, d);
    case CFM_KIND_CONV_MODULE:
      return conv_fwd<TA>(c, conv, x, nullptr, y, B, frames, d);
    case CFM_KIND_MHSA:
    case CFM_KIND_REL_ATTENTION: {
      // no table supplied: the module's own sinusoids, factorised (the workspace query covers both routes)
      const bool can_fact = fact_ok(static_cast<const TA*>(nullptr), attn, d, H, frames);
      const bool fact = pos == nullptr && can_fact;
      FactPos fp;
      TA* pp = nullptr;
      if (can_fact && (fact || c.ws.measuring())) CFM_TRY(fact_tables<TA>(c, div_term, frames, d, &fp));
      if (!fact || c.ws.measuring()) CFM_TRY(pos_project<TA>(c, attn.pos, div_term, pos, frames, d, &pp));
      return mhsa_fwd<TA>(c, attn, kind == CFM_KIND_MHSA, x, fact ? static_cast<const TA*>(nullptr) : pp, d, fact ? &fp : nullptr,
                          key_pad, nullptr, y, B, frames, d, H);
    }
    case CFM_KIND_BLOCK: {
      float* xw = ws_take<float>(c, static_cast<int64_t>(B) * frames * d);
      WS_CHECK(xw);
      const bool can_fact = fact_ok(static_cast<const TA*>(nullptr), blocks[0].attn, d, H, frames);
      const bool fact = pos == nullptr && can_fact;
      FactPos fp;
      TA* pp = nullptr;
      if (can_fact && (fact || c.ws.measuring())) CFM_TRY(fact_tables<TA>(c, div_term, frames, d, &fp));
      if (!fact || c.ws.measuring()) CFM_TRY(pos_project<TA>(c, blocks[0].attn.pos, div_term, pos, frames, d, &pp));
      return block_fwd<TA>(c, blocks[0], x, xw, y, fact ? static_cast<const TA*>(nullptr) : pp, d, fact ? &fp : nullptr, key_pad, B,
                           frames, d, H);
    }
    case CFM_KIND_SUBSAMPLING: {
      TA* h2 = nullptr;
      CFM_TRY(front_fwd<TA>(c, *this, x, B, frames, &h2));
      if (c.ws.measuring()) return 0;
      const int T2 = ((frames - 1) / 2 - 1) / 2, F2 = freq_bins();
      const int64_t n = static_cast<int64_t>(B) * T2 * F2 * d;
      ++c.launches;
      if (sizeof(TA) == 4)
        permute_front_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(reinterpret_cast<const float*>(h2), y, static_cast<int64_t>(B) * T2, F2, d);
      else
        permute_front_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c.stream>>>(reinterpret_cast<const bf16*>(h2), y, static_cast<int64_t>(B) * T2, F2, d);
      CFM_CUDA_OK(cudaGetLastError());
      if (len_in != nullptr && len_out != nullptr) {
        ++c.launches;
        CFM_TRY(lengths_and_mask(len_in, len_out, nullptr, B, T2, true, false, c.stream));
      }
      return 0;
    }
    case CFM_KIND_DECODER: {
      // Decoder.forward (decoder.py:15-27): per layer the input projection of all frames as one GEMM, then the
      // persistent recurrence; Swish fused into the last layer's output, BatchNorm folded into the vocabulary GEMM
      const int Hd = dec.H, T = frames;
      const int64_t M = static_cast<int64_t>(B) * T;
      float* xin = ws_take<float>(c, M * 4 * Hd);
      TA* seq_a = ws_take<TA>(c, M * (Hd > d ? Hd : d));
      TA* seq_b = ws_take<TA>(c, M * Hd);
      float* hst = ws_take<float>(c, 2LL * B * Hd);
      float* cst = ws_take<float>(c, static_cast<int64_t>(B) * Hd);
      uint8_t* tcs = ws_take<uint8_t>(c, static_cast<int64_t>(lstm_tc_scratch_bytes(Hd)));
      WS_CHECK(tcs);
      float* lg = ws_take<float>(c, M * dec.Vp);
      WS_CHECK(xin); WS_CHECK(seq_a); WS_CHECK(seq_b); WS_CHECK(hst); WS_CHECK(cst); WS_CHECK(lg);
      if (c.ws.measuring()) return 0;
      const TA* cur = reinterpret_cast<const TA*>(x);
      const int L = static_cast<int>(dec.layers.size());
      // tensor-core recurrence (bf16 mode, hidden size a multiple of 64): the whole head runs TIME-MAJOR ([T, B, .]), see lstm_tc.cu
      const bool tm = sizeof(TA) == 2 && L > 0 && dec.layers[0].w_hh16p != nullptr;
      if (sizeof(TA) == 2) {
        if (tm) KRUN(KC_MISC, 0, 6.0 * M * d, lstm_to_time_major(x, reinterpret_cast<bf16*>(seq_a), B, T, d, c.stream));
        else KRUN(KC_MISC, 0, 6.0 * M * d, cast_f32_to_bf16(x, reinterpret_cast<bf16*>(seq_a), M * d, c.stream));
        cur = seq_a;
      }
      for (int l = 0; l < L; ++l) {
        const int in = l == 0 ? d : Hd;
        CFM_TRY(Dispatch<TA>::gemm(c, cur, in, dec.layers[l].ih, xin, 4 * Hd, static_cast<int>(M), EPI_BIAS_F32, nullptr, 0.f, nullptr));
        TA* dst = (cur == seq_b) ? seq_a : seq_b;
        if (tm) {
          KRUN(KC_MISC, 8.0 * M * Hd * Hd, 4.0 * M * 5 * Hd,
               lstm_layer_forward_tc<TA>(xin, dec.layers[l].w_hh16p, len_in, tcs, dst, B, T, Hd, l == L - 1, c.num_sms, c.stream));
        } else {
          KRUN(KC_MISC, 8.0 * M * Hd * Hd, 4.0 * M * 5 * Hd,
               lstm_layer_forward<TA>(xin, dec.layers[l].w_hh, len_in, hst, cst, dst, B, T, Hd, l == L - 1, c.num_sms, c.stream));
        }
        cur = dst;
      }
      CFM_TRY(Dispatch<TA>::gemm(c, cur, Hd, dec.out, lg, dec.Vp, static_cast<int>(M), EPI_BIAS_F32, nullptr, 0.f, nullptr));
      if (tm) {
        KRUN(KC_MISC, 0, 8.0 * M * dec.V, lstm_logits_from_time_major(lg, y, B, T, dec.V, dec.Vp, c.stream));
      } else {
        CFM_CUDA_OK(cudaMemcpy2DAsync(y, static_cast<size_t>(dec.V) * 4, lg, static_cast<size_t>(dec.Vp) * 4,
                                      static_cast<size_t>(dec.V) * 4, static_cast<size_t>(M), cudaMemcpyDeviceToDevice, c.stream));
      }
      return 0;
    }
    case CFM_KIND_ENCODER: {
      const int T2 = ((frames - 1) / 2 - 1) / 2, F2 = freq_bins();
      const int64_t M = static_cast<int64_t>(B) * T2;
      TA* h2 = nullptr;
      CFM_TRY(front_fwd<TA>(c, *this, x, B, frames, &h2));
      float* xw = ws_take<float>(c, M * d);
      uint8_t* kp = nullptr;
      int64_t* lens = nullptr;
      if (len_in != nullptr) {
        kp = ws_take<uint8_t>(c, M);
        lens = ws_take<int64_t>(c, B);
        WS_CHECK(kp); WS_CHECK(lens);
      }
      WS_CHECK(xw);
      TA* pp = nullptr;
      FactPos fp;
      const bool fact = cfg.n_blocks > 0 && fact_ok(static_cast<const TA*>(nullptr), blocks[0].attn, d, H, T2);
      if (fact) CFM_TRY(fact_tables<TA>(c, div_term, T2, d, &fp));
      else CFM_TRY(pos_project<TA>(c, pos_all, div_term, nullptr, T2, d, &pp));
      if (!c.ws.measuring()) {
        // Encoder.linear on the channels-last flatten (weight columns permuted at pack time), encoder.py:23
        CFM_TRY(Dispatch<TA>::gemm(c, h2, F2 * d, lin, xw, d, static_cast<int>(M), EPI_BIAS_F32, nullptr, 0.f, nullptr));
        if (len_in != nullptr) {
          ++c.launches;
          CFM_TRY(lengths_and_mask(len_in, len_out ? len_out : lens, kp, B, T2, true, false, c.stream));
        }
      }
      const int L = cfg.n_blocks;
      // LN(block output; next block's ffn_1) travels between blocks as bf16 (fused-FFN fast path only)
      bf16* a1 = nullptr;
      if (L > 1 && ffn_fast_ok(static_cast<const TA*>(nullptr), d)) {
        a1 = ws_take<bf16>(c, M * d);
        WS_CHECK(a1);
      }
      for (int i = 0; i < L; ++i) {
        float* dst = (i == L - 1) ? y : xw;
        const bool to_next = a1 != nullptr && i + 1 < L;
        CFM_TRY(block_fwd<TA>(c, blocks[i], xw, xw, dst, fact ? nullptr : pp + static_cast<int64_t>(i) * d, L * d, fact ? &fp : nullptr,
                              kp, B, T2, d, H, (a1 != nullptr && i > 0) ? a1 : nullptr, to_next ? &blocks[i + 1].ffn1 : nullptr,
                              to_next ? a1 : nullptr, /*arbitrary_mask=*/false));
      }
      if (L == 0 && !c.ws.measuring())
        CFM_CUDA_OK(cudaMemcpyAsync(y, xw, static_cast<size_t>(M) * d * 4, cudaMemcpyDeviceToDevice, c.stream));
      return 0;
    }
  }
  set_error("unknown forward op");
  return 1;
}

int Module::forward(int op, const float* x, const float* pos, const uint8_t* key_pad, const int64_t* len_in,
                    float* y, int64_t* len_out, int B, int frames, void* workspace, int64_t ws_bytes,
                    cudaStream_t stream, int64_t* needed) {
  CFM_REQUIRE(finalized, "module not finalized (call cfm_module_finalize after registering parameters)");
  CFM_REQUIRE(op == kind || (op == CFM_KIND_MHSA && kind == CFM_KIND_REL_ATTENTION), "forward does not match the module kind");
  CFM_REQUIRE(B > 0 && frames > 0, "empty batch");
  Ctx c;
  c.stream = stream;
  c.num_sms = num_sms;
  c.prof = &prof;
  c.ws.init(needed ? nullptr : workspace, needed ? 0 : static_cast<size_t>(ws_bytes), needed != nullptr);
  int s = bf16_mode ? forward_impl<bf16>(c, op, x, pos, key_pad, len_in, y, len_out, B, frames)
                    : forward_impl<float>(c, op, x, pos, key_pad, len_in, y, len_out, B, frames);
  if (needed) *needed = static_cast<int64_t>(c.ws.peak());
  last_launches = c.launches;
  return s;
}

}  // namespace cfm
