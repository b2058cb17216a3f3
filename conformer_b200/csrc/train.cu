// Training step of the encoder (BASELINE config 4): forward in training mode with everything the backward needs kept on
// a caller-supplied "tape", and the backward pass producing fp32 gradients for every parameter.
//
// What the reference does (train.py:232-243): model(x, lengths) under autocast in TRAINING mode - dropout active
// (ffn.py:19,21, attention.py:12,67, convolution.py:19, encoder.py:25), BatchNorm1d with the batch statistics of every
// frame in the batch, padded frames included, and a running-statistics update (convolution.py:16,27) - then CTC loss,
// backward through autograd, optimizer step; DistributedDataParallel averages the gradients over the ranks
// (train.py:185-186).  Here the encoder's forward and backward are explicit: dense work on the GEMM kernels (tcgen05 in
// CFM_COMPUTE_BF16, fp32 CUDA cores in CFM_COMPUTE_F32 - the gradient-parity mode), with the weight-gradient GEMMs
// contracting over the rows (operands transposed by transpose_pad, split-K over the SMs, accumulated by reduction
// stores straight into the fp32 gradient buffer), attention by attention_train.cu, the rest by train_kernels.cu.
// The handle keeps POINTERS to the live fp32 master parameters and to their gradient buffers (cfm_module_set_param /
// cfm_module_set_grad); cfm_train_refresh re-derives the bf16 / transposed / packed copies after every optimizer step.
// The backward runs block by block (cfm_encoder_train_backward takes a block range), so the caller can start the
// gradient all-reduce of a finished block while the next one is still computing.
#include <cstring>

#include "module.h"
#include "train.h"

namespace cfm {

namespace {

// dropout sites
enum { SITE_INPUT = 0, SITE_FFN1_H = 1, SITE_FFN1_O = 2, SITE_ATTN_P = 3, SITE_ATTN_O = 4, SITE_CONV_O = 5, SITE_FFN2_H = 6, SITE_FFN2_O = 7, SITES_PER_BLOCK = 8 };

template <typename TA> struct WSel;
template <> struct WSel<float> {
  static const float* w(const TrainLinear& l) { return l.wg; }
  static const float* wT(const TrainLinear& l) { return static_cast<const float*>(l.wT); }
};
template <> struct WSel<bf16> {
  static const bf16* w(const TrainLinear& l) { return l.w16; }
  static const bf16* wT(const TrainLinear& l) { return static_cast<const bf16*>(l.wT); }
};

struct TC {                               // per-call context
  cudaStream_t st;
  int num_sms;
  Workspace* ws;
  float p;
  const uint32_t* seed;                   // device
  Drop drop(int block, int site) const { return make_drop(p, seed, static_cast<uint32_t>(1 + block * SITES_PER_BLOCK + site)); }
};

inline int gemm_any(const TC& c, const float* A, int lda, const float* W, const float* bias, void* C, int ldc, int M, int N, int K,
                    Epilogue e, const float* res, float scale, const ConvGeom* g = nullptr) {
  return gemm_f32(A, lda, W, bias, static_cast<float*>(C), ldc, M, N, K, e, res, scale, g, c.st);
}
inline int gemm_any(const TC& c, const bf16* A, int lda, const bf16* W, const float* bias, void* C, int ldc, int M, int N, int K,
                    Epilogue e, const float* res, float scale, const ConvGeom* g = nullptr) {
  return gemm_bf16_tc(A, lda, W, bias, C, ldc, M, N, K, e, res, scale, g, c.num_sms, c.st, /*split_k=*/true);
}

template <typename T> T* take(const TC& c, int64_t n) { return static_cast<T*>(c.ws->take(static_cast<size_t>(n) * sizeof(T))); }
#define TAKE(var, T, n)                                                                                          \
  T* var = take<T>(c, (n));                                                                                      \
  CFM_REQUIRE(var != nullptr || c.ws->measuring(), "training workspace too small; query cfm_encoder_train_workspace_bytes")

// training forward of the attention core: the tcgen05 kernel where it exists (bf16, head size 64), else the CUDA-core one
inline int attn_fwd_any(const float* qkv, const float* pp, const float* u, const float* v, const uint8_t* kp, float* ctx, float* lse, int B,
                        int Tn, int H, int dh, Drop dr, cudaStream_t st) {
  return attn_train_fwd<float>(qkv, pp, u, v, kp, ctx, lse, B, Tn, H, dh, dr, st);
}
inline int attn_fwd_any(const bf16* qkv, const bf16* pp, const float* u, const float* v, const uint8_t* kp, bf16* ctx, float* lse, int B,
                        int Tn, int H, int dh, Drop dr, cudaStream_t st) {
  if (dh == 64) return rel_attention_bf16_tc_train(qkv, pp, H * dh, u, v, kp, ctx, lse, dr, B, Tn, H, dh, st);
  return attn_train_fwd<bf16>(qkv, pp, u, v, kp, ctx, lse, B, Tn, H, dh, dr, st);
}

inline int round8(int64_t v) { return static_cast<int>((v + 7) & ~static_cast<int64_t>(7)); }

// dW[N_w, K_w] += dY[M, N_w]^T X[M, K_w]  and  db[N_w] += colsum(dY):  the weight-gradient GEMM contracts over the M rows.
// bf16: the tcgen05 GEMM reads both operands in place as MN-major tiles (gemm_bf16_tn); fp32: transposed copies + the
// CUDA-core GEMM.
inline int wgrad_gemm(const TC& c, const bf16* dY, int ld_dy, const bf16* X, int ld_x, float* dW, int M, int Nw, int Kw, bf16*, bf16*) {
  return gemm_bf16_tn(dY, ld_dy, X, ld_x, dW, Kw, Nw, Kw, M, c.num_sms, c.st);
}
inline int wgrad_gemm(const TC& c, const float* dY, int ld_dy, const float* X, int ld_x, float* dW, int M, int Nw, int Kw, float* tA,
                      float* tB) {
  const int Mp = round8(M);
  CFM_TRY(transpose_pad<float>(dY, ld_dy, tA, M, Nw, Mp, c.st));
  CFM_TRY(transpose_pad<float>(X, ld_x, tB, M, Kw, Mp, c.st));
  return gemm_any(c, static_cast<const float*>(tA), Mp, static_cast<const float*>(tB), nullptr, dW, Kw, Nw, Kw, Mp, EPI_BIAS_RES, dW, 1.f);
}
template <typename TA>
int wgrad(const TC& c, const TA* dY, int ld_dy, const TA* X, int ld_x, float* dW, float* db, int M, int Nw, int Kw, TA* tA, TA* tB) {
  CFM_TRY(wgrad_gemm(c, dY, ld_dy, X, ld_x, dW, M, Nw, Kw, tA, tB));
  if (db != nullptr) CFM_TRY(colsum_add<TA>(dY, ld_dy, db, M, Nw, c.st));
  return 0;
}

// conv_2 weight gradient in the packed layout: gpack[C, 9C] += dy2[M2, C]^T im2col(out1)[M2, 9C]
inline int conv2_wgrad(const TC& c, const bf16* dy2, const bf16* out1, float* gpack, const ConvGeom& g, int64_t M2, int64_t, bf16* big1, bf16*) {
  CFM_TRY(im2col_rows(out1, big1, g, M2, c.st));                  // row-major im2col, read in place by the TN GEMM
  return gemm_bf16_tn(dy2, g.C, big1, 9 * g.C, gpack, 9 * g.C, g.C, 9 * g.C, static_cast<int>(M2), c.num_sms, c.st);
}
inline int conv2_wgrad(const TC& c, const float* dy2, const float* out1, float* gpack, const ConvGeom& g, int64_t M2, int64_t M2p, float* big1,
                       float* big2) {
  const int d = g.C, M2i = static_cast<int>(M2), M2pi = static_cast<int>(M2p);
  CFM_TRY(transpose_pad<float>(dy2, d, big2, M2i, d, M2pi, c.st));                              // dy2^T [C, M2p]
  CFM_TRY(im2col_t<float>(out1, big1, g, M2, M2p, c.st));                                       // im2col^T [9C, M2p]
  return gemm_any(c, static_cast<const float*>(big2), M2pi, static_cast<const float*>(big1), nullptr, gpack, 9 * d, d, 9 * d, M2pi,
                  EPI_BIAS_RES, gpack, 1.f);
}

__global__ void unpack_conv2_grad_kernel(const float* __restrict__ gp, float* __restrict__ g, int C) {
  pdl_wait();
  // gp[n][(kw*3+kh)*C + c] -> g[n][c][kh][kw] (+=)
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(C) * C * 9) return;
  const int kw = static_cast<int>(i % 3), kh = static_cast<int>((i / 3) % 3);
  const int cc = static_cast<int>((i / 9) % C), n = static_cast<int>(i / (9 * static_cast<int64_t>(C)));
  g[i] += gp[static_cast<int64_t>(n) * 9 * C + static_cast<int64_t>(kw * 3 + kh) * C + cc];
}
__global__ void unpack_linear_grad_kernel(const float* __restrict__ gp, float* __restrict__ g, int d, int C, int F2) {
  pdl_wait();
  // gp[n][f*C + c] -> g[n][c*F2 + f] (+=)
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t K = static_cast<int64_t>(C) * F2;
  if (i >= K * d) return;
  const int64_t n = i / K, r = i - n * K;
  const int cc = static_cast<int>(r / F2), f = static_cast<int>(r - static_cast<int64_t>(cc) * F2);
  g[i] += gp[n * K + static_cast<int64_t>(f) * C + cc];
}

}  // namespace

// =================================================================================== parameter binding
int Module::train_bind(int compute_mode) {
  CFM_REQUIRE(kind == CFM_KIND_ENCODER, "training is implemented for the Encoder handle");
  CFM_REQUIRE(compute_mode == 0 || compute_mode == 1, "compute must be CFM_COMPUTE_F32 or CFM_COMPUTE_BF16");
  const int d = cfg.d_model, H = cfg.n_heads, k = cfg.kernel_size, L = cfg.n_blocks, F2 = freq_bins();
  CFM_REQUIRE(d % 8 == 0 && H > 0 && d % H == 0 && (d / H) % 4 == 0 && d / H <= 64, "training needs d_model % 8 == 0 and head size % 4 == 0, <= 64");
  CFM_REQUIRE(k % 2 == 1 && k <= 63, "training needs an odd depthwise kernel size up to 63 (the reference uses 31)");
  int dev = 0;
  CFM_CUDA_OK(cudaGetDevice(&dev));
  CFM_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  bf16_mode = compute_mode == 1;
  if (bf16_mode) CFM_TRY(tc_init());
  for (auto& kv : params) kv.second.used = false;
  // arena: bf16 copy + transposed copy of every matrix, packed conv_2 / linear copies
  size_t total = 0;
  for (const auto& kv : params) total += static_cast<size_t>(kv.second.numel);
  CFM_TRY(arena.reserve(total * 16 + (params.size() * 4 + 64) * 256 + (1 << 20)));
  int status = 0;
  auto P = [&](const std::string& name, int64_t numel) -> const float* {
    auto it = params.find(name);
    if (it == params.end() || it->second.numel != numel) {
      if (!status) set_error("missing or mis-sized parameter '" + name + "'");
      status = 1;
      return nullptr;
    }
    it->second.used = true;
    return it->second.data;
  };
  auto Gp = [&](const std::string& name, int64_t numel) -> float* {
    auto it = grads.find(name);
    if (it == grads.end() || it->second.numel != numel) {
      if (!status) set_error("missing or mis-sized gradient buffer for '" + name + "' (cfm_module_set_grad)");
      status = 1;
      return nullptr;
    }
    return const_cast<float*>(it->second.data);
  };
  const size_t esz = bf16_mode ? 2 : 4;
  auto lin = [&](TrainLinear& l, const std::string& pre, int N, int K) {
    l.N = N; l.K = K;
    l.w = P(pre + ".weight", static_cast<int64_t>(N) * K);
    l.wg = l.w;
    l.b = P(pre + ".bias", N);
    l.dw = Gp(pre + ".weight", static_cast<int64_t>(N) * K);
    l.db = Gp(pre + ".bias", N);
    l.w16 = bf16_mode ? static_cast<bf16*>(arena.take(static_cast<size_t>(N) * K * 2)) : nullptr;
    l.wT = arena.take(static_cast<size_t>(N) * K * esz);
    if (l.wT == nullptr && !status) { set_error("weight arena exhausted"); status = 4; }
  };
  auto ln = [&](TrainNorm& n, const std::string& pre) {
    n.g = P(pre + ".weight", d); n.b = P(pre + ".bias", d);
    n.dg = Gp(pre + ".weight", d); n.db = Gp(pre + ".bias", d);
  };
  TrainState& t = train;
  train_drop_graphs();                                       // every pointer below is baked into the captured launches
  t.seed_dev = static_cast<uint32_t*>(arena.take(256));
  if (t.seed_dev == nullptr) { set_error("weight arena exhausted"); return 4; }
  t.blocks.assign(L, TrainBlock{});
  lin(t.conv2, "downsampling_conv.conv_2", d, 9 * d);
  t.conv2_packed = static_cast<float*>(arena.take(static_cast<size_t>(d) * 9 * d * 4));
  t.conv2.wg = t.conv2_packed;
  t.c1_w = P("downsampling_conv.conv_1.weight", static_cast<int64_t>(d) * 9);
  t.c1_b = P("downsampling_conv.conv_1.bias", d);
  t.dc1_w = Gp("downsampling_conv.conv_1.weight", static_cast<int64_t>(d) * 9);
  t.dc1_b = Gp("downsampling_conv.conv_1.bias", d);
  lin(t.lin, "linear", d, d * F2);
  t.lin_packed = static_cast<float*>(arena.take(static_cast<size_t>(d) * d * F2 * 4));
  t.lin.wg = t.lin_packed;
  t.div_term = P("rel_pe.div_term", d / 2);
  for (int i = 0; i < L && !status; ++i) {
    TrainBlock& b = t.blocks[i];
    const std::string pre = "layers." + std::to_string(i) + ".";
    ln(b.ffn1_ln, pre + "ffn_1.layer_norm");
    lin(b.ffn1_h, pre + "ffn_1.hidden_linear", 4 * d, d);
    lin(b.ffn1_o, pre + "ffn_1.out_linear", d, 4 * d);
    ln(b.att_ln, pre + "attention.layer_norm");
    lin(b.q, pre + "attention.attention.query_proj", d, d);
    lin(b.k, pre + "attention.attention.key_proj", d, d);
    lin(b.v, pre + "attention.attention.value_proj", d, d);
    lin(b.pos, pre + "attention.attention.pos_proj", d, d);
    lin(b.o, pre + "attention.attention.out_proj", d, d);
    b.u = P(pre + "attention.attention.content_bias", d); b.du = Gp(pre + "attention.attention.content_bias", d);
    b.vb = P(pre + "attention.attention.position_bias", d); b.dvb = Gp(pre + "attention.attention.position_bias", d);
    ln(b.conv_ln, pre + "conv.layer_norm");
    lin(b.pw1, pre + "conv.pointwise_conv_1", 2 * d, d);
    lin(b.pw2, pre + "conv.pointwise_conv_2", d, d);
    b.dw_w = P(pre + "conv.deepwise_conv.weight", static_cast<int64_t>(d) * k); b.ddw_w = Gp(pre + "conv.deepwise_conv.weight", static_cast<int64_t>(d) * k);
    b.dw_b = P(pre + "conv.deepwise_conv.bias", d); b.ddw_b = Gp(pre + "conv.deepwise_conv.bias", d);
    b.bn_g = P(pre + "conv.batch_norm.weight", d); b.dbn_g = Gp(pre + "conv.batch_norm.weight", d);
    b.bn_b = P(pre + "conv.batch_norm.bias", d); b.dbn_b = Gp(pre + "conv.batch_norm.bias", d);
    b.bn_rm = const_cast<float*>(P(pre + "conv.batch_norm.running_mean", d));
    b.bn_rv = const_cast<float*>(P(pre + "conv.batch_norm.running_var", d));
    ln(b.ffn2_ln, pre + "ffn_2.layer_norm");
    lin(b.ffn2_h, pre + "ffn_2.hidden_linear", 4 * d, d);
    lin(b.ffn2_o, pre + "ffn_2.out_linear", d, 4 * d);
    ln(b.out_ln, pre + "layer_norm");
  }
  if (status) return status;
  for (const auto& kv : params)
    if (!kv.second.used) { set_error("unexpected parameter '" + kv.first + "'"); return 1; }
  train.bound = true;
  return 0;
}

// bf16 / transposed / packed copies of the (just updated) master weights
template <typename TA>
static int refresh_linear(const TrainLinear& l, const float* src, bool bf16_mode, cudaStream_t st) {
  // src: the [N,K] fp32 matrix the GEMMs see (the parameter itself, or its packed form)
  if (bf16_mode) {
    CFM_TRY(cast_f32_to_bf16(src, l.w16, static_cast<int64_t>(l.N) * l.K, st));
    CFM_TRY(transpose_pad<bf16>(l.w16, l.K, static_cast<bf16*>(l.wT), l.N, l.K, l.N, st));
  } else {
    CFM_TRY(transpose_pad<float>(src, l.K, static_cast<float*>(l.wT), l.N, l.K, l.N, st));
  }
  return 0;
}
static int train_graphed_refresh(Module& m, cudaStream_t st);

int Module::train_refresh(cudaStream_t st) {
  CFM_REQUIRE(train.bound, "call cfm_train_bind first");
  return train_graphed_refresh(*this, st);
}

static int train_refresh_body(Module& m, cudaStream_t st) {
  const bool bf16_mode = m.bf16_mode;
  const int d = m.cfg.d_model, F2 = m.freq_bins();
  TrainState& t = m.train;
  CFM_TRY(pack_conv2_weight(t.conv2.w, t.conv2_packed, d, st));
  CFM_TRY(pack_linear_weight(t.lin.w, t.lin_packed, d, d, F2, st));
  CFM_TRY(refresh_linear<float>(t.conv2, t.conv2_packed, bf16_mode, st));
  CFM_TRY(refresh_linear<float>(t.lin, t.lin_packed, bf16_mode, st));
  for (TrainBlock& b : t.blocks) {
    TrainLinear* ls[] = {&b.ffn1_h, &b.ffn1_o, &b.q, &b.k, &b.v, &b.pos, &b.o, &b.pw1, &b.pw2, &b.ffn2_h, &b.ffn2_o};
    for (TrainLinear* l : ls) CFM_TRY(refresh_linear<float>(*l, l->wg, bf16_mode, st));
  }
  return 0;
}

// =================================================================================== forward
template <typename TA>
int Module::train_forward_impl(const float* mel, const int64_t* len_in, float* y, int64_t* len_out, int B, int Tin, Workspace& ws,
                               float p_drop, cudaStream_t st) {
  const int d = cfg.d_model, H = cfg.n_heads, dh = d / H, ks = cfg.kernel_size, n_mel = cfg.n_mel_channels, L = cfg.n_blocks;
  const int T1 = (Tin - 1) / 2, T2 = (T1 - 1) / 2, F1 = (n_mel - 1) / 2, F2 = (F1 - 1) / 2, T1p = T1 + (T1 & 1);
  CFM_REQUIRE(Tin >= 7 && T2 >= 1 && F2 >= 1, "input too short for two stride-2 3x3 convolutions");
  const int64_t M = static_cast<int64_t>(B) * T2, M2 = M * F2;
  const int R = 2 * T2 - 1;
  TC c{st, num_sms, &ws, p_drop, train.seed_dev};
  TrainState& t = train;
  const bool run = !ws.measuring();
  t.tape = TrainTape{};
  TrainTape& tp = t.tape;
  tp.B = B; tp.Tin = Tin; tp.T2 = T2; tp.p = p_drop; tp.mel = mel;
  // ---- tape (lives until the backward) ----
  TAKE(out1, TA, static_cast<int64_t>(B) * T1p * F1 * d);
  TAKE(h2, TA, M2 * d);
  TAKE(x0, float, M * d);
  TAKE(pe, TA, static_cast<int64_t>(R) * d);
  uint8_t* kp = nullptr;
  int64_t* lens = nullptr;
  if (len_in != nullptr) {
    kp = take<uint8_t>(c, M);
    lens = take<int64_t>(c, B);
  }
  tp.out1 = out1; tp.h2 = h2; tp.x0 = x0; tp.pe = pe; tp.kp = kp;
  tp.blocks.assign(L, BlockTape{});
  for (int i = 0; i < L; ++i) {
    BlockTape& bt = tp.blocks[i];
    bt.a1 = take<TA>(c, M * d); bt.hp1 = take<TA>(c, M * 4 * d); bt.x1 = take<float>(c, M * d);
    bt.a2 = take<TA>(c, M * d); bt.qkv = take<TA>(c, M * 3 * d); bt.pp = take<TA>(c, static_cast<int64_t>(R) * d);
    bt.ctx = take<TA>(c, M * d); bt.lse = take<float>(c, static_cast<int64_t>(B) * H * T2); bt.x2 = take<float>(c, M * d);
    bt.a3 = take<TA>(c, M * d); bt.u = take<TA>(c, M * 2 * d); bt.g = take<TA>(c, M * d); bt.z = take<float>(c, M * d);
    bt.bn = take<float>(c, 2 * d); bt.x3 = take<float>(c, M * d);
    bt.a4 = take<TA>(c, M * d); bt.hp4 = take<TA>(c, M * 4 * d); bt.x4 = take<float>(c, M * d);
    bt.y = take<float>(c, M * d);
    CFM_REQUIRE(bt.y != nullptr || !run, "training workspace too small; query cfm_encoder_train_workspace_bytes");
  }
  tp.end_of_tape = ws.mark();
  // ---- scratch (dead after the forward) ----
  TAKE(hd, TA, M * 4 * d);            // dropout(swish(hidden)) / swish(bn(z))
  TAKE(tmp, float, M * d);            // branch output before the residual add
  TAKE(acc, float, 2 * d);
  if (!run) return 0;

  // front-end (convolution.py:42-57, encoder.py:23-25)
  CFM_TRY(subsample_conv1<TA>(mel, t.c1_w, t.c1_b, static_cast<TA*>(out1), B, n_mel, Tin, T1, T1p, F1, d, st));
  ConvGeom g{B, T1p, F1, d, T2, F2};
  CFM_TRY(gemm_any(c, static_cast<const TA*>(out1), 0, WSel<TA>::w(t.conv2), t.conv2.b, h2, d, static_cast<int>(M2), d, 9 * d,
                   EPI_BIAS_RELU, nullptr, 0.f, &g));
  CFM_TRY(gemm_any(c, static_cast<const TA*>(h2), F2 * d, WSel<TA>::w(t.lin), t.lin.b, x0, d, static_cast<int>(M), d, F2 * d, EPI_BIAS_F32,
                   nullptr, 0.f));
  CFM_TRY(drop_inplace(x0, M * d, make_drop(p_drop, t.seed_dev, SITE_INPUT), st));
  if (len_in != nullptr) CFM_TRY(lengths_and_mask(len_in, len_out ? len_out : lens, kp, B, T2, true, false, st));
  CFM_TRY(rel_pos_table<TA>(t.div_term, static_cast<TA*>(pe), T2, d, st));

  const float* x = x0;
  for (int i = 0; i < L; ++i) {
    const TrainBlock& w = t.blocks[i];
    BlockTape& bt = tp.blocks[i];
    TA *a1 = static_cast<TA*>(bt.a1), *hp1 = static_cast<TA*>(bt.hp1), *a2 = static_cast<TA*>(bt.a2), *qkv = static_cast<TA*>(bt.qkv);
    TA *pp = static_cast<TA*>(bt.pp), *ctx = static_cast<TA*>(bt.ctx), *a3 = static_cast<TA*>(bt.a3), *u = static_cast<TA*>(bt.u);
    TA *gg = static_cast<TA*>(bt.g), *a4 = static_cast<TA*>(bt.a4), *hp4 = static_cast<TA*>(bt.hp4);
    bt.x = x;
    const int Mi = static_cast<int>(M);
    // x1 = x + 0.5 * FFN1(x)                                                                    (block.py:19, ffn.py:15-23)
    CFM_TRY(layer_norm<TA>(x, w.ffn1_ln.g, w.ffn1_ln.b, a1, M, d, st));
    CFM_TRY(gemm_any(c, a1, d, WSel<TA>::w(w.ffn1_h), w.ffn1_h.b, hp1, 4 * d, Mi, 4 * d, d, EPI_BIAS, nullptr, 0.f));
    CFM_TRY(swish_drop_fwd<TA>(hp1, hd, M * 4 * d, c.drop(i, SITE_FFN1_H), st));
    CFM_TRY(gemm_any(c, static_cast<const TA*>(hd), 4 * d, WSel<TA>::w(w.ffn1_o), w.ffn1_o.b, tmp, d, Mi, d, 4 * d, EPI_BIAS_F32, nullptr, 0.f));
    CFM_TRY(residual_drop_add(x, tmp, bt.x1, M * d, 0.5f, c.drop(i, SITE_FFN1_O), st));
    // x2 = x1 + MHSA(x1)                                                                        (block.py:21, attention.py:14-18,74-92)
    CFM_TRY(layer_norm<TA>(bt.x1, w.att_ln.g, w.att_ln.b, a2, M, d, st));
    CFM_TRY(gemm_any(c, a2, d, WSel<TA>::w(w.q), w.q.b, qkv, 3 * d, Mi, d, d, EPI_BIAS, nullptr, 0.f));
    CFM_TRY(gemm_any(c, a2, d, WSel<TA>::w(w.k), w.k.b, qkv + d, 3 * d, Mi, d, d, EPI_BIAS, nullptr, 0.f));
    CFM_TRY(gemm_any(c, a2, d, WSel<TA>::w(w.v), w.v.b, qkv + 2 * d, 3 * d, Mi, d, d, EPI_BIAS, nullptr, 0.f));
    CFM_TRY(gemm_any(c, static_cast<const TA*>(pe), d, WSel<TA>::w(w.pos), w.pos.b, pp, d, R, d, d, EPI_BIAS, nullptr, 0.f));
    CFM_TRY(attn_fwd_any(qkv, pp, w.u, w.vb, kp, ctx, bt.lse, B, T2, H, dh, c.drop(i, SITE_ATTN_P), st));
    CFM_TRY(gemm_any(c, static_cast<const TA*>(ctx), d, WSel<TA>::w(w.o), w.o.b, tmp, d, Mi, d, d, EPI_BIAS_F32, nullptr, 0.f));
    CFM_TRY(residual_drop_add(bt.x1, tmp, bt.x2, M * d, 1.f, c.drop(i, SITE_ATTN_O), st));
    // x3 = x2 + Conv(x2)                                                                        (block.py:23, convolution.py:21-32)
    CFM_TRY(layer_norm<TA>(bt.x2, w.conv_ln.g, w.conv_ln.b, a3, M, d, st));
    CFM_TRY(gemm_any(c, a3, d, WSel<TA>::w(w.pw1), w.pw1.b, u, 2 * d, Mi, 2 * d, d, EPI_BIAS, nullptr, 0.f));
    CFM_TRY(glu_fwd<TA>(u, gg, M, d, st));
    CFM_TRY(dwconv_fwd<TA>(gg, w.dw_w, w.dw_b, bt.z, B, T2, d, ks, st));
    CFM_TRY(bn_train_stats(bt.z, acc, bt.bn, w.bn_rm, w.bn_rv, Mi, d, 0.1f, st));
    CFM_TRY(bn_swish_fwd<TA>(bt.z, bt.bn, w.bn_g, w.bn_b, hd, M, d, st));
    CFM_TRY(gemm_any(c, static_cast<const TA*>(hd), d, WSel<TA>::w(w.pw2), w.pw2.b, tmp, d, Mi, d, d, EPI_BIAS_F32, nullptr, 0.f));
    CFM_TRY(residual_drop_add(bt.x2, tmp, bt.x3, M * d, 1.f, c.drop(i, SITE_CONV_O), st));
    // x4 = x3 + 0.5 * FFN2(x3);  y = LayerNorm(x4)                                              (block.py:25-27)
    CFM_TRY(layer_norm<TA>(bt.x3, w.ffn2_ln.g, w.ffn2_ln.b, a4, M, d, st));
    CFM_TRY(gemm_any(c, a4, d, WSel<TA>::w(w.ffn2_h), w.ffn2_h.b, hp4, 4 * d, Mi, 4 * d, d, EPI_BIAS, nullptr, 0.f));
    CFM_TRY(swish_drop_fwd<TA>(hp4, hd, M * 4 * d, c.drop(i, SITE_FFN2_H), st));
    CFM_TRY(gemm_any(c, static_cast<const TA*>(hd), 4 * d, WSel<TA>::w(w.ffn2_o), w.ffn2_o.b, tmp, d, Mi, d, 4 * d, EPI_BIAS_F32, nullptr, 0.f));
    CFM_TRY(residual_drop_add(bt.x3, tmp, bt.x4, M * d, 0.5f, c.drop(i, SITE_FFN2_O), st));
    CFM_TRY(layer_norm<float>(bt.x4, w.out_ln.g, w.out_ln.b, bt.y, M, d, st));
    x = bt.y;
  }
  CFM_CUDA_OK(cudaMemcpyAsync(y, x, static_cast<size_t>(M) * d * 4, cudaMemcpyDeviceToDevice, st));
  tp.valid = true;
  return 0;
}

// =================================================================================== backward
// one FeedForwardModule half-step: given dx_out (fp32, in `dx`, updated in place to dx_in)
template <typename TA>
static int ffn_bwd(const TC& c, const TrainNorm& ln, const TrainLinear& h, const TrainLinear& o, const float* x_in, const TA* a, const TA* hp,
                   float* dx, int64_t M, int d, Drop drop_h, Drop drop_o, TA* dt, TA* hd, float* f32buf, TA* tA, TA* tB) {
  const int Mi = static_cast<int>(M);
  CFM_TRY(cast_scale_drop<TA>(dx, dt, M * d, 0.5f, drop_o, c.st));                               // d(out_linear output)
  CFM_TRY(swish_drop_fwd<TA>(hp, hd, M * 4 * d, drop_h, c.st));                                  // recompute the GEMM's input
  CFM_TRY(wgrad<TA>(c, dt, d, hd, 4 * d, o.dw, o.db, Mi, d, 4 * d, tA, tB));
  CFM_TRY(gemm_any(c, static_cast<const TA*>(dt), d, WSel<TA>::wT(o), nullptr, f32buf, 4 * d, Mi, 4 * d, d, EPI_BIAS_F32, nullptr, 0.f));
  CFM_TRY(swish_drop_bwd<TA>(f32buf, hp, hd, M * 4 * d, drop_h, c.st));                          // d(hidden pre-activation)
  CFM_TRY(wgrad<TA>(c, hd, 4 * d, a, d, h.dw, h.db, Mi, 4 * d, d, tA, tB));
  CFM_TRY(gemm_any(c, static_cast<const TA*>(hd), 4 * d, WSel<TA>::wT(h), nullptr, f32buf, d, Mi, d, 4 * d, EPI_BIAS_F32, nullptr, 0.f));
  return ln_bwd(f32buf, x_in, ln.g, dx, true, ln.dg, ln.db, M, d, c.st);
}

template <typename TA>
int Module::train_backward_impl(const float* dy, int block_hi, int block_lo, bool front, void* ws_base, int64_t ws_bytes, cudaStream_t st) {
  TrainState& t = train;
  TrainTape& tp = t.tape;
  CFM_REQUIRE(tp.valid, "no forward on the tape (cfm_encoder_train_forward must precede the backward)");
  const int d = cfg.d_model, H = cfg.n_heads, dh = d / H, ks = cfg.kernel_size, n_mel = cfg.n_mel_channels, L = cfg.n_blocks;
  const int B = tp.B, Tin = tp.Tin, T2 = tp.T2;
  const int T1 = (Tin - 1) / 2, F1 = (n_mel - 1) / 2, F2 = (F1 - 1) / 2, T1p = T1 + (T1 & 1);
  const int64_t M = static_cast<int64_t>(B) * T2, M2 = M * F2;
  const int Mi = static_cast<int>(M), R = 2 * T2 - 1, Mp = round8(M);
  CFM_REQUIRE(block_hi <= L && block_lo >= 0 && block_lo <= block_hi, "bad block range");
  // scratch lives behind the tape in the same workspace
  Workspace ws;
  ws.init(ws_base, static_cast<size_t>(ws_bytes), false);
  ws.take(tp.end_of_tape);
  TC c{st, num_sms, &ws, tp.p, train.seed_dev};
  TAKE(dx, float, M * d);                 // gradient of the residual stream (persists across calls: first allocation)
  TAKE(f32a, float, M * 4 * d);
  TAKE(f32b, float, M * d);
  TAKE(ta, TA, M * 4 * d);
  TAKE(tb, TA, M * 4 * d);
  TAKE(tA, TA, static_cast<int64_t>(4) * d * Mp);
  TAKE(tB, TA, static_cast<int64_t>(4) * d * Mp);
  TAKE(dpp, float, static_cast<int64_t>(R) * d);
  TAKE(dsc, float, static_cast<int64_t>(B) * H * T2);
  TAKE(acc, float, 2 * d);
  if (dy != nullptr) CFM_CUDA_OK(cudaMemcpyAsync(dx, dy, static_cast<size_t>(M) * d * 4, cudaMemcpyDeviceToDevice, st));

  for (int i = block_hi - 1; i >= block_lo; --i) {
    const TrainBlock& w = t.blocks[i];
    const BlockTape& bt = tp.blocks[i];
    const TA *a1 = static_cast<const TA*>(bt.a1), *hp1 = static_cast<const TA*>(bt.hp1), *a2 = static_cast<const TA*>(bt.a2);
    const TA *qkv = static_cast<const TA*>(bt.qkv), *pp = static_cast<const TA*>(bt.pp), *ctx = static_cast<const TA*>(bt.ctx);
    const TA *a3 = static_cast<const TA*>(bt.a3), *u = static_cast<const TA*>(bt.u), *gg = static_cast<const TA*>(bt.g);
    const TA *a4 = static_cast<const TA*>(bt.a4), *hp4 = static_cast<const TA*>(bt.hp4);
    // y = LN(x4)
    CFM_TRY(ln_bwd(dx, bt.x4, w.out_ln.g, f32b, false, w.out_ln.dg, w.out_ln.db, M, d, st));
    CFM_CUDA_OK(cudaMemcpyAsync(dx, f32b, static_cast<size_t>(M) * d * 4, cudaMemcpyDeviceToDevice, st));
    // x4 = x3 + 0.5 FFN2(x3)
    CFM_TRY(ffn_bwd<TA>(c, w.ffn2_ln, w.ffn2_h, w.ffn2_o, bt.x3, a4, hp4, dx, M, d, c.drop(i, SITE_FFN2_H), c.drop(i, SITE_FFN2_O), ta, tb,
                        f32a, tA, tB));
    // x3 = x2 + Conv(x2)
    {
      CFM_TRY(cast_scale_drop<TA>(dx, ta, M * d, 1.f, c.drop(i, SITE_CONV_O), st));                 // d(pointwise_conv_2 output)
      CFM_TRY(bn_swish_fwd<TA>(bt.z, bt.bn, w.bn_g, w.bn_b, tb, M, d, st));                         // recompute its input
      CFM_TRY(wgrad<TA>(c, ta, d, tb, d, w.pw2.dw, w.pw2.db, Mi, d, d, tA, tB));
      CFM_TRY(gemm_any(c, static_cast<const TA*>(ta), d, WSel<TA>::wT(w.pw2), nullptr, f32a, d, Mi, d, d, EPI_BIAS_F32, nullptr, 0.f));
      CFM_TRY(bn_swish_bwd(f32a, bt.z, bt.bn, w.bn_g, w.bn_b, acc, w.dbn_g, w.dbn_b, Mi, d, st));   // f32a: dz
      CFM_TRY(dwconv_bwd_weight<TA>(f32a, gg, w.ddw_w, w.ddw_b, B, T2, d, ks, st));
      CFM_TRY(dwconv_bwd_input(f32a, w.dw_w, f32b, B, T2, d, ks, st));                              // f32b: dg
      CFM_TRY(glu_bwd<TA>(f32b, u, ta, M, d, st));                                                  // ta: du [M, 2d]
      CFM_TRY(wgrad<TA>(c, ta, 2 * d, a3, d, w.pw1.dw, w.pw1.db, Mi, 2 * d, d, tA, tB));
      CFM_TRY(gemm_any(c, static_cast<const TA*>(ta), 2 * d, WSel<TA>::wT(w.pw1), nullptr, f32b, d, Mi, d, 2 * d, EPI_BIAS_F32, nullptr, 0.f));
      CFM_TRY(ln_bwd(f32b, bt.x2, w.conv_ln.g, dx, true, w.conv_ln.dg, w.conv_ln.db, M, d, st));
    }
    // x2 = x1 + MHSA(x1)
    {
      CFM_TRY(cast_scale_drop<TA>(dx, ta, M * d, 1.f, c.drop(i, SITE_ATTN_O), st));                 // d(out_proj output)
      CFM_TRY(wgrad<TA>(c, ta, d, ctx, d, w.o.dw, w.o.db, Mi, d, d, tA, tB));
      CFM_TRY(gemm_any(c, static_cast<const TA*>(ta), d, WSel<TA>::wT(w.o), nullptr, f32b, d, Mi, d, d, EPI_BIAS_F32, nullptr, 0.f));   // dctx
      CFM_CUDA_OK(cudaMemsetAsync(dpp, 0, static_cast<size_t>(R) * d * 4, st));
      CFM_TRY(attn_train_bwd<TA>(qkv, pp, w.u, w.vb, static_cast<const uint8_t*>(tp.kp), ctx, bt.lse, f32b, tb, dpp, w.du, w.dvb, dsc, B, T2,
                                 H, dh, c.drop(i, SITE_ATTN_P), st));                                // tb: dqkv [M, 3d]
      CFM_TRY(wgrad<TA>(c, tb, 3 * d, a2, d, w.q.dw, w.q.db, Mi, d, d, tA, tB));
      CFM_TRY(wgrad<TA>(c, tb + d, 3 * d, a2, d, w.k.dw, w.k.db, Mi, d, d, tA, tB));
      CFM_TRY(wgrad<TA>(c, tb + 2 * d, 3 * d, a2, d, w.v.dw, w.v.db, Mi, d, d, tA, tB));
      CFM_TRY(gemm_any(c, static_cast<const TA*>(tb), 3 * d, WSel<TA>::wT(w.q), nullptr, f32a, d, Mi, d, d, EPI_BIAS_F32, nullptr, 0.f));
      CFM_TRY(gemm_any(c, static_cast<const TA*>(tb) + d, 3 * d, WSel<TA>::wT(w.k), nullptr, f32a, d, Mi, d, d, EPI_BIAS_RES, f32a, 1.f));
      CFM_TRY(gemm_any(c, static_cast<const TA*>(tb) + 2 * d, 3 * d, WSel<TA>::wT(w.v), nullptr, f32a, d, Mi, d, d, EPI_BIAS_RES, f32a, 1.f));
      CFM_TRY(ln_bwd(f32a, bt.x1, w.att_ln.g, dx, true, w.att_ln.dg, w.att_ln.db, M, d, st));
      // pos_proj: pp = pe Wp^T + bp over the shared sinusoid table (its gradient summed over the batch by the attention kernel)
      CFM_TRY(cast_scale_drop<TA>(dpp, ta, static_cast<int64_t>(R) * d, 1.f, Drop{}, st));
      CFM_TRY(wgrad<TA>(c, ta, d, static_cast<const TA*>(tp.pe), d, w.pos.dw, w.pos.db, R, d, d, tA, tB));
    }
    // x1 = x + 0.5 FFN1(x)
    CFM_TRY(ffn_bwd<TA>(c, w.ffn1_ln, w.ffn1_h, w.ffn1_o, bt.x, a1, hp1, dx, M, d, c.drop(i, SITE_FFN1_H), c.drop(i, SITE_FFN1_O), ta, tb,
                        f32a, tA, tB));
  }
  if (!front) return 0;
  CFM_REQUIRE(block_lo == 0, "the front-end backward follows block 0");
  // ---- front-end: x0 = dropout(linear(flatten(relu(conv_2(relu(conv_1(mel)))))))
  ws.reset(tp.end_of_tape);
  ws.take(static_cast<size_t>(M) * d * 4);                       // dx stays where it is (first allocation above)
  const int64_t M2p = (M2 + 7) & ~static_cast<int64_t>(7);
  TAKE(dxl, TA, M * d);
  TAKE(dh2f, float, M2 * d);
  TAKE(dy2, TA, M2 * d);
  TAKE(big1, TA, static_cast<int64_t>(9) * d * M2p);            // colT, then dcol
  TAKE(big2, TA, static_cast<int64_t>(d) * (M2p > static_cast<int64_t>(F2) * d ? M2p : static_cast<int64_t>(F2) * d) + static_cast<int64_t>(F2) * d * Mp);
  TAKE(gpack, float, static_cast<int64_t>(d) * 9 * d > static_cast<int64_t>(d) * d * F2 ? static_cast<int64_t>(d) * 9 * d : static_cast<int64_t>(d) * d * F2);
  TAKE(dout1, float, static_cast<int64_t>(B) * T1p * F1 * d);
  const TA* h2 = static_cast<const TA*>(tp.h2);
  const TA* out1 = static_cast<const TA*>(tp.out1);
  CFM_TRY(cast_scale_drop<TA>(dx, dxl, M * d, 1.f, make_drop(tp.p, t.seed_dev, SITE_INPUT), st));
  // Encoder.linear on the channels-last flatten: gradient in the packed column order, then scattered to the parameter's
  {
    const int K = F2 * d;
    CFM_CUDA_OK(cudaMemsetAsync(gpack, 0, static_cast<size_t>(d) * K * 4, st));
    CFM_TRY(wgrad_gemm(c, static_cast<const TA*>(dxl), d, h2, K, gpack, Mi, d, K, big2, big2 + static_cast<int64_t>(d) * Mp));
    const int64_t n = static_cast<int64_t>(d) * K;
    CFM_CUDA_OK(launch_pdl(unpack_linear_grad_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, 1, gpack, t.lin.dw, d, d, F2));
    CFM_CUDA_OK(cudaGetLastError());
    CFM_TRY(colsum_add<TA>(dxl, d, t.lin.db, Mi, d, st));
    CFM_TRY(gemm_any(c, static_cast<const TA*>(dxl), d, WSel<TA>::wT(t.lin), nullptr, dh2f, K, Mi, K, d, EPI_BIAS_F32, nullptr, 0.f));
  }
  CFM_TRY(relu_bwd<TA>(dh2f, h2, dy2, M2 * d, st));
  // conv_2 as the implicit GEMM [M2, 9C] x [C, 9C]^T
  {
    ConvGeom g{B, T1p, F1, d, T2, F2};
    const int M2i = static_cast<int>(M2), M2pi = static_cast<int>(M2p);
    CFM_CUDA_OK(cudaMemsetAsync(gpack, 0, static_cast<size_t>(d) * 9 * d * 4, st));
    CFM_TRY(conv2_wgrad(c, dy2, out1, gpack, g, M2, M2p, big1, big2));
    const int64_t n = static_cast<int64_t>(d) * d * 9;
    CFM_CUDA_OK(launch_pdl(unpack_conv2_grad_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, st, 1, gpack, t.conv2.dw, d));
    CFM_CUDA_OK(cudaGetLastError());
    CFM_TRY(colsum_add<TA>(dy2, d, t.conv2.db, M2i, d, st));
    // dcol = dy2 W2p  ([M2, C] x [9C, C]^T), gathered back onto conv_1's output grid through its ReLU
    CFM_TRY(gemm_any(c, static_cast<const TA*>(dy2), d, WSel<TA>::wT(t.conv2), nullptr, big1, 9 * d, M2i, 9 * d, d, EPI_BIAS, nullptr, 0.f));
    // conv_1's parameter gradients straight from dcol: its output gradient (relu' * col2im(dcol), [B,T1,F1,C] fp32 - 0.96 GB
    // at C4) is gathered on the fly and never written
    if (d % 8 == 0 && d / 8 <= 256) {
      CFM_TRY(conv1_wgrad_fused<TA>(big1, out1, tp.mel, t.dc1_w, t.dc1_b, g, n_mel, Tin, T1, c.num_sms, st));
    } else {
      CFM_TRY(col2im_relu<TA>(big1, out1, dout1, g, st));
      CFM_TRY(conv1_wgrad(dout1, tp.mel, t.dc1_w, t.dc1_b, B, n_mel, Tin, T1, T1p, F1, d, st));
    }
  }
  return 0;
}

int64_t Module::train_workspace_bytes(int B, int Tin) {
  if (!train.bound) { set_error("call cfm_train_bind first"); return -1; }
  Workspace ws;
  ws.init(nullptr, 0, true);
  int s = bf16_mode ? train_forward_impl<bf16>(nullptr, reinterpret_cast<const int64_t*>(8), nullptr, nullptr, B, Tin, ws, 0.f, nullptr)
                    : train_forward_impl<float>(nullptr, reinterpret_cast<const int64_t*>(8), nullptr, nullptr, B, Tin, ws, 0.f, nullptr);
  if (s != 0) return -1;
  const size_t tape = train.tape.end_of_tape, fwd_peak = ws.peak();
  // backward scratch (the larger of the block phase and the front-end phase), sized generously from the shapes
  const int d = cfg.d_model, H = cfg.n_heads, n_mel = cfg.n_mel_channels;
  const int T1 = (Tin - 1) / 2, T2 = (T1 - 1) / 2, F1 = (n_mel - 1) / 2, F2 = (F1 - 1) / 2, T1p = T1 + (T1 & 1);
  const size_t M = static_cast<size_t>(B) * T2, M2 = M * F2, Mp = (M + 7) & ~size_t(7), M2p = (M2 + 7) & ~size_t(7), es = bf16_mode ? 2 : 4;
  const size_t blocks = M * d * 4 + M * 4 * d * 4 + M * d * 4 + 2 * M * 4 * d * es + 2 * 4 * d * Mp * es + size_t(2 * T2) * d * 4 +
                        size_t(B) * H * T2 * 4 + size_t(d) * (2 * T2 + 8) * es + 64 * 256;
  const size_t front = M * d * 4 + M * d * es + M2 * d * 4 + M2 * d * es + 9 * size_t(d) * M2p * es +
                       (size_t(d) * (M2p > size_t(F2) * d ? M2p : size_t(F2) * d) + size_t(F2) * d * Mp) * es + size_t(d) * d * (F2 > 9 ? F2 : 9) * 4 +
                       size_t(B) * T1p * F1 * d * 4 + 64 * 256;
  const size_t bwd = tape + (blocks > front ? blocks : front);
  train.tape = TrainTape{};
  return static_cast<int64_t>((fwd_peak > bwd ? fwd_peak : bwd) + 4096);
}

// ---- CUDA graphs of the step's launch sequences -------------------------------------------------------------------
// The training step is ~2 500 small launches and the host cannot issue them as fast as the GPU retires them.  Every
// argument of those launches is a function of (pointers, shapes, dropout rate) - the dropout seed lives in device memory -
// so a sequence seen twice with the same key is captured (on a private stream: the caller's may be the legacy default
// stream, which cannot capture) and replayed from then on with one cudaGraphLaunch on the caller's stream.
namespace {
__global__ void set_seed_kernel(uint32_t* dst, uint32_t lo, uint32_t hi) {
  pdl_wait();
  dst[0] = lo;
  dst[1] = hi;
}
uint64_t key_ptr(const void* p) { return static_cast<uint64_t>(reinterpret_cast<uintptr_t>(p)); }
}  // namespace

void Module::train_drop_graphs() {
  for (TrainGraph& g : train.graphs)
    if (g.exec != nullptr) cudaGraphExecDestroy(g.exec);
  train.graphs.clear();
}

template <class Body>
static int train_graphed(Module& m, const uint64_t (&key)[10], bool is_forward, cudaStream_t st, Body&& body) {
  TrainState& t = m.train;
  if (!t.graphs_on) {
    ++t.n_direct;
    return body(st);
  }
  TrainGraph* g = nullptr;
  for (TrainGraph& c : t.graphs)
    if (std::memcmp(c.key, key, sizeof(key)) == 0) {
      g = &c;
      break;
    }
  if (g != nullptr && g->exec != nullptr) {
    if (is_forward) t.tape = g->tape;
    CFM_CUDA_OK(cudaGraphLaunch(g->exec, st));
    ++t.n_replayed;
    return 0;
  }
  if (g == nullptr) {                                            // first sighting: run directly (also sets kernel attributes)
    if (t.graphs.size() < 512) {
      TrainGraph n;
      std::memcpy(n.key, key, sizeof(key));
      n.seen = 1;
      t.graphs.push_back(n);
    }
    ++t.n_direct;
    return body(st);
  }
  if (g->dead) {
    ++t.n_direct;
    return body(st);
  }
  if (t.cap_stream == nullptr) CFM_CUDA_OK(cudaStreamCreateWithFlags(&t.cap_stream, cudaStreamNonBlocking));
  if (cudaStreamBeginCapture(t.cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    g->dead = true;
    ++t.n_failed;
    return body(st);
  }
  const int s = body(t.cap_stream);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(t.cap_stream, &graph);
  if (s != 0 || e != cudaSuccess || graph == nullptr) {
    if (graph != nullptr) cudaGraphDestroy(graph);
    cudaGetLastError();
    g->dead = true;
    ++t.n_failed;
    return s != 0 ? s : body(st);
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess || exec == nullptr) {
    cudaGetLastError();
    g->dead = true;
    ++t.n_failed;
    return body(st);
  }
  g->exec = exec;
  ++t.n_captured;
  if (is_forward) g->tape = t.tape;
  CFM_CUDA_OK(cudaGraphLaunch(g->exec, st));
  return 0;
}

static int train_graphed_refresh(Module& m, cudaStream_t st) {
  const uint64_t key[10] = {3};
  return train_graphed(m, key, false, st, [&](cudaStream_t s) { return train_refresh_body(m, s); });
}

int Module::train_forward(const float* mel, const int64_t* len_in, float* y, int64_t* len_out, int B, int Tin, void* wsp, int64_t ws_bytes,
                          float p_drop, uint64_t seed, cudaStream_t st) {
  CFM_REQUIRE(train.bound, "call cfm_train_bind first");
  CFM_REQUIRE(mel != nullptr && y != nullptr && wsp != nullptr && B > 0, "bad training-forward arguments");
  CFM_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "dropout rate must be in [0, 1)");
  CFM_CUDA_OK(launch_pdl(set_seed_kernel, dim3(1), dim3(1), 0, st, 1, train.seed_dev, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  CFM_CUDA_OK(cudaGetLastError());
  uint32_t pbits;
  std::memcpy(&pbits, &p_drop, 4);
  const uint64_t key[10] = {1, key_ptr(mel), key_ptr(len_in), key_ptr(y), key_ptr(len_out), key_ptr(wsp), static_cast<uint64_t>(ws_bytes),
                            static_cast<uint64_t>(B), static_cast<uint64_t>(Tin), pbits};
  return train_graphed(*this, key, true, st, [&](cudaStream_t s) {
    Workspace ws;
    ws.init(wsp, static_cast<size_t>(ws_bytes), false);
    return bf16_mode ? train_forward_impl<bf16>(mel, len_in, y, len_out, B, Tin, ws, p_drop, s)
                     : train_forward_impl<float>(mel, len_in, y, len_out, B, Tin, ws, p_drop, s);
  });
}

int Module::train_backward(const float* dy, int block_hi, int block_lo, bool front, void* wsp, int64_t ws_bytes, cudaStream_t st) {
  CFM_REQUIRE(train.bound, "call cfm_train_bind first");
  const TrainTape& tp = train.tape;
  CFM_REQUIRE(tp.valid, "cfm_encoder_train_backward needs the tape of a preceding cfm_encoder_train_forward");
  uint32_t pbits;
  std::memcpy(&pbits, &tp.p, 4);
  // the tape's layout is a function of (workspace, B, Tin, lengths given); mel is read again by conv_1's weight gradient
  const uint64_t key[10] = {2, key_ptr(dy), key_ptr(wsp), static_cast<uint64_t>(ws_bytes), key_ptr(tp.mel),
                            (static_cast<uint64_t>(tp.B) << 32) | static_cast<uint64_t>(tp.Tin), tp.kp != nullptr ? 1u : 0u, pbits,
                            (static_cast<uint64_t>(block_hi) << 32) | static_cast<uint64_t>(block_lo), front ? 1u : 0u};
  return train_graphed(*this, key, false, st, [&](cudaStream_t s) {
    return bf16_mode ? train_backward_impl<bf16>(dy, block_hi, block_lo, front, wsp, ws_bytes, s)
                     : train_backward_impl<float>(dy, block_hi, block_lo, front, wsp, ws_bytes, s);
  });
}

}  // namespace cfm
