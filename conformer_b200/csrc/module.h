// Module handle internals (see include/conformer_b200.h for the public ABI).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/conformer_b200.h"
#include "ops.h"

namespace cfm {

class Arena {                        // one cudaMalloc holding every packed weight of a module
 public:
  ~Arena() { release(); }
  int reserve(size_t bytes);
  void release();
  void* take(size_t bytes);          // 256-byte aligned; nullptr when exhausted
 private:
  void* base_ = nullptr;
  size_t cap_ = 0, used_ = 0;
};

class Workspace {                    // bump allocator over the caller's scratch; can also just measure
 public:
  void init(void* base, size_t cap, bool measuring) {
    base_ = static_cast<char*>(base); cap_ = cap; used_ = peak_ = 0; measuring_ = measuring;
  }
  void* take(size_t bytes) {
    const size_t a = (used_ + 255) & ~static_cast<size_t>(255);
    used_ = a + bytes;
    if (used_ > peak_) peak_ = used_;
    if (measuring_ || used_ > cap_) return nullptr;
    return base_ + a;
  }
  size_t mark() const { return used_; }
  void reset(size_t m) { used_ = m; }
  size_t peak() const { return peak_; }
  bool measuring() const { return measuring_; }
 private:
  char* base_ = nullptr;
  size_t cap_ = 0, used_ = 0, peak_ = 0;
  bool measuring_ = false;
};

// kernel classes for the optional per-launch CUDA-event profile (bench.py's roofline numbers)
enum KernelClass { KC_GEMM_TC = 0, KC_GEMM_F32, KC_LAYERNORM, KC_ATTENTION, KC_DWCONV, KC_CONV1, KC_MISC, KC_FFN_FUSED, KC_CONV_FUSED, KC_COUNT };

struct ProfRec { int cls; double flops, bytes; cudaEvent_t e0, e1; };
struct Profiler {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  size_t next = 0;
  cudaEvent_t get() {
    if (next == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[next++];
  }
  ~Profiler() { for (auto e : pool) cudaEventDestroy(e); }
};

struct Ctx {
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  int64_t launches = 0;
  Workspace ws;
  Profiler* prof = nullptr;
};

// brackets one kernel launch with events on the launching stream when profiling is enabled
struct Timed {
  Ctx& c;
  ProfRec rec;
  bool live;
  Timed(Ctx& ctx, int cls, double flops, double bytes) : c(ctx), live(ctx.prof != nullptr && ctx.prof->on) {
    if (!live) return;
    rec = ProfRec{cls, flops, bytes, c.prof->get(), c.prof->get()};
    cudaEventRecord(rec.e0, c.stream);
  }
  ~Timed() {
    if (!live) return;
    cudaEventRecord(rec.e1, c.stream);
    c.prof->recs.push_back(rec);
  }
};

struct Linear {
  const float* w32 = nullptr;        // [N,K]
  bf16* w16 = nullptr;               // [N,K], bf16 mode only
  const float* b = nullptr;          // [N]
  int N = 0, K = 0;
};
struct FfnW { const float *ln_g = nullptr, *ln_b = nullptr; Linear hidden, out; };
struct AttnW {
  const float *ln_g = nullptr, *ln_b = nullptr, *u = nullptr, *v = nullptr;
  Linear qkv, pos, out;
  bf16* wpT = nullptr;               // pos_proj packed per head for the factorised positional term (bf16, d = 256)
  Linear qkv_fact;                   // q|k|v with the query rows as c (Wq x + bq + u) for the factorised kernel
  const float* pos_const = nullptr;  // [H][d]: c (v - u) Wp_h (factorised kernel)
};
struct ConvW {
  const float *ln_g = nullptr, *ln_b = nullptr, *dw_w = nullptr, *dw_b = nullptr;
  int ksize = 0;
  bool glu_fused = false;            // pw1 rows packed for the GEMM's GLU epilogue (bf16 fast path)
  Linear pw1, pw2;
};
struct BlockW { FfnW ffn1, ffn2; AttnW attn; ConvW conv; const float *ln_g = nullptr, *ln_b = nullptr; };
struct FrontW {
  const float *c1_w = nullptr, *c1_b = nullptr;
  bf16* c1_w16 = nullptr;            // conv_1 weight packed [C][64] for the tensor-core kernel (bf16 mode, C = 256)
  Linear conv2;
};

struct LstmLayerW {
  Linear ih;                         // [4H, in] with bias b_ih + b_hh (tensor-core recurrence: rows in its packed order)
  const float* w_hh = nullptr;       // [4H, H] fp32, nn.LSTM row order (CUDA-core recurrence)
  bf16* w_hh16p = nullptr;           // [4H, H] bf16, packed row order (lstm_tc.cu); null: CUDA-core recurrence
};
struct DecW {
  std::vector<LstmLayerW> layers;
  Linear out;                        // vocabulary projection with the eval-mode BatchNorm folded in, rows padded to a multiple of 8
  int V = 0, Vp = 0, H = 0;
};

// ---- training step (train.cu): pointers to the LIVE master parameters / gradient buffers + derived copies
struct TrainLinear {
  const float* w = nullptr;          // the parameter [N,K] (fp32 master)
  const float* wg = nullptr;         // the [N,K] fp32 matrix the GEMMs use (w itself, or its packed form for conv_2 / linear)
  const float* b = nullptr;
  float *dw = nullptr, *db = nullptr;
  bf16* w16 = nullptr;               // bf16 copy of wg (bf16 mode)
  void* wT = nullptr;                // [K,N] transposed copy in the activation type (input-gradient GEMMs)
  int N = 0, K = 0;
};
struct TrainNorm { const float *g = nullptr, *b = nullptr; float *dg = nullptr, *db = nullptr; };
struct TrainBlock {
  TrainNorm ffn1_ln, att_ln, conv_ln, ffn2_ln, out_ln;
  TrainLinear ffn1_h, ffn1_o, q, k, v, pos, o, pw1, pw2, ffn2_h, ffn2_o;
  const float *u = nullptr, *vb = nullptr, *dw_w = nullptr, *dw_b = nullptr, *bn_g = nullptr, *bn_b = nullptr;
  float *du = nullptr, *dvb = nullptr, *ddw_w = nullptr, *ddw_b = nullptr, *dbn_g = nullptr, *dbn_b = nullptr;
  float *bn_rm = nullptr, *bn_rv = nullptr;      // running statistics, updated in place by the training forward
};
struct BlockTape {                   // what one ConformerBlock's backward needs from its forward
  const float* x = nullptr;
  void *a1 = nullptr, *hp1 = nullptr, *a2 = nullptr, *qkv = nullptr, *pp = nullptr, *ctx = nullptr, *a3 = nullptr, *u = nullptr,
       *g = nullptr, *a4 = nullptr, *hp4 = nullptr;
  float *x1 = nullptr, *lse = nullptr, *x2 = nullptr, *z = nullptr, *bn = nullptr, *x3 = nullptr, *x4 = nullptr, *y = nullptr;
};
struct TrainTape {
  bool valid = false;
  int B = 0, Tin = 0, T2 = 0;
  float p = 0.f;
  const float* mel = nullptr;
  void *out1 = nullptr, *h2 = nullptr, *pe = nullptr, *kp = nullptr;
  float* x0 = nullptr;
  size_t end_of_tape = 0;
  std::vector<BlockTape> blocks;
};
// One captured launch sequence of the training step (a whole forward, one block range of the backward, the weight
// refresh): the second call with the same key captures it on a private stream, later calls replay it.
struct TrainGraph {
  uint64_t key[10] = {};
  int seen = 0;
  bool dead = false;                   // capture failed once: this key always launches directly
  cudaGraphExec_t exec = nullptr;
  TrainTape tape;                      // forward graphs: the tape the captured run laid out
};
struct TrainState {
  bool bound = false;
  bool graphs_on = true;
  uint32_t* seed_dev = nullptr;        // {lo, hi} of the step's dropout seed
  cudaStream_t cap_stream = nullptr;
  std::vector<TrainGraph> graphs;
  int64_t n_direct = 0, n_captured = 0, n_replayed = 0, n_failed = 0;   // cfm_train_graph_stats
  TrainLinear conv2, lin;
  float *conv2_packed = nullptr, *lin_packed = nullptr;
  const float *c1_w = nullptr, *c1_b = nullptr, *div_term = nullptr;
  float *dc1_w = nullptr, *dc1_b = nullptr;
  std::vector<TrainBlock> blocks;
  TrainTape tape;
};

struct ParamRef { const float* data; int64_t numel; bool used; };

struct Module {
  int kind = 0;
  cfm_config_t cfg{};
  std::unordered_map<std::string, ParamRef> params;
  std::unordered_map<std::string, ParamRef> grads;       // cfm_module_set_grad: fp32 gradient buffer of each parameter
  TrainState train;
  bool finalized = false, bf16_mode = false;
  int num_sms = 148;
  int64_t last_launches = 0;
  Arena arena;
  Profiler prof;
  // cfm_encoder_forward_host_pipelined: ordering events of each in-flight slot (a slot = its staging workspace)
  struct HostSlot { cudaEvent_t copied_in = nullptr, computed = nullptr, copied_out = nullptr; };
  std::unordered_map<void*, HostSlot> host_slots;
  ~Module() {
    train_drop_graphs();
    if (train.cap_stream != nullptr) cudaStreamDestroy(train.cap_stream);
    for (auto& kv : host_slots) {
      if (kv.second.copied_in) cudaEventDestroy(kv.second.copied_in);
      if (kv.second.computed) cudaEventDestroy(kv.second.computed);
      if (kv.second.copied_out) cudaEventDestroy(kv.second.copied_out);
    }
  }
  // packed weights (which members are populated depends on `kind`)
  FfnW ffn;
  AttnW attn;
  ConvW conv;
  FrontW front;
  Linear lin, pos_all;
  const float* div_term = nullptr;
  std::vector<BlockW> blocks;
  DecW dec;

  int freq_bins() const { return ((cfg.n_mel_channels - 1) / 2 - 1) / 2; }
  int finalize(int compute_mode, cudaStream_t stream);
  // needed != nullptr: only measure the workspace the call would use
  int forward(int op, const float* x, const float* pos, const uint8_t* key_pad, const int64_t* len_in, float* y,
              int64_t* len_out, int B, int frames, void* workspace, int64_t ws_bytes, cudaStream_t stream,
              int64_t* needed);
  // training step (train.cu)
  int train_bind(int compute_mode);
  int train_refresh(cudaStream_t stream);
  int64_t train_workspace_bytes(int B, int Tin);
  int train_forward(const float* mel, const int64_t* len_in, float* y, int64_t* len_out, int B, int Tin, void* ws, int64_t ws_bytes,
                    float p_drop, uint64_t seed, cudaStream_t stream);
  int train_backward(const float* dy, int block_hi, int block_lo, bool front, void* ws, int64_t ws_bytes, cudaStream_t stream);
  void train_drop_graphs();
  template <typename TA>
  int train_forward_impl(const float* mel, const int64_t* len_in, float* y, int64_t* len_out, int B, int Tin, Workspace& ws, float p_drop,
                         cudaStream_t stream);
  template <typename TA>
  int train_backward_impl(const float* dy, int block_hi, int block_lo, bool front, void* ws, int64_t ws_bytes, cudaStream_t stream);
  template <typename TA>
  int forward_impl(Ctx& c, int op, const float* x, const float* pos, const uint8_t* key_pad, const int64_t* len_in,
                   float* y, int64_t* len_out, int B, int frames);
};

}  // namespace cfm

struct cfm_module : public cfm::Module {};
