// extern "C" surface of libconformer_b200.so (declared in include/conformer_b200.h).
#include <cstring>
#include <new>

#include "module.h"

namespace cfm {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
}  // namespace cfm

namespace cfm { extern unsigned long long* g_gemm_trace; }
using namespace cfm;

namespace {
inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

int run(cfm_module_t m, int op, const float* x, const float* pos, const uint8_t* key_pad, const int64_t* len_in,
        float* y, int64_t* len_out, int B, int frames, void* ws, int64_t ws_bytes, void* stream) {
  if (m == nullptr) { set_error("null module handle"); return 1; }
  if (x == nullptr || y == nullptr) { set_error("null input/output pointer"); return 1; }
  return m->forward(op, x, pos, key_pad, len_in, y, len_out, B, frames, ws, ws_bytes, S(stream), nullptr);
}
}  // namespace

extern "C" {

int cfm_abi_version(void) { return CFM_ABI_VERSION; }
const char* cfm_last_error(void) { return g_error.c_str(); }

int cfm_device_check(void) {
  int dev = 0;
  CFM_CUDA_OK(cudaGetDevice(&dev));
  int major = 0;
  CFM_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CFM_REQUIRE(major == 10, "this library only carries sm_100a code (Blackwell B200)");
  return tc_init();
}

int cfm_module_create(int kind, const cfm_config_t* cfg, cfm_module_t* out) {
  CFM_REQUIRE(cfg != nullptr && out != nullptr, "null argument");
  CFM_REQUIRE(kind >= CFM_KIND_ENCODER && kind <= CFM_KIND_DECODER, "unknown module kind");
  cfm_module* m = new (std::nothrow) cfm_module();
  CFM_REQUIRE(m != nullptr, "out of host memory");
  m->kind = kind;
  m->cfg = *cfg;
  *out = m;
  return 0;
}

int cfm_module_set_param(cfm_module_t m, const char* name, const float* data, int64_t numel) {
  CFM_REQUIRE(m != nullptr && name != nullptr && data != nullptr && numel > 0, "bad parameter registration");
  m->params[name] = ParamRef{data, numel, false};
  m->finalized = false;
  return 0;
}

int cfm_module_set_grad(cfm_module_t m, const char* name, float* grad, int64_t numel) {
  CFM_REQUIRE(m != nullptr && name != nullptr && grad != nullptr && numel > 0, "bad gradient registration");
  m->grads[name] = ParamRef{grad, numel, false};
  m->train.bound = false;
  return 0;
}
int cfm_train_bind(cfm_module_t m, int compute) {
  CFM_REQUIRE(m != nullptr, "null module handle");
  return m->train_bind(compute);
}
int cfm_set_pdl(int32_t enable) {
  cfm::g_pdl.store(enable != 0 ? 1 : 0, std::memory_order_relaxed);
  return 0;
}
int cfm_train_set_graphs(cfm_module_t m, int32_t enable) {
  if (m == nullptr) return 1;
  m->train.graphs_on = enable != 0;
  if (!m->train.graphs_on) m->train_drop_graphs();
  return 0;
}
int cfm_train_graph_stats(cfm_module_t m, int64_t* out4) {
  if (m == nullptr || out4 == nullptr) return 1;
  out4[0] = m->train.n_direct;
  out4[1] = m->train.n_captured;
  out4[2] = m->train.n_replayed;
  out4[3] = m->train.n_failed;
  return 0;
}
int cfm_train_refresh(cfm_module_t m, void* stream) {
  CFM_REQUIRE(m != nullptr, "null module handle");
  return m->train_refresh(S(stream));
}
int64_t cfm_encoder_train_workspace_bytes(cfm_module_t m, int32_t batch, int32_t Tin) {
  if (m == nullptr) { set_error("null module handle"); return -1; }
  return m->train_workspace_bytes(batch, Tin);
}
int cfm_encoder_train_forward(cfm_module_t m, const float* mel, const int64_t* lengths_in, float* y, int64_t* lengths_out, int32_t batch,
                              int32_t Tin, float dropout_p, uint64_t seed, void* workspace, int64_t workspace_bytes, void* stream) {
  CFM_REQUIRE(m != nullptr, "null module handle");
  return m->train_forward(mel, lengths_in, y, lengths_out, batch, Tin, workspace, workspace_bytes, dropout_p, seed, S(stream));
}
int cfm_encoder_train_backward(cfm_module_t m, const float* dy, int32_t block_hi, int32_t block_lo, int32_t front_end, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  CFM_REQUIRE(m != nullptr && workspace != nullptr, "null argument");
  return m->train_backward(dy, block_hi, block_lo, front_end != 0, workspace, workspace_bytes, S(stream));
}

int cfm_module_finalize(cfm_module_t m, int compute, void* stream) {
  CFM_REQUIRE(m != nullptr, "null module handle");
  return m->finalize(compute, S(stream));
}

int cfm_module_destroy(cfm_module_t m) {
  delete m;
  return 0;
}

int cfm_padding_mask(const int64_t* lengths, int32_t batch, int32_t max_len, uint8_t* valid, void* stream) {
  CFM_REQUIRE(lengths != nullptr && valid != nullptr && batch > 0 && max_len > 0, "bad padding-mask arguments");
  return lengths_and_mask(lengths, nullptr, valid, batch, max_len, false, true, S(stream));
}

int cfm_ctc_greedy_decode(const float* logits, int32_t batch, int32_t T, int32_t vocab, int32_t pad_id, int32_t unk_id,
                          int32_t* tokens, int32_t* counts, void* stream) {
  CFM_REQUIRE(logits != nullptr && tokens != nullptr && counts != nullptr, "bad CTC decode arguments");
  return ctc_greedy_decode(logits, batch, T, vocab, pad_id, unk_id, tokens, counts, S(stream));
}

int cfm_rel_pos_encoding(const float* div_term, int32_t batch, int32_t T, int32_t d_model, float* out, void* stream) {
  CFM_REQUIRE(div_term != nullptr && out != nullptr && batch > 0, "bad positional-encoding arguments");
  CFM_TRY(rel_pos_table<float>(div_term, out, T, d_model, S(stream)));
  const size_t bytes = static_cast<size_t>(2 * T - 1) * d_model * sizeof(float);
  for (int b = 1; b < batch; ++b)                      // position.py:26 repeats the table over the batch
    CFM_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<char*>(out) + b * bytes, out, bytes, cudaMemcpyDeviceToDevice, S(stream)));
  return 0;
}

int cfm_swish(const float* x, float* y, int64_t numel, void* stream) {
  CFM_REQUIRE(x != nullptr && y != nullptr && numel > 0, "bad swish arguments");
  return swish_f32(x, y, numel, S(stream));
}
int cfm_glu(const float* x, float* y, int64_t rows, int32_t d_out, void* stream) {
  CFM_REQUIRE(x != nullptr && y != nullptr && rows > 0 && d_out > 0, "bad glu arguments");
  return glu_lastdim_f32(x, y, rows, d_out, S(stream));
}
int cfm_subsampled_lengths(const int64_t* lengths, int64_t* out, int32_t batch, void* stream) {
  CFM_REQUIRE(lengths != nullptr && out != nullptr && batch > 0, "bad lengths arguments");
  return lengths_and_mask(lengths, out, nullptr, batch, 0, true, false, S(stream));
}

int64_t cfm_workspace_bytes(cfm_module_t m, int32_t batch, int32_t frames) {
  if (m == nullptr) { set_error("null module handle"); return -1; }
  int64_t need = 0;
  int s = m->forward(m->kind, nullptr, nullptr, nullptr, reinterpret_cast<const int64_t*>(8), nullptr, nullptr, batch,
                     frames, nullptr, 0, nullptr, &need);
  return s == 0 ? need + 256 : -1;
}

int cfm_ffn_forward(cfm_module_t m, const float* x, float* y, int32_t batch, int32_t T, void* ws, int64_t wsb, void* stream) {
  return run(m, CFM_KIND_FFN, x, nullptr, nullptr, nullptr, y, nullptr, batch, T, ws, wsb, stream);
}
int cfm_mhsa_forward(cfm_module_t m, const float* x, const float* pos, const uint8_t* key_pad, float* y, int32_t batch,
                     int32_t T, void* ws, int64_t wsb, void* stream) {
  return run(m, CFM_KIND_MHSA, x, pos, key_pad, nullptr, y, nullptr, batch, T, ws, wsb, stream);
}
int cfm_conv_module_forward(cfm_module_t m, const float* x, float* y, int32_t batch, int32_t T, void* ws, int64_t wsb,
                            void* stream) {
  return run(m, CFM_KIND_CONV_MODULE, x, nullptr, nullptr, nullptr, y, nullptr, batch, T, ws, wsb, stream);
}
int cfm_block_forward(cfm_module_t m, const float* x, const float* pos, const uint8_t* key_pad, float* y, int32_t batch,
                      int32_t T, void* ws, int64_t wsb, void* stream) {
  return run(m, CFM_KIND_BLOCK, x, pos, key_pad, nullptr, y, nullptr, batch, T, ws, wsb, stream);
}
int cfm_subsampling_forward(cfm_module_t m, const float* mel, const int64_t* lengths_in, float* y, int64_t* lengths_out,
                            int32_t batch, int32_t Tin, void* ws, int64_t wsb, void* stream) {
  return run(m, CFM_KIND_SUBSAMPLING, mel, nullptr, nullptr, lengths_in, y, lengths_out, batch, Tin, ws, wsb, stream);
}
int cfm_encoder_forward(cfm_module_t m, const float* mel, const int64_t* lengths_in, float* y, int64_t* lengths_out,
                        int32_t batch, int32_t Tin, void* ws, int64_t wsb, void* stream) {
  return run(m, CFM_KIND_ENCODER, mel, nullptr, nullptr, lengths_in, y, lengths_out, batch, Tin, ws, wsb, stream);
}

int cfm_decoder_forward(cfm_module_t m, const float* x, const int64_t* lengths, float* logits, int32_t batch, int32_t T,
                        void* ws, int64_t wsb, void* stream) {
  return run(m, CFM_KIND_DECODER, x, nullptr, nullptr, lengths, logits, nullptr, batch, T, ws, wsb, stream);
}

int64_t cfm_encoder_host_workspace_bytes(cfm_module_t m, int32_t batch, int32_t Tin) {
  const int64_t inner = cfm_workspace_bytes(m, batch, Tin);
  if (inner < 0) return -1;
  const int64_t T2 = ((Tin - 1) / 2 - 1) / 2;
  const int64_t mel = static_cast<int64_t>(batch) * m->cfg.n_mel_channels * Tin * 4;
  const int64_t y = static_cast<int64_t>(batch) * T2 * m->cfg.d_model * 4;
  // fp32 staging of both tensors + their bf16 transfer copies (cfm_encoder_forward_host_io) + lengths
  return inner + mel + y + mel / 2 + y / 2 + 2 * (static_cast<int64_t>(batch) * 8 + 256) + 2048;
}

namespace {
// One host-buffer forward: copy-in on s_in, forward on s_fwd, copy-out on s_out (all three may be the same stream, in
// which case no events are needed).  bf16 transfer dtypes halve the PCIe bytes: a bf16 mel is widened on the device (the
// bf16 compute path rounds the mel to bf16 in its first kernel anyway, so nothing is lost there), a bf16 result is
// narrowed on the device before it leaves.
int host_forward(cfm_module_t m, const void* mel_host, int mel_dtype, const int64_t* lengths_in_host, void* y_host, int y_dtype,
                 int64_t* lengths_out_host, int32_t batch, int32_t Tin, void* workspace, int64_t workspace_bytes,
                 cudaStream_t s_fwd, cudaStream_t s_in, cudaStream_t s_out) {
  CFM_REQUIRE(m != nullptr && mel_host != nullptr && y_host != nullptr && workspace != nullptr, "null argument");
  CFM_REQUIRE((mel_dtype == CFM_DTYPE_F32 || mel_dtype == CFM_DTYPE_BF16) && (y_dtype == CFM_DTYPE_F32 || y_dtype == CFM_DTYPE_BF16),
              "host transfer dtype must be CFM_DTYPE_F32 or CFM_DTYPE_BF16");
  const int64_t T2 = ((Tin - 1) / 2 - 1) / 2;
  const int64_t mel_n = static_cast<int64_t>(batch) * m->cfg.n_mel_channels * Tin;
  const int64_t y_n = static_cast<int64_t>(batch) * T2 * m->cfg.d_model;
  const size_t len_b = static_cast<size_t>(batch) * 8;
  Workspace w;
  w.init(workspace, static_cast<size_t>(workspace_bytes), false);
  float* mel_d = static_cast<float*>(w.take(static_cast<size_t>(mel_n) * 4));
  float* y_d = static_cast<float*>(w.take(static_cast<size_t>(y_n) * 4));
  bf16* mel16_d = static_cast<bf16*>(w.take(static_cast<size_t>(mel_n) * 2));
  bf16* y16_d = static_cast<bf16*>(w.take(static_cast<size_t>(y_n) * 2));
  int64_t* li_d = static_cast<int64_t*>(w.take(len_b));
  int64_t* lo_d = static_cast<int64_t*>(w.take(len_b));
  char* rest = static_cast<char*>(w.take(256));
  CFM_REQUIRE(mel_d && y_d && mel16_d && y16_d && li_d && lo_d && rest, "workspace too small; query cfm_encoder_host_workspace_bytes");
  const int64_t rest_bytes = workspace_bytes - (rest - static_cast<char*>(workspace));
  const bool piped = s_in != s_fwd || s_out != s_fwd;
  Module::HostSlot* slot = nullptr;
  if (piped) {
    // the slot's events (created at its first use; waiting on an event that was never recorded is a no-op)
    slot = &m->host_slots[workspace];
    if (slot->copied_in == nullptr) {
      CFM_CUDA_OK(cudaEventCreateWithFlags(&slot->copied_in, cudaEventDisableTiming));
      CFM_CUDA_OK(cudaEventCreateWithFlags(&slot->computed, cudaEventDisableTiming));
      CFM_CUDA_OK(cudaEventCreateWithFlags(&slot->copied_out, cudaEventDisableTiming));
    }
    // copy-in: after the slot's previous forward has finished reading its staged input
    CFM_CUDA_OK(cudaStreamWaitEvent(s_in, slot->computed, 0));
  }
  if (mel_dtype == CFM_DTYPE_F32) CFM_CUDA_OK(cudaMemcpyAsync(mel_d, mel_host, static_cast<size_t>(mel_n) * 4, cudaMemcpyHostToDevice, s_in));
  else CFM_CUDA_OK(cudaMemcpyAsync(mel16_d, mel_host, static_cast<size_t>(mel_n) * 2, cudaMemcpyHostToDevice, s_in));
  if (lengths_in_host) CFM_CUDA_OK(cudaMemcpyAsync(li_d, lengths_in_host, len_b, cudaMemcpyHostToDevice, s_in));
  if (piped) {
    CFM_CUDA_OK(cudaEventRecord(slot->copied_in, s_in));
    // forward: after the copy-in, and after the slot's previous result has left its staging buffer
    CFM_CUDA_OK(cudaStreamWaitEvent(s_fwd, slot->copied_in, 0));
    CFM_CUDA_OK(cudaStreamWaitEvent(s_fwd, slot->copied_out, 0));
  }
  if (mel_dtype == CFM_DTYPE_BF16) CFM_TRY(cast_bf16_to_f32(mel16_d, mel_d, mel_n, s_fwd));
  CFM_TRY(m->forward(CFM_KIND_ENCODER, mel_d, nullptr, nullptr, lengths_in_host ? li_d : nullptr, y_d,
                     lengths_in_host ? lo_d : nullptr, batch, Tin, rest, rest_bytes, s_fwd, nullptr));
  if (y_dtype == CFM_DTYPE_BF16) CFM_TRY(cast_f32_to_bf16(y_d, y16_d, y_n, s_fwd));
  if (piped) {
    CFM_CUDA_OK(cudaEventRecord(slot->computed, s_fwd));
    CFM_CUDA_OK(cudaStreamWaitEvent(s_out, slot->computed, 0));      // copy-out: after the forward
  }
  if (y_dtype == CFM_DTYPE_F32) CFM_CUDA_OK(cudaMemcpyAsync(y_host, y_d, static_cast<size_t>(y_n) * 4, cudaMemcpyDeviceToHost, s_out));
  else CFM_CUDA_OK(cudaMemcpyAsync(y_host, y16_d, static_cast<size_t>(y_n) * 2, cudaMemcpyDeviceToHost, s_out));
  if (lengths_in_host && lengths_out_host)
    CFM_CUDA_OK(cudaMemcpyAsync(lengths_out_host, lo_d, len_b, cudaMemcpyDeviceToHost, s_out));
  if (piped) CFM_CUDA_OK(cudaEventRecord(slot->copied_out, s_out));
  return 0;
}
}  // namespace

int cfm_encoder_forward_host(cfm_module_t m, const float* mel_host, const int64_t* lengths_in_host, float* y_host,
                             int64_t* lengths_out_host, int32_t batch, int32_t Tin, void* workspace,
                             int64_t workspace_bytes, void* stream) {
  return host_forward(m, mel_host, CFM_DTYPE_F32, lengths_in_host, y_host, CFM_DTYPE_F32, lengths_out_host, batch, Tin, workspace,
                      workspace_bytes, S(stream), S(stream), S(stream));
}

int cfm_encoder_forward_host_pipelined(cfm_module_t m, const float* mel_host, const int64_t* lengths_in_host, float* y_host,
                                       int64_t* lengths_out_host, int32_t batch, int32_t Tin, void* workspace,
                                       int64_t workspace_bytes, void* stream, void* copy_in_stream, void* copy_out_stream) {
  return host_forward(m, mel_host, CFM_DTYPE_F32, lengths_in_host, y_host, CFM_DTYPE_F32, lengths_out_host, batch, Tin, workspace,
                      workspace_bytes, S(stream), S(copy_in_stream), S(copy_out_stream));
}

int cfm_encoder_forward_host_io(cfm_module_t m, const void* mel_host, int32_t mel_dtype, const int64_t* lengths_in_host,
                                void* y_host, int32_t y_dtype, int64_t* lengths_out_host, int32_t batch, int32_t Tin,
                                void* workspace, int64_t workspace_bytes, void* stream, void* copy_in_stream,
                                void* copy_out_stream) {
  return host_forward(m, mel_host, mel_dtype, lengths_in_host, y_host, y_dtype, lengths_out_host, batch, Tin, workspace,
                      workspace_bytes, S(stream), S(copy_in_stream ? copy_in_stream : stream),
                      S(copy_out_stream ? copy_out_stream : stream));
}

int cfm_encoder_host_slot_release(cfm_module_t m, void* workspace) {
  CFM_REQUIRE(m != nullptr, "null module handle");
  auto it = m->host_slots.find(workspace);
  if (it == m->host_slots.end()) return 0;
  // everything queued against the slot (copy-in, forward, copy-out) has finished once its last copy-out has
  if (it->second.copied_out) CFM_CUDA_OK(cudaEventSynchronize(it->second.copied_out));
  if (it->second.computed) CFM_CUDA_OK(cudaEventSynchronize(it->second.computed));
  if (it->second.copied_in) { CFM_CUDA_OK(cudaEventSynchronize(it->second.copied_in)); cudaEventDestroy(it->second.copied_in); }
  if (it->second.computed) cudaEventDestroy(it->second.computed);
  if (it->second.copied_out) cudaEventDestroy(it->second.copied_out);
  m->host_slots.erase(it);
  return 0;
}

int64_t cfm_last_launch_count(cfm_module_t m) { return m ? m->last_launches : -1; }

int cfm_profile_enable(cfm_module_t m, int32_t on) {
  CFM_REQUIRE(m != nullptr, "null module handle");
  m->prof.on = on != 0;
  m->prof.recs.clear();
  m->prof.next = 0;
  return 0;
}

int cfm_profile_collect(cfm_module_t m, double* out_host, int32_t n_classes) {
  CFM_REQUIRE(m != nullptr && out_host != nullptr && n_classes >= KC_COUNT, "bad profile buffer");
  for (int i = 0; i < n_classes * 4; ++i) out_host[i] = 0.0;
  for (const ProfRec& r : m->prof.recs) {
    CFM_CUDA_OK(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    CFM_CUDA_OK(cudaEventElapsedTime(&ms, r.e0, r.e1));
    double* o = out_host + r.cls * 4;
    o[0] += 1.0; o[1] += ms; o[2] += r.flops; o[3] += r.bytes;
  }
  m->prof.recs.clear();
  m->prof.next = 0;
  return 0;
}

int cfm_gemm_bf16(const void* A, const void* W, const float* bias, void* C, int32_t M, int32_t N, int32_t K, int32_t epi,
                  const float* res, float res_scale, void* stream) {
  CFM_REQUIRE(A && W && C && epi >= 0 && epi <= 5, "bad GEMM arguments");
  int dev = 0, sms = 0;
  CFM_CUDA_OK(cudaGetDevice(&dev));
  CFM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (epi == EPI_BIAS_RES) {
    CFM_REQUIRE(res != nullptr, "residual epilogue without residual");
    if (res != static_cast<const float*>(C))       // the kernel accumulates into C in place
      CFM_CUDA_OK(cudaMemcpyAsync(C, res, static_cast<size_t>(M) * N * 4, cudaMemcpyDeviceToDevice, S(stream)));
    res = static_cast<const float*>(C);
  }
  return gemm_bf16_tc(static_cast<const bf16*>(A), K, static_cast<const bf16*>(W), bias, C,
                      epi == EPI_BIAS_GLU ? N / 2 : N, M, N, K, static_cast<Epilogue>(epi), res, res_scale, nullptr, sms,
                      S(stream));
}
/* debug: device buffer receiving 8 globaltimer stamps per tile of CTA 0 of the next tcgen05 GEMMs (NULL = off) */
__attribute__((visibility("default"))) int cfm_debug_set_gemm_trace(void* dev_buf) {
  cfm::g_gemm_trace = static_cast<unsigned long long*>(dev_buf);
  return 0;
}

int cfm_gemm_bf16_tn(const void* A, const void* B, float* C, int32_t M, int32_t N, int32_t K, void* stream) {
  CFM_REQUIRE(A && B && C, "bad GEMM arguments");
  int dev = 0, sms = 0;
  CFM_CUDA_OK(cudaGetDevice(&dev));
  CFM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return gemm_bf16_tn(static_cast<const bf16*>(A), M, static_cast<const bf16*>(B), N, C, N, M, N, K, sms, S(stream));
}

int cfm_gemm_tf32(const float* A, const float* W, const float* bias, float* C, int32_t M, int32_t N, int32_t K, void* stream) {
  CFM_REQUIRE(A && W && C, "bad GEMM arguments");
  int dev = 0, sms = 0;
  CFM_CUDA_OK(cudaGetDevice(&dev));
  CFM_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return gemm_tf32_tc(A, K, W, K, bias, C, N, M, N, K, sms, S(stream));
}

int cfm_gemm_f32(const float* A, const float* W, const float* bias, float* C, int32_t M, int32_t N, int32_t K, int32_t epi,
                 const float* res, float res_scale, void* stream) {
  CFM_REQUIRE(A && W && C && epi >= 0 && epi <= 4, "bad GEMM arguments");
  return gemm_f32(A, K, W, bias, C, N, M, N, K, static_cast<Epilogue>(epi), res, res_scale, nullptr, S(stream));
}
int cfm_dwconv_glu_bn_swish(const void* u, const float* w, const float* bias, void* y, int32_t batch, int32_t T, int32_t d,
                            int32_t ksize, int32_t dtype, void* stream) {
  CFM_REQUIRE(u && w && bias && y && batch > 0 && T > 0 && d > 0, "bad depthwise-conv arguments");
  if (dtype == 0)
    return dwconv_glu_bn_swish<float>(static_cast<const float*>(u), w, bias, static_cast<float*>(y), batch, T, d, ksize, S(stream));
  return dwconv_glu_bn_swish<bf16>(static_cast<const bf16*>(u), w, bias, static_cast<bf16*>(y), batch, T, d, ksize, S(stream));
}

}  // extern "C"
