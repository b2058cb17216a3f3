/* conformer_b200 - C ABI of the B200 (sm_100a) Conformer encoder forward path.
 *
 * The reference (Alan-404/Conformer) is pure PyTorch and has no FFI: its boundary for this
 * path is the nn.Module API.  This header is the C-ABI that the drop-in
 * modules in conformer_b200/modules.py bind with ctypes; each entry point names the reference
 * interface it stands in for.  INTEGRATION.md shows the binding a maintainer of the reference
 * would add.
 *
 * Conventions
 *   - every pointer argument is a DEVICE pointer unless its name ends in _host;
 *   - tensors are dense row-major fp32 unless stated; `lengths` are int64 like torch.long;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - functions return 0 on success; on failure they return non-zero and cfm_last_error()
 *     (thread-local) describes it.  Nothing throws across the ABI, nothing allocates device
 *     memory inside a forward call (the caller supplies `workspace`), nothing synchronises;
 *   - a module handle owns device copies of its (packed) weights; it is tied to the CUDA device
 *     current at cfm_module_finalize (which must also be current at every forward) and may be
 *     used from one host thread at a time; different handles may be used from different threads
 *     and on different devices of one process;
 *   - aliasing: a forward's output `y` may alias its input `x` for cfm_ffn_forward only; every
 *     other forward needs distinct buffers.  `workspace` must not overlap any argument;
 *   - masks: key_pad marks padded KEYS (1 = masked).  A sample whose keys are all masked gets the
 *     reference's result (uniform softmax = mean of V, attention.py:60-66) from the module-level
 *     forwards (cfm_mhsa_forward, cfm_block_forward).  cfm_encoder_forward builds its masks from
 *     `lengths` and requires every utterance to keep at least one output frame (lengths >= 7 mel
 *     frames), which the reference's collate guarantees for real audio;
 *   - there is no CPU path: without a CUDA device every compute entry point fails.
 */
#ifndef CONFORMER_B200_H_
#define CONFORMER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFM_ABI_VERSION 2

#if defined(__GNUC__)
#define CFM_API __attribute__((visibility("default")))
#else
#define CFM_API
#endif

typedef struct cfm_module* cfm_module_t;

/* Arithmetic of a finalized module. */
enum {
  CFM_COMPUTE_F32 = 0,  /* fp32 storage, fp32 FMA on CUDA cores (parity path, rtol 1e-4)            */
  CFM_COMPUTE_BF16 = 1  /* bf16 GEMM operands on tcgen05 tensor cores, fp32 accumulate,             */
                        /* fp32 residual stream / LayerNorm / softmax (= the reference under        */
                        /* torch.autocast, train.py:232, with bf16 in place of fp16)                */
};

/* Element type of a HOST transfer buffer (cfm_encoder_forward_host_io). */
enum { CFM_DTYPE_F32 = 0, CFM_DTYPE_BF16 = 1 };

/* Which reference module a handle stands for (constructor arguments in cfm_config_t). */
enum {
  CFM_KIND_ENCODER = 0,      /* Encoder                      model/modules/encoder.py:9-37      */
  CFM_KIND_BLOCK = 1,        /* ConformerBlock               model/utils/block.py:8-29          */
  CFM_KIND_FFN = 2,          /* FeedForwardModule            model/utils/ffn.py:5-23            */
  CFM_KIND_MHSA = 3,         /* MultiHeadSelfAttentionModule model/utils/attention.py:7-18      */
  CFM_KIND_REL_ATTENTION = 4,/* RelativeMultiHeadAttention   model/utils/attention.py:20-102    */
  CFM_KIND_CONV_MODULE = 5,  /* ConvolutionModule            model/utils/convolution.py:7-32    */
  CFM_KIND_SUBSAMPLING = 6,  /* ConvolutionSubsampling       model/utils/convolution.py:34-57   */
  CFM_KIND_DECODER = 7       /* Decoder (CTC head)           model/modules/decoder.py:7-27      */
};

typedef struct {
  int32_t n_mel_channels; /* Encoder(n_mel_channels=...)            encoder.py:10 */
  int32_t n_blocks;       /* Encoder(n_blocks=...)                                */
  int32_t d_model;        /* d_model / dim / channels of every module             */
  int32_t n_heads;        /* n_heads                                              */
  int32_t kernel_size;    /* ConvolutionModule(kernel_size=...)  convolution.py:8 */
  int32_t vocab_size;     /* Decoder(vocab_size, d_model, hidden_dim, n_layers)   decoder.py:8; 0 for the other kinds */
  int32_t lstm_hidden_dim;
  int32_t n_lstm_layers;
} cfm_config_t;

CFM_API int cfm_abi_version(void);
CFM_API const char* cfm_last_error(void);
/* 0 when a CUDA device with compute capability 10.x is current and usable. */
CFM_API int cfm_device_check(void);

/* ---- module lifecycle --------------------------------------------------------------------- */
CFM_API int cfm_module_create(int kind, const cfm_config_t* cfg, cfm_module_t* out);
/* Register one parameter or buffer under the name it has in the reference module's state_dict
 * (e.g. "layers.3.conv.batch_norm.running_var"); `data` is fp32 on the device and must stay valid
 * until cfm_module_finalize returns.  Integer buffers (num_batches_tracked) are not registered. */
CFM_API int cfm_module_set_param(cfm_module_t m, const char* name, const float* data, int64_t numel);
/* Check that every tensor of the reference state_dict was registered with the right size, then
 * build the device-side packed weights (BatchNorm folded into the depthwise conv, q/k/v fused,
 * conv2/linear reordered to channels-last, bf16 copies).  May be called again after new
 * cfm_module_set_param calls (e.g. after load_state_dict). */
CFM_API int cfm_module_finalize(cfm_module_t m, int compute, void* stream);
CFM_API int cfm_module_destroy(cfm_module_t m);

/* ---- stateless operators ------------------------------------------------------------------- */
/* generate_padding_mask (model/utils/masking.py:4-13): valid[b,t] = t < lengths[b], uint8 [B,max_len] */
CFM_API int cfm_padding_mask(const int64_t* lengths, int32_t batch, int32_t max_len, uint8_t* valid, void* stream);
/* RelativePositionalEncoding.forward (model/utils/position.py:11-28): out fp32 [batch, 2T-1, d];
 * div_term fp32 [d/2] (the module's non-trainable parameter).  batch >= 1 replicates like :26. */
CFM_API int cfm_rel_pos_encoding(const float* div_term, int32_t batch, int32_t T, int32_t d_model, float* out, void* stream);
/* ConformerProcessor.greedy_decode at token-id level (processing/processor.py:301-328), the step right after the
 * path: ids = argmax over the vocabulary (first maximum), frames whose id is pad_id (the CTC blank,
 * <PAD> = 0, processor.py:114) or unk_id (last vocabulary entry) are skipped WITHOUT resetting the previous id, an id is
 * emitted when it differs from the previously emitted one.  Every row is decoded over the full padded length T, as the
 * reference does.  logits fp32 [batch, T, vocab]; tokens int32 [batch, T] receives the collapsed ids followed by -1;
 * counts int32 [batch] their number. */
CFM_API int cfm_ctc_greedy_decode(const float* logits, int32_t batch, int32_t T, int32_t vocab, int32_t pad_id,
                                  int32_t unk_id, int32_t* tokens, int32_t* counts, void* stream);
/* Swish.forward / GLU.forward over the last dimension (model/utils/activation.py:4-17) */
CFM_API int cfm_swish(const float* x, float* y, int64_t numel, void* stream);
CFM_API int cfm_glu(const float* x, float* y, int64_t rows, int32_t d_out, void* stream);
/* ConvolutionSubsampling length rule ((L-1)//2-1)//2 (convolution.py:55) */
CFM_API int cfm_subsampled_lengths(const int64_t* lengths, int64_t* out, int32_t batch, void* stream);

/* ---- module forwards ------------------------------------------------------------------------ */
/* Bytes of scratch a forward of `m` needs for `batch` sequences of `frames` frames (for the
 * encoder and the subsampling module `frames` counts INPUT mel frames, otherwise time steps). */
CFM_API int64_t cfm_workspace_bytes(cfm_module_t m, int32_t batch, int32_t frames);

/* FeedForwardModule.forward: y[rows,d] = out_linear(swish(hidden_linear(LN(x)))) */
CFM_API int cfm_ffn_forward(cfm_module_t m, const float* x, float* y, int32_t batch, int32_t T, void* workspace,
                    int64_t workspace_bytes, void* stream);
/* MultiHeadSelfAttentionModule.forward (KIND_MHSA: LayerNorm first) or
 * RelativeMultiHeadAttention.forward with q=k=v=x (KIND_REL_ATTENTION).
 * pos: fp32 [2T-1, d] positional table (one copy; the reference's batch replicas are identical),
 *      or NULL to build it from the module-independent default div_term.
 * key_pad: uint8 [B,T], 1 = masked key (the reference's mask[:,0,0,:]), or NULL. */
CFM_API int cfm_mhsa_forward(cfm_module_t m, const float* x, const float* pos, const uint8_t* key_pad, float* y,
                     int32_t batch, int32_t T, void* workspace, int64_t workspace_bytes, void* stream);
/* ConvolutionModule.forward (eval-mode BatchNorm) */
CFM_API int cfm_conv_module_forward(cfm_module_t m, const float* x, float* y, int32_t batch, int32_t T, void* workspace,
                            int64_t workspace_bytes, void* stream);
/* ConformerBlock.forward(x, pos_embedding, mask) */
CFM_API int cfm_block_forward(cfm_module_t m, const float* x, const float* pos, const uint8_t* key_pad, float* y,
                      int32_t batch, int32_t T, void* workspace, int64_t workspace_bytes, void* stream);
/* ConvolutionSubsampling.forward: mel fp32 [B,n_mel,Tin] -> y fp32 [B,T',d*F2] with feature index
 * c*F2+f (convolution.py:51-52); lengths_in/lengths_out may both be NULL. */
CFM_API int cfm_subsampling_forward(cfm_module_t m, const float* mel, const int64_t* lengths_in, float* y,
                            int64_t* lengths_out, int32_t batch, int32_t Tin, void* workspace,
                            int64_t workspace_bytes, void* stream);
/* Encoder.forward(x, lengths) -> (y [B,T',d] fp32, lengths').  lengths_in NULL = no mask. */
CFM_API int cfm_encoder_forward(cfm_module_t m, const float* mel, const int64_t* lengths_in, float* y,
                        int64_t* lengths_out, int32_t batch, int32_t Tin, void* workspace,
                        int64_t workspace_bytes, void* stream);
/* Decoder.forward(x, lengths) in eval mode (model/modules/decoder.py:15-27), the step right after the path: x fp32
 * [B,T,d_model] (the encoder output), lengths int64 [B] or NULL (sorted or not; lengths[b] <= T, and like the
 * reference's pad_packed_sequence the output keeps T frames when max(lengths) == T) -> logits fp32 [B,T,vocab].
 * Frames at or beyond a sample's length carry the constant Linear(BatchNorm(0)), as pad_packed_sequence + the head give. */
CFM_API int cfm_decoder_forward(cfm_module_t m, const float* x, const int64_t* lengths, float* logits, int32_t batch,
                                int32_t T, void* workspace, int64_t workspace_bytes, void* stream);
/* Same call with HOST buffers (pinned or pageable): copies mel/lengths in, runs the forward and
 * copies y/lengths' out on `stream`, using workspace for the device staging.  This is the
 * end-to-end entry point a non-torch caller would use. */
CFM_API int cfm_encoder_forward_host(cfm_module_t m, const float* mel_host, const int64_t* lengths_in_host, float* y_host,
                             int64_t* lengths_out_host, int32_t batch, int32_t Tin, void* workspace,
                             int64_t workspace_bytes, void* stream);
CFM_API int64_t cfm_encoder_host_workspace_bytes(cfm_module_t m, int32_t batch, int32_t Tin);
/* Pipelined form for serving loops: the copy-in runs on `copy_in_stream`, the forward on `stream`, the copy-out on
 * `copy_out_stream`, ordered by events the handle keeps per SLOT (a slot is identified by its `workspace`, sized by
 * cfm_encoder_host_workspace_bytes): a slot's copy-in waits for its previous forward, its forward for that copy-in and
 * for its previous copy-out, its copy-out for the forward.  Rotating K slots over fewer compute streams keeps the
 * compute streams free of copies (the reference's loop does .cuda() / .cpu() around every forward, test.py:145-149).
 * y_host (and lengths_out_host) are valid once the work queued on `copy_out_stream` has completed. */
CFM_API int cfm_encoder_forward_host_pipelined(cfm_module_t m, const float* mel_host, const int64_t* lengths_in_host,
                                               float* y_host, int64_t* lengths_out_host, int32_t batch, int32_t Tin,
                                               void* workspace, int64_t workspace_bytes, void* stream,
                                               void* copy_in_stream, void* copy_out_stream);

/* General form of the two calls above.  mel_dtype / y_dtype select what crosses PCIe: CFM_DTYPE_F32 (the reference's
 * tensors, test.py:145-149) or CFM_DTYPE_BF16 (half the bytes; the bf16 compute path rounds the mel to bf16 in its first
 * kernel anyway, so a bf16 mel loses nothing there; a bf16 result is the fp32 result rounded once).  copy_in_stream /
 * copy_out_stream may be NULL (= `stream`, i.e. the unpipelined form). */
CFM_API int cfm_encoder_forward_host_io(cfm_module_t m, const void* mel_host, int32_t mel_dtype,
                                        const int64_t* lengths_in_host, void* y_host, int32_t y_dtype,
                                        int64_t* lengths_out_host, int32_t batch, int32_t Tin, void* workspace,
                                        int64_t workspace_bytes, void* stream, void* copy_in_stream, void* copy_out_stream);
/* Forget a pipelined slot before its workspace is freed or replaced: waits for the work queued against it and drops
 * its events. */
CFM_API int cfm_encoder_host_slot_release(cfm_module_t m, void* workspace);

/* Number of kernel launches the last forward on `m` enqueued (for bench.py's gpu_launches). */
CFM_API int64_t cfm_last_launch_count(cfm_module_t m);

/* Optional per-launch profile: when enabled every kernel launch of a forward on `m` is bracketed by
 * CUDA events on the launching stream.  cfm_profile_collect waits for them and fills, per kernel class
 * (0 tcgen05 GEMM, 1 fp32 GEMM, 2 LayerNorm, 3 attention, 4 depthwise conv, 5 conv1, 6 misc, 7 fused FFN, 8 fused conv module), four
 * doubles: launches, total milliseconds, algorithmic FLOPs, algorithmic bytes. */
CFM_API int cfm_profile_enable(cfm_module_t m, int32_t on);
CFM_API int cfm_profile_collect(cfm_module_t m, double* out_host, int32_t n_classes);

/* ---- training step of the encoder (BASELINE config 4; train.py:232-243) ---------------------- */
/* The handle keeps POINTERS to the live fp32 master parameters (cfm_module_set_param, which must then stay valid) and to
 * one fp32 gradient buffer per parameter (cfm_module_set_grad; the backward ACCUMULATES into them, the caller zeroes
 * them).  cfm_train_bind checks the reference's state_dict contract and allocates the derived copies; cfm_train_refresh
 * re-derives them (bf16 / transposed / packed) and must be called after every optimizer step, before the next forward.
 * Buffers (running_mean / running_var) are registered like parameters and are UPDATED IN PLACE by the training forward
 * (momentum 0.1, unbiased variance), as nn.BatchNorm1d does; num_batches_tracked stays with the caller. */
CFM_API int cfm_module_set_grad(cfm_module_t m, const char* name, float* grad, int64_t numel);
CFM_API int cfm_train_bind(cfm_module_t m, int compute);
CFM_API int cfm_train_refresh(cfm_module_t m, void* stream);
/* Programmatic dependent launch of the forward kernels (process-wide, default: on).  Each kernel of the path is launched
 * with programmatic stream serialization: its CTAs become resident and set up (barriers, tensor memory, descriptors)
 * while the previous kernel's last wave drains, then wait (griddepcontrol.wait) until that kernel has completed and its
 * memory is visible before they touch any operand.  Results are identical to plain stream order; 0 turns it off. */
CFM_API int cfm_set_pdl(int32_t enable);
/* CUDA graphs of the training step (default: on).  A training forward, a block range of the backward or the weight
 * refresh that is called a second time with the same arguments (pointers, shapes, dropout rate) is captured and from
 * then on replayed with one graph launch on the caller's stream; the dropout seed is passed through device memory, so
 * it may change from step to step.  Any other argument that changes makes that call launch directly (and is remembered
 * in turn).  cfm_train_bind discards all graphs.  Buffers named in a captured call (input, lengths, output, workspace,
 * output gradient) must stay allocated while the handle may replay it. */
CFM_API int cfm_train_set_graphs(cfm_module_t m, int32_t enable);
/* out4 = {calls launched directly, sequences captured, graph replays, captures that failed (those keys stay direct)} */
CFM_API int cfm_train_graph_stats(cfm_module_t m, int64_t* out4);
/* One buffer holds the tape (activations kept for the backward) followed by scratch. */
CFM_API int64_t cfm_encoder_train_workspace_bytes(cfm_module_t m, int32_t batch, int32_t Tin);
/* Encoder.forward(x, lengths) in TRAINING mode: dropout with rate dropout_p at every nn.Dropout of the path (stateless
 * masks from `seed`, recomputed by the backward), BatchNorm1d with batch statistics over all B*T' frames. */
CFM_API int cfm_encoder_train_forward(cfm_module_t m, const float* mel, const int64_t* lengths_in, float* y, int64_t* lengths_out,
                                      int32_t batch, int32_t Tin, float dropout_p, uint64_t seed, void* workspace,
                                      int64_t workspace_bytes, void* stream);
/* Backward of the blocks [block_lo, block_hi) in reverse order, then (front_end != 0, block_lo == 0) of Encoder.linear and
 * the two subsampling convolutions.  dy fp32 [B,T',d] = gradient of the encoder output, passed with the FIRST call of a
 * step (block_hi == n_blocks) and NULL afterwards: the gradient of the residual stream stays in the workspace between
 * calls, so a caller can all-reduce the gradients of finished blocks while the next call computes. */
CFM_API int cfm_encoder_train_backward(cfm_module_t m, const float* dy, int32_t block_hi, int32_t block_lo, int32_t front_end,
                                       void* workspace, int64_t workspace_bytes, void* stream);

/* ---- log-mel front-end: the step right before the path ---------------------- */
/* ConformerProcessor.mel_spectrogram (processing/processor.py:53-63,155-158):
 *   log(clamp(torchaudio.transforms.MelSpectrogram(...)(signal), min=1e-5))
 * i.e. centred reflect-padded STFT, periodic Hann window, power spectrum, triangular mel filterbank.  Both contractions
 * (DFT, filterbank) run on the tcgen05 tensor cores as 3xTF32 GEMMs with fp32 accumulation.  The field defaults of the
 * reference are {16000, 400, 400, 160, 80, 0.0, 8000.0, 1, 1, 1e-5}. */
typedef struct {
  int32_t sample_rate, n_fft, win_length, hop_length, n_mels;
  float f_min, f_max;               /* f_max <= 0: sample_rate / 2 */
  int32_t norm_slaney;              /* 1: norm="slaney" (area normalisation), 0: none */
  int32_t mel_scale_slaney;         /* 1: mel_scale="slaney", 0: "htk" */
  float log_floor;                  /* clamp before the log (1e-5, processor.py:157) */
} cfm_logmel_config_t;
typedef struct cfm_logmel* cfm_logmel_t;
/* window_host [n_fft] and fb_host [n_fft/2+1, n_mels] (HOST pointers, nullable) are the module's constant buffers as
 * torchaudio holds them (`spectrogram.window`, `mel_scale.fb`): torchaudio builds them in fp32, so a caller that wants
 * the reference's exact filters passes its own; NULL = computed here in fp64 from the configuration. */
CFM_API int cfm_logmel_create(const cfm_logmel_config_t* cfg, const float* window_host, const float* fb_host,
                              cfm_logmel_t* out);
CFM_API int cfm_logmel_destroy(cfm_logmel_t h);
CFM_API int64_t cfm_logmel_workspace_bytes(cfm_logmel_t h, int32_t batch, int32_t samples);
/* wave fp32 [batch, samples] (zero-padded to the longest utterance, processor.py:376-388) -> mel fp32
 * [batch, n_mels, samples / hop_length + 1], the tensor Encoder.forward takes. */
CFM_API int cfm_logmel_forward(cfm_logmel_t h, const float* wave, float* mel, int32_t batch, int32_t samples,
                               void* workspace, int64_t workspace_bytes, void* stream);
/* frames = samples // hop_length + 1 (processor.py:392) */
CFM_API int cfm_logmel_lengths(const int64_t* samples, int64_t* frames, int32_t batch, int32_t hop_length, void* stream);

/* ---- low-level kernels exposed for unit tests and benchmarks -------------------------------- */
/* bf16 tcgen05 GEMM: C = epi(A[M,K] W[N,K]^T + bias); epi: 0 bias, 1 bias+swish, 2 bias+relu,
 * 3 res + scale*(acc+bias), 4 acc+bias (C,res fp32; otherwise C bf16), 5 GLU: C is [M,N/2] and W's rows are
 * packed per 64-row group as 32 value rows followed by their 32 gate rows.  A,W bf16 (uint16 storage). */
CFM_API int cfm_gemm_bf16(const void* A, const void* W, const float* bias, void* C, int32_t M, int32_t N, int32_t K,
                  int32_t epi, const float* res, float res_scale, void* stream);
/* bf16 tcgen05 GEMM in the weight-gradient form: C fp32 [M,N] += A[K,M]^T B[K,N] (row-major bf16 operands read in place as
 * MN-major tiles; split over K when the output has few tiles, so the summation order may vary between runs). */
CFM_API int cfm_gemm_bf16_tn(const void* A, const void* B, float* C, int32_t M, int32_t N, int32_t K, void* stream);
/* TF32 tcgen05 GEMM (kind::tf32): C fp32 [M,N] = A fp32 [M,K] W fp32 [N,K]^T + bias; operands are read as TF32 (10-bit
 * mantissa).  For fp32-grade accuracy pass operands split into a TF32-exact high part and the remainder, concatenated
 * along K: A' = [Ah | Al | Ah], W' = [Wh | Wh | Wl] (the log-mel front-end does). */
CFM_API int cfm_gemm_tf32(const float* A, const float* W, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                          void* stream);
/* fp32 CUDA-core GEMM with the same epilogues (all fp32). */
CFM_API int cfm_gemm_f32(const float* A, const float* W, const float* bias, float* C, int32_t M, int32_t N, int32_t K,
                 int32_t epi, const float* res, float res_scale, void* stream);
/* depthwise stage of the ConvolutionModule on its own: u [B,T,2d] -> y [B,T,d]
 * (GLU, depthwise conv with BN-folded weights w[k][d] and bias[d], Swish); dtype 0 fp32, 1 bf16. */
CFM_API int cfm_dwconv_glu_bn_swish(const void* u, const float* w, const float* bias, void* y, int32_t batch, int32_t T,
                            int32_t d, int32_t ksize, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CONFORMER_B200_H_ */
