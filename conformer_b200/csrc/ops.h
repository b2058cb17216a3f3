// Internal launcher declarations (host side).  Every function enqueues work on `stream`,
// returns 0 on success and a non-zero status with cfm::set_error() otherwise.  No launcher
// allocates or synchronises.
#pragma once
#include "common.cuh"

namespace cfm {

enum Epilogue {
  EPI_BIAS = 0,        // C = acc + bias
  EPI_BIAS_SWISH = 1,  // C = swish(acc + bias)
  EPI_BIAS_RELU = 2,   // C = relu(acc + bias)
  EPI_BIAS_RES = 3,    // C(fp32) = res + res_scale * (acc + bias)
  EPI_BIAS_F32 = 4,    // C(fp32) = acc + bias
  EPI_BIAS_GLU = 5,    // C[:, c] = (acc+bias)[a(c)] * sigmoid((acc+bias)[g(c)]), N/2 outputs; weight rows packed
                       // per 64-row group as 32 value rows then their 32 gate rows (tensor-core path only)
};

// Implicit-im2col description of the A operand for the second subsampling conv
// (reference: Conv2d(d,d,3,stride 2), model/utils/convolution.py:39,46).
// The source is conv1's output stored channels-last as [B][T1p][F1][C]; GEMM row m = (b, t2, f2),
// GEMM k = (kw*3 + kh)*C + c  ->  src[b][2*t2+kw][2*f2+kh][c].
struct ConvGeom {
  int B, T1p, F1, C, T2, F2;
};

// ---- fp32 SIMT path (true fp32 accumulate; the rtol-1e-4 parity path) -------------------------
int gemm_f32(const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int M, int N,
             int K, Epilogue epi, const float* res, float res_scale, const ConvGeom* conv,
             cudaStream_t stream);

// ---- bf16 tcgen05/TMEM/TMA path -----------------------------------------------------------------
// A [M,K] bf16 row-major (lda elements, multiple of 8), W [N,K] bf16 row-major.  Output bf16 for
// EPI_BIAS/SWISH/RELU, fp32 for EPI_BIAS_RES/EPI_BIAS_F32.  When conv != nullptr, A is conv1's output and the
// output rows are ordered (b, t2, f2).
int gemm_bf16_tc(const bf16* A, int lda, const bf16* W, const float* bias, void* C, int ldc, int M, int N,
                 int K, Epilogue epi, const float* res, float res_scale, const ConvGeom* conv, int num_sms,
                 cudaStream_t stream, bool split_k = false);
// split_k (EPI_BIAS_RES only): when the output has few tiles and K is long - the weight-gradient GEMMs of the training
// step - the contraction is split over the idle SMs, every split reduce-adding its partial product into C (summation
// order then varies from run to run, so the inference path keeps it off)
// C[M,N] (fp32) += A^T B with A [K rows, M cols] (lda) and B [K rows, N cols] (ldb) bf16 row-major: the weight-gradient
// form dW += dY^T X, contraction over the rows, operands consumed in place as MN-major tiles (no transposed copies);
// split over K when the output has few tiles
int gemm_bf16_tn(const bf16* A, int lda, const bf16* B, int ldb, float* C, int ldc, int M, int N, int K, int num_sms,
                 cudaStream_t stream);
// fp32 operands multiplied as TF32 on tcgen05 (kind::tf32), fp32 accumulate and output: C = A[M,K] W[N,K]^T + bias.
// For fp32-grade accuracy pass TF32-exact high/low splits concatenated along K (A' = [Ahi|Alo|Ahi], W' = [Whi|Whi|Wlo]).
int gemm_tf32_tc(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M, int N, int K,
                 int num_sms, cudaStream_t stream);
int tc_init();   // resolves cuTensorMapEncodeTiled through the runtime; 0 when usable
// 128-byte-swizzled tiled tensor map (call tc_init() first)
int encode_tmap(CUtensorMap* tm, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                const cuuint64_t* strides, const cuuint32_t* box, const char* what);

// Fused FeedForwardModule half-step on the tensor cores (d_model = 256):
//   x[M,256] += scale * (swish(a_pre W1^T + b1) W2^T + b2)        (ffn.py:15-23, block.py:19,25)
// a_pre = LN(x_in) as bf16 [M,d], normalised by the producer of x_in (layer_norm_dual, the fused conv module's epilogue,
// or a layer_norm<bf16> launch).  The [M,1024] hidden activation never leaves the SM.  x must already hold the residual
// (or zeros); the kernel only adds into it (TMA reduction stores) and never reads it.
int ffn_fused_bf16(const bf16* a_pre, float* x, const bf16* W1, const float* b1, const bf16* W2, const float* b2, float scale,
                   int M, int d, int num_sms, cudaStream_t stream);
bool ffn_fused_supported(int d);

// Fused ConvolutionModule on the tensor cores (d_model = 256, kernel size 31):
//   out = res + pw2(swish(bn(dw(glu(pw1(LN(x)))))))        (convolution.py:21-32, block.py:23)
// W1p/b1p packed by pack_glu_rows, dw_w/dw_b BN-folded (fold_dwconv_bn).  res may be null or alias x; out must
// not alias x (tiles read a 15-frame halo of their neighbours).
int conv_module_fused_bf16(const float* x, const float* res, float* out, const float* ln_g, const float* ln_b,
                           const bf16* W1p, const float* b1p, const float* dw_w, const float* dw_b, const bf16* W2,
                           const float* b2, const float* ln2_g, const float* ln2_b, bf16* a_out /*nullable: LN(out; ln2) as bf16*/,
                           int B, int Tn, int d, int ksize, cudaStream_t stream);
bool conv_module_fused_supported(int d, int ksize);

// ---- CTC head (model/modules/decoder.py:15-27; decoder.cu) -------------------------------------
// one LSTM layer over packed sequences: xin [B,T,4H] = x W_ih^T + b_ih + b_hh (gate order i|f|g|o), out [B,T,H] = h_t (or
// swish(h_t) for the last layer), zeros at frames >= lengths[b]; h_state [2,B,H] and c_state [B,H] are scratch
template <typename TOut>
int lstm_layer_forward(const float* xin, const float* w_hh, const int64_t* lengths, float* h_state, float* c_state, TOut* out,
                       int B, int T, int H, bool last_layer, int num_sms, cudaStream_t stream);
// tensor-core recurrence (lstm_tc.cu, bf16 mode): W_ih / W_hh rows in the packed order, see lstm_tc.cu
bool lstm_tc_supported(int H);
int lstm_pack_rows_f32(const float* src, float* dst, int H, int K, cudaStream_t stream);
int lstm_pack_rows_bf16(const float* src, bf16* dst, int H, int K, cudaStream_t stream);
size_t lstm_tc_scratch_bytes(int H);
// xin_tm / out_tm are TIME-MAJOR: [T, B, 4H] and [T, B, H]
template <typename TOut>
int lstm_layer_forward_tc(const float* xin_tm, const bf16* w_hh_packed, const int64_t* lengths, void* scratch, TOut* out_tm, int B, int T,
                          int H, bool last_layer, int num_sms, cudaStream_t stream);
int lstm_to_time_major(const float* x, bf16* y, int B, int T, int d, cudaStream_t stream);
int lstm_logits_from_time_major(const float* lg, float* y, int B, int T, int V, int Vp, cudaStream_t stream);
int fold_bn_linear(const float* w, const float* b, const float* gamma, const float* beta, const float* mean, const float* var,
                   float eps, float* w_out /*[Vp,H]*/, float* b_out /*[Vp]*/, int V, int Vp, int H, cudaStream_t stream);
int add_vec(const float* a, const float* b, float* o, int n, cudaStream_t stream);

// ---- greedy CTC decode (processing/processor.py:301-328, id level; ctc.cu) ----------------------
// tokens int32 [B,T]: collapsed ids, -1 after counts[b] entries
int ctc_greedy_decode(const float* logits /*[B,T,V]*/, int B, int T, int V, int pad_id, int unk_id, int32_t* tokens,
                      int32_t* counts, cudaStream_t stream);

// ---- normalisation / elementwise ---------------------------------------------------------------
template <typename TOut>
int layer_norm(const float* x, const float* gamma, const float* beta, TOut* y, int64_t rows, int d,
               cudaStream_t stream);
// y = LN(x; g1, b1) (fp32) and a = LN(y; g2, b2) (bf16) in one pass (d = 256): block-final LayerNorm plus the next
// block's ffn_1 LayerNorm
int layer_norm_dual(const float* x, const float* g1, const float* b1, float* y, const float* g2, const float* b2, bf16* a,
                    int64_t rows, int d, cudaStream_t stream);

// depthwise conv over time fused with GLU (on load), folded BatchNorm and Swish.
// u [B,T,2d] (pointwise_conv_1 output, value half then gate half), w [ksize][d] (BN-folded,
// tap-major), bias [d] (BN-folded), y [B,T,d].
template <typename T>
int dwconv_glu_bn_swish(const T* u, const float* w, const float* bias, T* y, int B, int Tn, int d, int ksize,
                        cudaStream_t stream);
// bf16 variant whose input already went through GLU (fused into the pointwise_conv_1 GEMM epilogue):
// g [B,T,d] -> y [B,T,d]; 16-byte loads, packed fp32x2 FMAs, time halo in shared memory.
int dwconv_bn_swish_bf16(const bf16* g, const float* w, const float* bias, bf16* y, int B, int Tn, int d, int ksize,
                         cudaStream_t stream);
// reorder the 2d rows of pointwise_conv_1 (weight [2d,K] or bias with K=1) for EPI_BIAS_GLU
int pack_glu_rows(const float* w, float* out, int d, int K, cudaStream_t stream);

// relative-position multi-head self-attention core (model/utils/attention.py:47-72).
// qkv [B,T,3d] (q | k | v, head h at columns h*dh), pos [2T-1, ldp] (projected table, head h at
// columns h*dh), u/v = content/position bias [H,dh] fp32, key_pad [B,T] bytes (1 = masked) or null.
template <typename T>
int rel_attention(const T* qkv, const T* pos, int ldp, const float* u, const float* v,
                  const uint8_t* key_pad, T* ctx, int B, int Tn, int H, int dh, cudaStream_t stream);

// bf16 tcgen05/TMEM flash variant (S, shifted positional scores and P.V on 5th-gen tensor cores), head size 64
int rel_attention_bf16_tc(const bf16* qkv, const bf16* pos, int ldp, const float* u, const float* v,
                          const uint8_t* key_pad, bf16* ctx, int B, int Tn, int H, int dh, cudaStream_t stream);

// bf16 tcgen05 variant for d_model 256 / head size 64 with the sinusoidal positional term factorised into a rotated
// query times an absolute key table (no relative shift, no projected [2T-1,d] table; attention_fact.cu).
// wpT = pack_pos_fact(pos_proj.weight); e16 (bf16) / r16 (fp16) = pos_fact_tables(div_term): [T][d] = cos(t w_k) | sin(t w_k).
// The query block of qkv must come from a projection packed by pack_q_fact (c (q + u), c = log2(e)/sqrt(dh));
// pos_const = pos_const_fact(...) = c (v - u) Wp per head.
int rel_attention_bf16_fact(const bf16* qkv, const bf16* wpT, const bf16* e16, const void* r16, const float* pos_const,
                            const uint8_t* key_pad, bf16* ctx, int B, int Tn, int H, int dh, cudaStream_t stream);
int pack_q_fact(float* wq /*[d,d] in place*/, float* bq /*[d] in place*/, const float* u, int d, int dh, cudaStream_t stream);
int pos_const_fact(const float* wp /*[d,d]*/, const float* u, const float* v, float* out /*[H][d]*/, int d, int dh,
                   cudaStream_t stream);
bool rel_attention_fact_supported(int d, int dh);
int pos_fact_tables(const float* div_term, bf16* e16, void* r16, int Tn, int d, cudaStream_t stream);
int pack_pos_fact(const float* wp /*[d,d]*/, bf16* out /*[H][d][dh]*/, int d, int dh, cudaStream_t stream);

// samples whose keys are all masked get ctx = mean over time of V, as the reference's uniform softmax gives
// (attention.py:60-66); a no-op for every sample with at least one valid key
template <typename T>
int all_masked_mean_v(const uint8_t* key_pad, const T* qkv /*[B,T,3d]*/, T* ctx /*[B,T,d]*/, int B, int Tn, int d,
                      cudaStream_t stream);

// sinusoid table [2T-1, d] (model/utils/position.py:11-25), row p <-> relative position T-1-p
template <typename T>
int rel_pos_table(const float* div_term, T* pe, int Tn, int d, cudaStream_t stream);

// first subsampling conv: x [B,n_mel,Tin] fp32 -> out [B][T1p][F1][C] channels-last, ReLU fused
template <typename T>
int subsample_conv1(const float* x, const float* w, const float* bias, T* out, int B, int n_mel, int Tin,
                    int T1, int T1p, int F1, int C, cudaStream_t stream);

// tensor-core variant for C = 256 (bf16 out; the 3x3 patch is one K=16 MMA step): w_packed = pack_conv1_weight(w)
int subsample_conv1_tc(const float* mel, const bf16* w_packed, const float* bias, bf16* out, int B, int n_mel, int Tin, int T1,
                       int T1p, int F1, int C, int num_sms, cudaStream_t stream);
int pack_conv1_weight(const float* w /*[C,1,3,3]*/, bf16* out /*[C][64]*/, int C, cudaStream_t stream);
bool conv1_tc_supported(int C);

// lengths' = ((L-1)//2-1)//2 ; key_pad[b,t] = t >= lengths'[b]   (convolution.py:55, masking.py:9-13)
// (mark_valid = true writes the complement, i.e. generate_padding_mask itself)
int lengths_and_mask(const int64_t* lengths_in, int64_t* lengths_out, uint8_t* key_pad, int B, int Tn,
                     bool subsample, bool mark_valid, cudaStream_t stream);

// elementwise helpers used by the module-level API and weight packing
int swish_f32(const float* x, float* y, int64_t n, cudaStream_t stream);
int glu_lastdim_f32(const float* x, float* y, int64_t rows, int d, cudaStream_t stream);
int cast_f32_to_bf16(const float* x, bf16* y, int64_t n, cudaStream_t stream);
int cast_bf16_to_f32(const bf16* x, float* y, int64_t n, cudaStream_t stream);
// weight packing
int pack_conv2_weight(const float* w /*[C,C,3,3]*/, float* out /*[C][(kw*3+kh)*C+c]*/, int C, cudaStream_t);
int pack_linear_weight(const float* w /*[d, C*F2] (c*F2+f)*/, float* out /*[d][f*C+c]*/, int d, int C, int F2,
                       cudaStream_t);
int fold_dwconv_bn(const float* w /*[d,1,k]*/, const float* b, const float* bn_w, const float* bn_b,
                   const float* rm, const float* rv, float eps, float* w_out /*[k][d]*/, float* b_out, int d,
                   int k, cudaStream_t);

}  // namespace cfm
