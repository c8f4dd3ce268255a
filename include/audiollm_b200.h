/* audiollm_b200 — C ABI of the B200 (sm_100a) audio-conditioning path.
 *
 * The reference (cdreetz/audio-llama) has no FFI: its seam is the Python module API under
 * /root/reference/src/models plus HuggingFace calls. Each entry point below names the reference
 * interface it replaces; the Python host side (audio_llama_b200/) binds them with ctypes and keeps the
 * reference's module / method names (INTEGRATION.md shows the binding).
 *
 * Conventions: every pointer is a DEVICE pointer unless the name ends in _host; every function is
 * asynchronous on `stream` (a cudaStream_t), returns 0 on success, -1 on an argument error, -2 on a CUDA
 * error (al_last_error() has the text), never throws and never allocates on the hot path (workspaces are
 * passed in; the only library-owned device memory is the constant tables built on first use per process:
 * FFT twiddles, Hann window, mel filter bank). Caller guarantees dtype, contiguity and 16-byte alignment.
 * bf16 = raw __nv_bfloat16 bits (torch.bfloat16).
 */
#ifndef AUDIOLLM_B200_H
#define AUDIOLLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* al_stream_t; /* cudaStream_t */

int al_version(void);
const char* al_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
long long al_launch_count(void);

/* ---- M1 / M2: log-mel features --------------------------------------------------------------------
 * Replaces WhisperFeatureExtractor.__call__ -> _torch_extract_fbank_features
 * (transformers/models/whisper/feature_extraction_whisper.py:135-164, 296-303) as called by
 * process_audio, /root/reference/src/inference.py:100-105 (mode 0), and the MelSpectrogram + log of
 * AudioLLMDataset._process_audio, /root/reference/src/dataset.py:125-133 (mode 1).
 *   wave       [n_clips][wave_stride] float32 mono 16 kHz
 *   n_samples  [n_clips] int32 valid samples per clip (NULL = wave_stride); clips are zero-padded /
 *              truncated to 480 000 samples exactly as the reference does
 *   out        [n_clips][n_mels][3000] float32
 *   clip_max_ws[n_clips] uint32 scratch (mode 0)
 */
int al_mel_forward(const float* wave, const int* n_samples, int n_clips, long long wave_stride, int n_mels,
                   int mode, float* out, unsigned int* clip_max_ws, al_stream_t stream);
/* flags: AL_MEL_RAW (mode 0) = stop before the per-clip floor: out holds log10(max(mel, 1e-10)) and clip_max_ws the
 * per-clip maximum (ordered-uint encoding); al_encoder_forward_ex / al_pack_mel_ex finish the job while they repack
 * the features, which saves the floor pass's launch and its re-read + re-write of the f32 features. */
enum { AL_MEL_RAW = 1 };
int al_mel_forward_ex(const float* wave, const int* n_samples, int n_clips, long long wave_stride, int n_mels,
                      int mode, int flags, float* out, unsigned int* clip_max_ws, al_stream_t stream);
/* Host-side copy of the filter bank the kernel uses, [201][n_mels] float64 row-major (for parity tests
 * against transformers.audio_utils.mel_filter_bank / torchaudio melscale_fbanks). */
int al_mel_filterbank_host(int n_mels, int mode, double* out_host);
/* Install a caller-built bank [201][n_mels] float64 for (n_mels, mode) before its first use. The host side uses
 * this for mode 1: torchaudio builds the HTK bank with float32 torch ops whose last-ulp behaviour moves the very
 * narrow low filters by ~1e-3 relative, so the bank is rebuilt with the same torch ops and handed in. */
int al_mel_set_filterbank_host(int n_mels, int mode, const double* fb_host);
/* Kernel form of al_mel_forward: 1 = tensor-core folded DFT (fp16 hi/lo operands, fp32 accumulation; default when
 * the filter bank is banded), 0 = CUDA-core FFT. Same results within the mel tolerance; the environment variable
 * AUDIOLLM_B200_MEL=tc|fft sets the default. */
int al_mel_set_mode(int tc);

/* ---- waveform ingest ----------------------------------------------------------
 * Channel mean + sinc resampling to `new_freq` + zero padding / truncation to out_cap samples, one launch.
 * Replaces torch.mean(waveform, dim=0) and torchaudio.transforms.Resample(orig_freq, new_freq) of
 * /root/reference/src/inference.py:87-98 and /root/reference/src/dataset.py:105-123 (torchaudio's default
 * sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99).
 *   in     [n_clips][n_chan][...] f32: clip b channel c sample s at in[b*clip_stride + c*chan_stride + s]
 *   n_in   [n_clips] int32 valid samples per clip (NULL = n_in_cap); every clip is first cut to n_in_cap samples
 *          (the training path truncates BEFORE resampling, dataset.py:106-112; pass a large cap otherwise)
 *   out    [n_clips][out_stride] f32, samples >= n_out[b] zero — the layout al_mel_forward reads
 *   n_out  [n_clips] int32 = min(ceil(new*len/orig), out_cap), or NULL */
int al_ingest_forward(const float* in, long long clip_stride, long long chan_stride, int n_chan, const int* n_in,
                      int n_in_cap, int orig_freq, int new_freq, float* out, long long out_stride, int out_cap,
                      int* n_out, int n_clips, al_stream_t stream);

/* ---- building blocks ------------------------------------------------------------------------------ */
/* epilogue flags for al_gemm_bf16* */
#define AL_EPI_GELU 1
#define AL_EPI_OUT_F32 2
#define AL_EPI_REDUCE_ADD 4
#define AL_EPI_ROWAUX 8
#define AL_EPI_RESIDUAL 16

/* out = epilogue(A @ W^T + bias).  A [batch][m_per_batch][K] bf16 with row / batch strides in ELEMENTS
 * (rows may overlap — the conv stem uses that), W [N][K] bf16 (nn.Linear layout), bias [N] f32 or NULL,
 * out [batch][m_per_batch][N] bf16 (or f32 with AL_EPI_OUT_F32), aux [m_per_batch][aux_ld] f32, resid (with
 * AL_EPI_RESIDUAL) f32 with the output's strides — it may BE the output buffer (in-place x += ...).
 * Replaces nn.Linear / nn.Conv1d forward (cuBLAS / cuDNN in the reference's stack). */
int al_gemm_bf16(const void* A, long long a_row_stride, long long a_batch_stride, int m_per_batch, int batch,
                 const void* W, int N, int K, const float* bias, void* out, long long o_row_stride,
                 long long o_batch_stride, int flags, const float* aux, int aux_ld, const float* resid,
                 al_stream_t stream);

/* Weight-gradient form: out[M][N] (f32, row pitch ldo) += A_src^T W_src with A_src [K][lda] (M valid columns) and W_src
 * [K][ldw] (N valid columns) bf16, both row-major over the contraction index (what autograd computes for nn.Linear's
 * weight: grad_out^T x). The operands are staged MN-major (64-column swizzle atoms, a_major = b_major = 1 in the tcgen05
 * instruction descriptor), nothing is transposed in memory; split-K with TMA reduce-add, so `out` must be initialised. */
int al_gemm_tn_accumulate(const void* A_src, long long lda, int M, const void* W_src, long long ldw, int N, int K, float* out,
                          long long ldo, al_stream_t stream);

/* GEMM kernel form: 1 = CTA pair per 256x256 tile (tcgen05.mma.cta_group::2, default), 0 = one CTA per 128x256
 * tile. Same results; the environment variable AUDIOLLM_B200_GEMM=single|pair sets the default. */
int al_gemm_set_mode(int pair);

/* LayerNorm(x[rows][d] f32) -> out (out_dtype 0 bf16 / 1 f32). Output row of input row r:
 * (r / rows_per_group) * out_group_stride + out_row_offset + r % rows_per_group, rows of out_ld elements. */
int al_layernorm(const float* x, const float* gamma, const float* beta, void* out, int rows, int d, float eps,
                 int out_dtype, long long out_ld, int rows_per_group, long long out_group_stride,
                 long long out_row_offset, al_stream_t stream);

/* softmax(Q K^T) V, no mask, head_dim 64, q pre-scaled. qkv [B][T][3*H*64] bf16 -> out [B][T][H*64] bf16.
 * Replaces WhisperAttention's attention_interface call (modeling_whisper.py:339-349).
 * al_attention_ex flags: AL_ATT_Q_LOG2 = q carries log2(e) as well as head_dim^-1/2 (scores arrive in log2 units:
 * the kernel then takes exp2 straight from the accumulator; same result). al_attention = flags 0. */
enum { AL_ATT_Q_LOG2 = 1 };
int al_attention(const void* qkv, void* out, int B, int T, int H, al_stream_t stream);
int al_attention_ex(const void* qkv, void* out, int B, int T, int H, int flags, al_stream_t stream);

int al_pack_mel(const float* mel, void* out_bf16, int B, int n_mels, int T, int c_pad, al_stream_t stream);
/* clip_max_ws != NULL: mel is the AL_MEL_RAW output of al_mel_forward_ex and the extractor's floor + affine step is
 * applied while packing (bit-identical to al_mel_forward followed by al_pack_mel). */
int al_pack_mel_ex(const float* mel, const unsigned int* clip_max_ws, void* out_bf16, int B, int n_mels, int T,
                   int c_pad, al_stream_t stream);
int al_f32_to_bf16(const float* x, void* out_bf16, long long n, al_stream_t stream);

/* ---- E1/E2: frozen Whisper encoder forward --------------------------------------------------------
 * Replaces AudioLLM._process_audio_features (/root/reference/src/models/allm.py:198-221) ->
 * WhisperEncoder.forward (transformers/models/whisper/modeling_whisper.py:593-647).
 * Weights are bf16 matrices packed by the host (audio_llama_b200/encoder.py): conv weights as
 * [d][3*c_in] with column kk*c_in+ci, wqkv = [q*hd^-0.5 ; k ; v] ([3d][d]), biases / LayerNorm / position
 * table in f32. The plan keeps pointers only; the caller owns weights and workspace. */
typedef struct al_encoder al_encoder;
size_t al_encoder_workspace_bytes(int d_model, int n_layers, int n_heads, int ffn_dim, int n_mels, int max_batch);
int al_encoder_create(al_encoder** out, int d_model, int n_layers, int n_heads, int ffn_dim, int n_mels,
                      int max_batch, void* workspace, size_t workspace_bytes);
int al_encoder_set_stem(al_encoder* e, const void* conv1_w, const float* conv1_b, const void* conv2_w,
                        const float* conv2_b, const float* pos, const float* lnf_g, const float* lnf_b);
int al_encoder_set_layer(al_encoder* e, int layer, const float* ln1_g, const float* ln1_b, const void* wqkv,
                         const float* bqkv, const void* wo, const float* bo, const float* ln2_g,
                         const float* ln2_b, const void* w1, const float* b1, const void* w2, const float* b2);
/* attention_flags: AL_ATT_* the plan passes to its attention launches (AL_ATT_Q_LOG2 when the host packed
 * wqkv's q rows and bias with hd^-0.5 * log2(e), as audio_llama_b200/encoder.py does). Default 0. */
int al_encoder_set_options(al_encoder* e, int attention_flags);
/* mel [B][n_mels][3000] f32 -> out [B][1500][d] (out_dtype 0 bf16 / 1 f32). n_layers_run < 0 = all. */
int al_encoder_forward(al_encoder* e, const float* mel, int B, void* out, int out_dtype, int n_layers_run,
                       al_stream_t stream);
/* The same with mel = the AL_MEL_RAW output of al_mel_forward_ex and its clip_max_ws: the per-clip floor and the
 * (x + 4) / 4 step run inside the plan's first kernel (the f32 mel is written once and read once). */
int al_encoder_forward_ex(al_encoder* e, const float* mel, const unsigned int* clip_max_ws, int B, void* out,
                          int out_dtype, int n_layers_run, al_stream_t stream);
/* Live per-kernel timing of the plan's launches (CUDA events recorded on the launch stream around every
 * kernel while profiling is on). al_encoder_profile_read synchronises, sums milliseconds and launch counts per
 * kind since the last read / set, and resets. Kinds index ms_by_kind_host[AL_K_COUNT]. */
enum { AL_K_PACK = 0, AL_K_CONV1, AL_K_CONV2, AL_K_LN, AL_K_QKV, AL_K_ATTN, AL_K_OPROJ, AL_K_FC1, AL_K_FC2, AL_K_COUNT };
int al_encoder_set_profiling(al_encoder* e, int on);
int al_encoder_profile_read(al_encoder* e, float* ms_by_kind_host, int* launches_by_kind_host);
/* fp32 residual stream [max_batch*1500][d] inside the workspace (tests read intermediate states). */
float* al_encoder_hidden(al_encoder* e);
int al_encoder_destroy(al_encoder* e);

/* ---- P1: AudioProjector forward -------------------------------------------------------------------
 * Replaces AudioProjector.forward (/root/reference/src/models/projector.py:18-19):
 * LN(W2 gelu(W1 x + b1) + b2). x [rows][d_in] bf16; W1 [hidden][d_in], W2 [d_out][hidden] bf16; h_ws
 * [rows][hidden] bf16 and y_ws [rows][d_out] f32 scratch. The LayerNorm stores with al_layernorm's row
 * mapping, i.e. straight into inputs_embeds[b, 1 + t] when out = inputs_embeds,
 * rows_per_group = 1500, out_group_stride = S, out_row_offset = 1. */
int al_projector_forward(const void* x, int rows, int d_in, int hidden, int d_out, const void* W1, const float* b1,
                         const void* W2, const float* b2, const float* gamma, const float* beta, void* h_ws,
                         float* y_ws, void* out, int out_dtype, long long out_ld, int rows_per_group,
                         long long out_group_stride, long long out_row_offset, al_stream_t stream);

/* Backward of the projector (it is the trainable part of the path). Inputs: the forward's operands, the saved
 * h = gelu(W1 x + b1) (bf16 [rows][hidden]) and y = W2 h + b2 (f32 [rows][d_out], the LayerNorm input), and the
 * upstream gradient dout (f32 [rows][d_out]). Outputs (f32, overwritten): dW1 [hidden][d_in], db1, dW2
 * [d_out][hidden], db2, dgamma, dbeta. x has no gradient (the encoder is frozen, base.py:8-9).
 * LayerNorm backward kernel -> dy; dW2 = dy^T h and dW1 = da^T x as split-K tcgen05 GEMMs on transposed (K-major)
 * copies with TMA reduce-add; dh = dy W2; da = dh * gelu'(.) with the pre-activation recomputed in the GEMM epilogue. */
size_t al_projector_backward_workspace_bytes(int rows, int d_in, int hidden, int d_out);
int al_projector_backward(const void* x, int rows, int d_in, int hidden, int d_out, const void* W1, const float* b1,
                          const void* W2, const float* gamma, const void* h_saved, const float* y_saved,
                          const float* dout, void* workspace, float* dW1, float* db1, float* dW2, float* db2,
                          float* dgamma, float* dbeta, al_stream_t stream);

/* ---- L1: frozen linear + LoRA update ---------------------------------------------------------------
 * Replaces lora_forward_hook(module, input, output, lora_layer) = output + (x @ (B @ A).T) * scaling
 * (/root/reference/src/models/lora.py:20-21, 41-43) together with the frozen nn.Linear it hooks:
 *   out = x W^T + bias + (x A^T)(s B)^T
 * as two launches: T = x A^T ([rows][rank] bf16, t_ws) and one GEMM whose K loop runs over in_dim and then over
 * rank, both into the same TMEM accumulator — the dense [out][in] delta of the reference is never formed.
 * x [rows][in] bf16, W [out][in] bf16, bias [out] f32 or NULL, lora_A [rank][in] bf16, lora_B_scaled = scaling*B
 * [out][rank] bf16 (rank a multiple of 8), out [rows][out] bf16 (out_dtype 0) or f32 (1). */
int al_lora_linear_forward(const void* x, int rows, int in_dim, int out_dim, int rank, const void* W,
                           const float* bias, const void* lora_A, const void* lora_B_scaled, void* t_ws, void* out,
                           int out_dtype, al_stream_t stream);

/* Backward of al_lora_linear_forward with W frozen (what autograd derives from lora.py:20-21, 41-43 plus the hooked
 * nn.Linear): U = dy (sB); dx = dy W + U A (one GEMM, second operand pair in the K loop; NULL dx skips it);
 * dA = U^T x [rank][in] f32; dB_raw = dy^T T [out][rank] f32 with T = x A^T saved by the forward (t_ws) — the
 * gradient of the UNSCALED lora_B is scaling * dB_raw, and dA already carries the scaling through U.
 * W_T is the frozen weight transposed, [in][out] bf16 (transposed once by the caller, it never changes). */
/* The fused GEMM's LoRA operands from the fp32 parameters of lora.py's LoRALayer in one launch: a_pad [round8(rank)][in]
 * bf16 = A (padding rows zero), b_scaled_pad [out][round8(rank)] bf16 = scaling * B (padding columns zero). */
int al_lora_pack(const float* lora_A, const float* lora_B, int rank, int in_dim, int out_dim, float scaling, void* a_pad,
                 void* b_scaled_pad, al_stream_t stream);
size_t al_lora_linear_backward_workspace_bytes(int rows, int in_dim, int out_dim, int rank);
/* The _ex forms fold an elementwise add into the GEMM epilogue (what HF's LlamaDecoderLayer.forward writes as
 * `residual + hidden_states`, and what autograd does when several projections share one input):
 *   al_lora_linear_forward_ex   out = x W^T + b + (x A^T)(sB)^T + addend      addend [rows][out_dim] bf16 or NULL
 *   al_lora_linear_backward_ex  dx  = dy W + U A + dx_addend                  dx_addend [rows][in_dim] bf16 or NULL
 *   al_linear_add_bf16          out = x W^T + b + addend                      (a frozen linear without LoRA: o_proj)
 * The addend may be the output buffer itself (accumulate in place). */
int al_lora_linear_forward_ex(const void* x, int rows, int in_dim, int out_dim, int rank, const void* W,
                              const float* bias, const void* lora_A, const void* lora_B_scaled, void* t_ws,
                              const void* addend, void* out, int out_dtype, al_stream_t stream);
int al_lora_linear_backward_ex(const void* x, const void* dy, int rows, int in_dim, int out_dim, int rank, const void* W_T,
                               const void* lora_A, const void* lora_B_scaled, const void* t_saved, void* workspace,
                               const void* dx_addend, void* dx, float* dA, float* dB_raw, al_stream_t stream);
int al_linear_add_bf16(const void* x, int rows, int in_dim, int out_dim, const void* W, const float* bias,
                       const void* addend, void* out, al_stream_t stream);
int al_lora_linear_backward(const void* x, const void* dy, int rows, int in_dim, int out_dim, int rank, const void* W_T,
                            const void* lora_A, const void* lora_B_scaled, const void* t_saved, void* workspace,
                            void* dx, float* dA, float* dB_raw, al_stream_t stream);

/* ---- LLaMA-side row kernels (first slice) -------------------------------------------------
 * The elementwise / row work of the HF LlamaForCausalLM the reference drives (/root/reference/src/models/allm.py:99-104
 * -> HF models/llama/modeling_llama.py), bf16 in / out, fp32 arithmetic, forward and backward:
 *   al_rmsnorm_*      LlamaRMSNorm.forward: y = weight * bf16(x * rsqrt(mean(x^2) + eps)); rstd [rows] f32 saved for the
 *                     backward (weight frozen: dx only)
 *   al_swiglu_*       act_fn(gate_proj(x)) * up_proj(x) of LlamaMLP.forward (SiLU)
 *   al_rope           apply_rotary_pos_emb on x [B][S][H][head_dim] with cos / sin [cos_batch][S][head_dim]
 *                     (cos_batch 1 or B); backward = 1 applies the transposed rotation to a gradient
 *   al_cross_entropy_inplace   rows of logits [rows][ld] bf16 -> loss_sum += lse - logit[label] (label -100 ignored),
 *                     the row is overwritten with (softmax - onehot) * grad_scale
 *   al_linear_ce      lm_head + the loss of LlamaForCausalLM.forward without materialising [rows][vocab] logits: per
 *                     chunk_rows rows, logits = h W^T, cross-entropy in place, dh = dlogits W (W_T = W transposed,
 *                     [d][round8(vocab)] bf16). labels are the SHIFTED labels (labels[t + 1], -100 at the end). */
int al_rmsnorm_forward(const void* x, const void* weight, void* y, float* rstd, int rows, int d, float eps, al_stream_t stream);
int al_rmsnorm_backward(const void* x, const void* weight, const float* rstd, const void* dy, void* dx, int rows, int d,
                        al_stream_t stream);
/* dx = rmsnorm backward + dx_addend (bf16 [rows][d] or NULL; may be dx itself): the gradient arriving over the residual
 * connection joins in the same pass. */
int al_rmsnorm_backward_ex(const void* x, const void* weight, const float* rstd, const void* dy, const void* dx_addend,
                           void* dx, int rows, int d, al_stream_t stream);
int al_swiglu_forward(const void* gate, const void* up, void* h, long long n, al_stream_t stream);
int al_swiglu_backward(const void* gate, const void* up, const void* dh, void* dgate, void* dup, long long n, al_stream_t stream);
int al_rope(const void* x, const void* cos, const void* sin, void* out, int B, int S, int H, int head_dim, int cos_batch,
            int backward, al_stream_t stream);
int al_cross_entropy_inplace(void* logits, const long long* labels, int rows, int vocab, long long ld, float grad_scale,
                             float* loss_sum, al_stream_t stream);
size_t al_linear_ce_workspace_bytes(int chunk_rows, int vocab);
int al_linear_ce(const void* h, const void* W, const void* W_T, const long long* labels, int rows, int d, int vocab,
                 float grad_scale, int chunk_rows, void* workspace, float* loss_sum, void* dh, al_stream_t stream);

/* ---- §8 f-1: causal grouped-query attention of the LLaMA layers (head_dim 128) ---------------------
 * Replaces the attention_interface call of LlamaAttention.forward (transformers/models/llama/modeling_llama.py),
 * which the reference reaches through self.llama.model(inputs_embeds=..., attention_mask=..., labels=...)
 * (/root/reference/src/models/allm.py:99-104), together with what autograd derives from it.
 *   q [B][S][Hq][128], k / v [B][S][Hkv][128] bf16 (q, k after the rotary embedding), out [B][S][Hq][128] bf16,
 *   lse [B][Hq][S] f32 (base-2 log-sum-exp of the scaled scores; saved for the backward),
 *   kv_len [B] int32 or NULL: keys at positions >= kv_len[b] are masked (a right-padded attention_mask).
 * softmax(scale * q k^T + causal mask + key-padding mask) v, query head h reading kv head h / (Hq / Hkv). */
int al_gqa_attention_forward(const void* q, const void* k, const void* v, void* out, float* lse, const int* kv_len,
                             int B, int S, int Hq, int Hkv, int head_dim, float scale, al_stream_t stream);
/* Backward of the above (what autograd derives from scaled_dot_product_attention in the reference's stack): given the
 * forward's operands, its out and lse, and d_out [B][S][Hq][128] bf16, writes dq / dk / dv (bf16, layouts of q / k / v;
 * dk / dv summed over the query heads of each group). dsum_ws: [B][Hq][S] f32 scratch. Scores are recomputed. */
int al_gqa_attention_backward(const void* q, const void* k, const void* v, const void* out, const float* lse,
                              const void* d_out, const int* kv_len, void* dq, void* dk, void* dv, float* dsum_ws,
                              int B, int S, int Hq, int Hkv, int head_dim, float scale, al_stream_t stream);

/* ---- S1 / S2: splice ------------------------------------------------------------------------------
 * Replaces AudioLLM._combine_text_and_audio_embeddings (allm.py:143-170), _extend_attention_mask
 * (allm.py:176-196) and the label extension (allm.py:81-89). Row map per sample (bit-exact contract):
 * 0 <- table[start_id]; 1..n_audio <- audio rows; n_audio+1 <- table[end_id]; n_audio+2+j <- table[ids[j]].
 *   table [vocab][d] (elem_bytes 2 or 4), input_ids / attn_mask / labels [B][t_txt] int64 (mask, labels may
 *   be NULL), audio_rows [B][n_audio][d] or NULL when the projector already wrote them in place,
 *   out [B][n_audio+2+t_txt][d]; mask_out float32, labels_out int64 (either may be NULL).
 * vocab = rows of `table`. Delimiter ids >= vocab are refused (-1; the reference's ValueError, allm.py:140-141).
 * An input id outside [0, vocab) -- on which the reference's embed_tokens(input_ids) raises (allm.py:64) -- is never
 * used as an address: its output row is zeroed and *bad_id_flag (device int32, may be NULL) is OR-ed with 1; the
 * host wrapper turns the flag into the exception. */
int al_splice(const void* table, int elem_bytes, int d, const long long* input_ids, const long long* attn_mask,
              const long long* labels, int B, int t_txt, int n_audio, long long start_id, long long end_id,
              const void* audio_rows, void* out, float* mask_out, long long* labels_out, long long vocab,
              int* bad_id_flag, al_stream_t stream);
/* Ragged extension (config 5; not in the reference): see audio_llama_b200/splice.py. span_start_out
 * [B][max_spans] int32 receives the device-computed exclusive prefix sums. */
int al_splice_ragged(const void* table, int elem_bytes, int d, const long long* input_ids,
                     const long long* attn_mask, const long long* labels, int B, int t_txt, int S_out,
                     const int* span_rows, const int* span_src_row, const int* n_spans, int max_spans,
                     const void* audio_rows, long long start_id, long long end_id, void* out, float* mask_out,
                     long long* labels_out, int* span_start_out, long long vocab, int* bad_id_flag,
                     al_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AUDIOLLM_B200_H */
