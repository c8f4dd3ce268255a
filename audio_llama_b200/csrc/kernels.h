// Internal launch interface between api.cu (the C ABI) and the kernel translation units.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace al {

// ------------------------------------------------------------------ GEMM (gemm_sm100.cu)
enum GemmEpilogue : int {
  EPI_GELU = 1,         // exact-erf GELU after the bias
  EPI_OUT_F32 = 2,      // fp32 output (default bf16)
  EPI_REDUCE_ADD = 4,   // out += result (TMA reduce-add; residual connection on the fp32 stream)
  EPI_ROWAUX = 8,       // + aux[row_in_batch][col] (fp32; Whisper's position table)
  EPI_RESIDUAL = 16,    // + resid[batch][row][col] (fp32, may alias out: in-place residual add, no atomics)
  EPI_GELU_GRAD = 32,   // out = grad_in[row][col] * gelu'(acc + bias)  (backward of a Linear+GELU, recomputed)
  EPI_TN = 64,          // operand layout, not an epilogue: out = A_src^T W_src with A_src [K][M], W_src [K][N] row-major
  EPI_ADD_BF16 = 128,   // + addend[row][col] (bf16 rows of grad_ld elements through grad_in; may alias a bf16 out: residual
                        //   adds and gradient accumulation of the LLaMA layers without a separate elementwise pass); the addend's tensor map is
                        //   passed to launch_gemm2_add, its chunks arrive through TMA
                        // (contraction over the ROWS: weight gradients); both operands are staged MN-major
};

struct GemmParams {
  int m_per_batch;
  int batch;
  int N;
  int K;
  int K2;                  // extra contraction length of the second operand pair (LoRA rank), 0 = none
  int tiles_m_per_batch;   // filled by launch_gemm
  int tiles_n;             // filled by launch_gemm
  const float* bias;       // [N] or nullptr
  const float* aux;        // [m_per_batch][aux_ld] or nullptr
  int aux_ld;
  const float* resid;      // [batch][m_per_batch][resid_ld] or nullptr (EPI_RESIDUAL)
  long long resid_ld;
  long long resid_batch_stride;
  const __nv_bfloat16* grad_in;   // [batch*m_per_batch][grad_ld] bf16 (EPI_GELU_GRAD)
  long long grad_ld;
  int k_splits;            // > 1: split the contraction over this many work items (needs EPI_REDUCE_ADD, zeroed out)
};

void gemm_set_mode(int pair);        // 1 = CTA-pair (cta_group::2) kernel, 0 = single-CTA kernel
int gemm_out_box_cols(int flags);   // inner box extent of the output tensor map (32 fp32 / 64 bf16)
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, GemmParams p, int flags,
                int num_sms, cudaStream_t stream);
// out = epilogue(A W^T + A2 W2^T + bias): second operand pair over p.K2, same accumulator (fused LoRA update).
int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmA2,
                 const CUtensorMap& tmB2, GemmParams p, int flags, int num_sms, cudaStream_t stream);
int launch_gemm2_add(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmA2,
                     const CUtensorMap& tmB2, const CUtensorMap& tmAdd, GemmParams p, int flags, int num_sms, cudaStream_t stream);

// ------------------------------------------------------------------ attention (attention_sm100.cu)
// qkv: [B][T][3*H*64] bf16 (q pre-scaled), out: [B][T][H*64] bf16. tm_qkv: 3-D map, box {64,128,1}, SW128.
// q_log2 != 0: q carries log2(e) as well as head_dim^-1/2 (scores in log2 units; the encoder folds both into Wq).
int launch_attention(const CUtensorMap& tm_qkv, const void* qkv, void* out, int B, int T, int H, int q_log2, cudaStream_t stream);

// ------------------------------------------------------------------ causal GQA attention, head_dim 128 (gqa_attention_sm100.cu)
// Tensor maps: rank 4 over [B][S][H][128] bf16, box {64, 1, 128, 1}, SW128 (see api.cu tmap_bshd).
int launch_gqa_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, float* lse,
                   const int* kv_len, int B, int S, int Hq, int Hkv, float scale, cudaStream_t stream);
// dsum_ws: [B][Hq][S] f32 scratch (rowsum(dO * O)); dq / dk / dv in the layouts of q / k / v.
int launch_gqa_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo, const CUtensorMap& tdq,
                   const void* out,
                   const void* d_out, const float* lse, float* dsum_ws, const int* kv_len, void* dq, void* dk, void* dv, int B,
                   int S, int Hq, int Hkv, float scale, cudaStream_t stream);

// ------------------------------------------------------------------ row kernels (rowwise.cu)
// LayerNorm over the last dim of fp32 rows. out_dtype 0 = bf16, 1 = fp32. Output row of input row r is
// (r / rows_per_group) * out_group_stride + out_row_offset + (r % rows_per_group)   (in rows of out_ld elements),
// which is how the projector's LayerNorm writes straight into inputs_embeds.
int launch_layernorm(const float* x, const float* gamma, const float* beta, void* out, int rows, int d, float eps,
                     int out_dtype, long long out_ld, int rows_per_group, long long out_group_stride,
                     long long out_row_offset, cudaStream_t stream);
// mel [B][n_mels][T] fp32 -> time-major bf16 [B][T+2][c_pad], rows 0 and T+1 and channels >= n_mels zero.
int launch_pack_mel(const float* mel, void* out_bf16, int B, int n_mels, int T, int c_pad,
                    const unsigned int* clip_max_bits /* null = mel is final */, cudaStream_t stream);
// Splice gather (S1/S2): see include/audiollm_b200.h al_splice.
int launch_splice(const void* table, int elem_bytes, int d, const long long* input_ids, const long long* attn_mask,
                  const long long* labels, int B, int t_txt, int n_audio, long long start_id, long long end_id,
                  const void* audio_rows, void* out, float* mask_out, long long* labels_out, long long vocab,
                  int* bad_id_flag, cudaStream_t stream);
int launch_splice_ragged(const void* table, int elem_bytes, int d, const long long* input_ids,
                         const long long* attn_mask, const long long* labels, int B, int t_txt, int S_out,
                         const int* span_rows, const int* span_src_row, const int* n_spans, int max_spans,
                         const void* audio_rows, long long start_id, long long end_id, void* out, float* mask_out,
                         long long* labels_out, int* span_start_out, long long vocab, int* bad_id_flag,
                         cudaStream_t stream);
int launch_f32_to_bf16(const float* x, void* out, long long n, cudaStream_t stream);
int launch_transpose_bf16(const void* in, void* out, int R, int C, int out_ld, cudaStream_t stream);
int launch_lora_pack(const float* A, const float* Bm, int rank, int r_pad, int in_dim, int out_dim, float scaling, void* a_pad,
                     void* b_pad, cudaStream_t stream);
int launch_layernorm_bwd(const float* y, const float* dout, const float* gamma, void* dy_bf16, float* dgamma,
                         float* dbeta, float* dbias, int rows, int d, float eps, cudaStream_t stream);
int launch_colsum_bf16(const void* x, float* out, int rows, int n, cudaStream_t stream);

// ------------------------------------------------------------------ mel (mel.cu)
struct MelTables {
  const float* window;     // [400]
  const float2* tw200;     // [200] exp(-2 pi i k / 200)
  const float2* tw400;     // [101] exp(-2 pi i k / 400)
  const int* col_start;    // [n_mels + 1] CSC offsets into nz_*
  const int* nz_freq;      // [nnz]
  const float* nz_w;       // [nnz]
  int n_mels;
  int nnz;
  // tensor-core form (mel_tc.cu): tc_bank = index of the compiled bank structure (mel_bank_struct.h) this bank's
  // non-zero pattern equals, or -1 (only the FFT kernel applies)
  const uint8_t* tc_b_image;   // [2 CTA ranks][MEL_TC_B_BYTES] fp16 twiddle operands in their shared-memory layout
  int tc_bank;
};
// twiddle operand image of one CTA of the pair: 8 matrices (Ce, Se, Co, So x hi, lo), each 56 rows x 112 K halves as
// [14 K chunks of 8 halves][7 row groups][8 rows][16 B]. K position 16 b + e holds sample index i = b + 7 e.
constexpr int MEL_TC_KCHUNK_BYTES = 7 * 128;
constexpr int MEL_TC_MAT_BYTES = 14 * MEL_TC_KCHUNK_BYTES;
constexpr int MEL_TC_B_BYTES = 8 * MEL_TC_MAT_BYTES;
int mel_tc_set_window(const float* host_win_2x112);
int mel_tc_set_weights(int bank, const float* host_w_201x2);
int launch_mel_tc(const float* wave, const int* n_samples, int B, long long wave_stride, const MelTables& tb, int mode,
                  float* out, unsigned int* clip_max_bits, int num_sms, cudaStream_t stream);
// mode 0: whisper (log10, per-clip max written to clip_max for finalize); mode 1: ln(x + 1e-9).
int launch_mel(const float* wave, const int* n_samples, int B, long long wave_stride, const MelTables& tb, int mode,
               float* out, unsigned int* clip_max_bits, cudaStream_t stream);
int launch_mel_finalize(float* out, const unsigned int* clip_max_bits, int B, int n_mels, cudaStream_t stream);

// ------------------------------------------------------------------ LLaMA-side row kernels (llama_rows.cu)
int launch_rmsnorm(const void* x, const void* w, void* y, float* rstd, const void* dy, const void* addend, void* dx, int rows, int d,
                   float eps, bool backward, cudaStream_t st);
int launch_swiglu(const void* gate, const void* up, const void* dh, void* out0, void* out1, long long n, bool backward,
                  int num_sms, cudaStream_t st);
int launch_rope(const void* x, const void* cs, const void* sn, void* out, int B, int S, int H, int hd, int cos_batch,
                int backward, int num_sms, cudaStream_t st);
int launch_ce_inplace(void* logits, const long long* labels, int rows, int vocab, long long ld, float grad_scale,
                      float* loss_sum, cudaStream_t st);

// ------------------------------------------------------------------ ingest (ingest.cu)
struct ResampleTable {
  const float* weights;   // [new_f][max_taps] non-zero run of each phase's filter (zero padded)
  const int* first;       // [new_f] index of the run's first tap inside the dense 2*width + orig_f filter
  int orig_f, new_f;      // gcd-reduced rates
  int width;              // torchaudio's `width` (left zero padding of the dense convolution)
  int max_taps;
};
// in: [B][n_chan][..] fp32 (strides in elements); tb == nullptr means same rate (channel mean + pad only).
int launch_ingest(const float* in, long long clip_stride, long long chan_stride, int n_chan, const int* n_in,
                  int n_in_cap, const ResampleTable* tb, float* out, long long out_stride, int out_cap, int* n_out,
                  int B, cudaStream_t stream);

}  // namespace al
