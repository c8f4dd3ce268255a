// HBM-bound row kernels: LayerNorm, mel repack for the conv-stem GEMM, the splice gather, casts.
#include "common.cuh"
#include "kernels.h"

namespace al {

// ----------------------------------------------------------------------------- LayerNorm
// One warp per row, row held in registers (d <= 128*MAXV), two-pass mean / variance in fp32.
// Replaces nn.LayerNorm of HF WhisperEncoderLayer (modeling_whisper.py:393, 403), the encoder's final
// layer_norm (:643) and AudioProjector.layers[3] (/root/reference/src/models/projector.py:15).
template <int MAXV, bool OUT_F32>
__global__ void __launch_bounds__(256, (MAXV <= 12) ? 4 : 1)   // d <= 1536: <= 64 registers, 32 warps / SM in flight
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 void* __restrict__ out, int rows, int d, float eps, long long out_ld, int rows_per_group,
                 long long out_group_stride, long long out_row_offset) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = d >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * d);
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      v[i] = xr[idx];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + b * b) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / d + eps);
  const long long orow = static_cast<long long>(row / rows_per_group) * out_group_stride + out_row_offset +
                         (row % rows_per_group);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nv) {
      const float4 g = __ldg(g4 + idx), bb = __ldg(b4 + idx);
      float4 y;
      y.x = (v[i].x - mean) * rstd * g.x + bb.x;
      y.y = (v[i].y - mean) * rstd * g.y + bb.y;
      y.z = (v[i].z - mean) * rstd * g.z + bb.z;
      y.w = (v[i].w - mean) * rstd * g.w + bb.w;
      if constexpr (OUT_F32) {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + orow * out_ld)[idx] = y;
      } else {
        uint2 pk;
        pk.x = pack_bf16(y.x, y.y);
        pk.y = pack_bf16(y.z, y.w);
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + orow * out_ld)[idx] = pk;
      }
    }
  }
}

template <int MAXV>
static int ln_dispatch(const float* x, const float* g, const float* b, void* out, int rows, int d, float eps,
                       int out_dtype, long long out_ld, int rpg, long long ogs, long long oro, cudaStream_t st) {
  dim3 grid((rows + 7) / 8);
  if (out_dtype == 1)
    layernorm_kernel<MAXV, true><<<grid, 256, 0, st>>>(x, g, b, out, rows, d, eps, out_ld, rpg, ogs, oro);
  else
    layernorm_kernel<MAXV, false><<<grid, 256, 0, st>>>(x, g, b, out, rows, d, eps, out_ld, rpg, ogs, oro);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, void* out, int rows, int d, float eps,
                     int out_dtype, long long out_ld, int rows_per_group, long long out_group_stride,
                     long long out_row_offset, cudaStream_t stream) {
  AL_REQUIRE(d % 4 == 0 && d <= 128 * 32, "layernorm: d=%d must be a multiple of 4 and <= 4096", d);
  AL_REQUIRE(rows_per_group > 0, "layernorm: rows_per_group must be positive");
  if (rows == 0) return 0;
  if (d <= 128 * 4) return ln_dispatch<4>(x, gamma, beta, out, rows, d, eps, out_dtype, out_ld, rows_per_group, out_group_stride, out_row_offset, stream);
  if (d <= 128 * 12) return ln_dispatch<12>(x, gamma, beta, out, rows, d, eps, out_dtype, out_ld, rows_per_group, out_group_stride, out_row_offset, stream);
  if (d <= 128 * 24) return ln_dispatch<24>(x, gamma, beta, out, rows, d, eps, out_dtype, out_ld, rows_per_group, out_group_stride, out_row_offset, stream);
  return ln_dispatch<32>(x, gamma, beta, out, rows, d, eps, out_dtype, out_ld, rows_per_group, out_group_stride, out_row_offset, stream);
}

// ----------------------------------------------------------------------------- mel repack
// [B][n_mels][T] fp32 (what the reference hands the encoder, allm.py:214) -> [B][T+2][c_pad] bf16, time-major
// with one zero row before and after each clip: row t+1 holds frame t. In this layout the im2col row of
// conv1 (k=3, pad=1) for output frame t is the 3*c_pad contiguous elements starting at row t.
__global__ void __launch_bounds__(256)
pack_mel_kernel(const float* __restrict__ mel, __nv_bfloat16* __restrict__ out, int n_mels, int T, int c_pad) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int m = m0 + i, t = t0 + tx;
    tile[i][tx] = (m < n_mels && t < T) ? mel[(static_cast<long long>(b) * n_mels + m) * T + t] : 0.f;
  }
  __syncthreads();
  __nv_bfloat16* ob = out + static_cast<long long>(b) * (T + 2) * c_pad;
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, m = m0 + tx;
    if (t < T && m < c_pad) ob[static_cast<long long>(t + 1) * c_pad + m] = __float2bfloat16_rn(tile[tx][i]);
  }
  // zero rows 0 and T+1 (first / last time-tile of each channel-tile)
  if (blockIdx.x == 0 && ty == 0 && m0 + tx < c_pad) ob[m0 + tx] = __float2bfloat16_rn(0.f);
  if (blockIdx.x == gridDim.x - 1 && ty == 0 && m0 + tx < c_pad)
    ob[static_cast<long long>(T + 1) * c_pad + m0 + tx] = __float2bfloat16_rn(0.f);
}

int launch_pack_mel(const float* mel, void* out, int B, int n_mels, int T, int c_pad, cudaStream_t stream) {
  dim3 grid((T + 31) / 32, (c_pad + 31) / 32, B);
  pack_mel_kernel<<<grid, 256, 0, stream>>>(mel, reinterpret_cast<__nv_bfloat16*>(out), n_mels, T, c_pad);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------------------- splice (S1 / S2)
// One warp per output row of inputs_embeds. Row map of sample b (allm.py:165-170; bit-exact contract):
//   0 <- E[<audio>] ; 1..A <- audio rows (copied only if `audio_rows` != null — the projector's LayerNorm
//   normally writes them in place) ; A+1 <- E[</audio>] ; A+2+j <- E[input_ids[b][j]].
// mask_out = [1.0f x (A+2), attention_mask] (float32, allm.py:192-194); labels_out = [-100 x (A+2), labels].
__device__ __forceinline__ void copy_row_16B(void* dst, const void* src, int n16, int lane) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (int i = lane; i < n16; i += 32) d[i] = __ldg(s + i);
}

__global__ void __launch_bounds__(256)
splice_kernel(const uint8_t* __restrict__ table, long long row_bytes, const long long* __restrict__ input_ids,
              const long long* __restrict__ attn_mask, const long long* __restrict__ labels, int B, int t_txt,
              int n_audio, long long start_id, long long end_id, const uint8_t* __restrict__ audio_rows,
              uint8_t* __restrict__ out, float* __restrict__ mask_out, long long* __restrict__ labels_out) {
  const int S = n_audio + 2 + t_txt;
  const long long gw = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (gw >= static_cast<long long>(B) * S) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(gw / S), r = static_cast<int>(gw % S);
  uint8_t* dst = out + gw * row_bytes;
  const int n16 = static_cast<int>(row_bytes >> 4);
  float mk = 1.0f;
  long long lb = -100;
  if (r == 0) {
    copy_row_16B(dst, table + start_id * row_bytes, n16, lane);
  } else if (r <= n_audio) {
    if (audio_rows) copy_row_16B(dst, audio_rows + (static_cast<long long>(b) * n_audio + (r - 1)) * row_bytes, n16, lane);
  } else if (r == n_audio + 1) {
    copy_row_16B(dst, table + end_id * row_bytes, n16, lane);
  } else {
    const int j = r - n_audio - 2;
    const long long id = input_ids[static_cast<long long>(b) * t_txt + j];
    copy_row_16B(dst, table + id * row_bytes, n16, lane);
    if (attn_mask) mk = static_cast<float>(attn_mask[static_cast<long long>(b) * t_txt + j]);
    if (labels) lb = labels[static_cast<long long>(b) * t_txt + j];
  }
  if (lane == 0) {
    if (mask_out) mask_out[gw] = mk;
    if (labels_out) labels_out[gw] = lb;
  }
}

int launch_splice(const void* table, int elem_bytes, int d, const long long* input_ids, const long long* attn_mask,
                  const long long* labels, int B, int t_txt, int n_audio, long long start_id, long long end_id,
                  const void* audio_rows, void* out, float* mask_out, long long* labels_out, cudaStream_t stream);

// Ragged extension (SURVEY.md §8 extension row; not in the reference). Sample b has n_spans[b] <= max_spans
// spans; span k keeps span_rows[b][k] encoder rows taken from audio_rows at row span_src_row[b][k]. Output rows:
// [<audio>, a_1 rows, </audio>, <audio>, a_2 rows, </audio>, ..., text..., zero pad to S_out]. The span start
// offsets are the exclusive prefix sum over (a_k + 2), computed here on the device (and optionally written to
// span_start_out for the bit-exact index check).
__global__ void __launch_bounds__(256)
splice_ragged_kernel(const uint8_t* __restrict__ table, long long row_bytes, const long long* __restrict__ input_ids,
                     const long long* __restrict__ attn_mask, const long long* __restrict__ labels, int B, int t_txt,
                     int S_out, const int* __restrict__ span_rows, const int* __restrict__ span_src_row,
                     const int* __restrict__ n_spans, int max_spans, const uint8_t* __restrict__ audio_rows,
                     long long start_id, long long end_id, uint8_t* __restrict__ out, float* __restrict__ mask_out,
                     long long* __restrict__ labels_out, int* __restrict__ span_start_out) {
  const long long gw = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (gw >= static_cast<long long>(B) * S_out) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(gw / S_out), r = static_cast<int>(gw % S_out);
  uint8_t* dst = out + gw * row_bytes;
  const int n16 = static_cast<int>(row_bytes >> 4);
  const int ns = n_spans[b];
  // exclusive prefix sum over (a_k + 2); every warp of the sample recomputes it (max_spans is tiny)
  int start = 0, kind = 3 /*0 <audio>, 1 audio row, 2 </audio>, 3 after spans*/, src = 0;
  for (int k = 0; k < ns; ++k) {
    const int a = span_rows[b * max_spans + k];
    if (r == 0 && lane == 0 && span_start_out) span_start_out[b * max_spans + k] = start;
    if (r >= start && r < start + a + 2) {
      const int o = r - start;
      kind = (o == 0) ? 0 : (o == a + 1 ? 2 : 1);
      src = span_src_row[b * max_spans + k] + o - 1;
    }
    start += a + 2;
  }
  const int text_off = start;
  float mk = 1.0f;
  long long lb = -100;
  if (kind == 0) {
    copy_row_16B(dst, table + start_id * row_bytes, n16, lane);
  } else if (kind == 1) {
    copy_row_16B(dst, audio_rows + static_cast<long long>(src) * row_bytes, n16, lane);
  } else if (kind == 2) {
    copy_row_16B(dst, table + end_id * row_bytes, n16, lane);
  } else if (r < text_off + t_txt) {
    const int j = r - text_off;
    const long long id = input_ids[static_cast<long long>(b) * t_txt + j];
    copy_row_16B(dst, table + id * row_bytes, n16, lane);
    mk = attn_mask ? static_cast<float>(attn_mask[static_cast<long long>(b) * t_txt + j]) : 1.0f;
    if (labels) lb = labels[static_cast<long long>(b) * t_txt + j];
  } else {
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < n16; i += 32) reinterpret_cast<uint4*>(dst)[i] = z;
    mk = 0.f;
  }
  if (lane == 0) {
    if (mask_out) mask_out[gw] = mk;
    if (labels_out) labels_out[gw] = lb;
  }
}

int launch_splice(const void* table, int elem_bytes, int d, const long long* input_ids, const long long* attn_mask,
                  const long long* labels, int B, int t_txt, int n_audio, long long start_id, long long end_id,
                  const void* audio_rows, void* out, float* mask_out, long long* labels_out, cudaStream_t stream) {
  const long long row_bytes = static_cast<long long>(d) * elem_bytes;
  AL_REQUIRE(row_bytes % 16 == 0, "splice: row of %lld bytes is not a multiple of 16", row_bytes);
  const long long rows = static_cast<long long>(B) * (n_audio + 2 + t_txt);
  if (rows == 0) return 0;
  splice_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const uint8_t*>(table), row_bytes, input_ids, attn_mask, labels, B, t_txt, n_audio, start_id,
      end_id, reinterpret_cast<const uint8_t*>(audio_rows), reinterpret_cast<uint8_t*>(out), mask_out, labels_out);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_splice_ragged(const void* table, int elem_bytes, int d, const long long* input_ids,
                         const long long* attn_mask, const long long* labels, int B, int t_txt, int S_out,
                         const int* span_rows, const int* span_src_row, const int* n_spans, int max_spans,
                         const void* audio_rows, long long start_id, long long end_id, void* out, float* mask_out,
                         long long* labels_out, int* span_start_out, cudaStream_t stream) {
  const long long row_bytes = static_cast<long long>(d) * elem_bytes;
  AL_REQUIRE(row_bytes % 16 == 0, "splice: row of %lld bytes is not a multiple of 16", row_bytes);
  const long long rows = static_cast<long long>(B) * S_out;
  if (rows == 0) return 0;
  splice_ragged_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const uint8_t*>(table), row_bytes, input_ids, attn_mask, labels, B, t_txt, S_out, span_rows,
      span_src_row, n_spans, max_spans, reinterpret_cast<const uint8_t*>(audio_rows), start_id, end_id,
      reinterpret_cast<uint8_t*>(out), mask_out, labels_out, span_start_out);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------------------- casts
__global__ void f32_to_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = x[i];
    out[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
}
int launch_f32_to_bf16(const float* x, void* out, long long n, cudaStream_t stream) {
  AL_REQUIRE(n % 4 == 0, "f32_to_bf16: n=%lld must be a multiple of 4", n);
  if (n == 0) return 0;
  const long long n4 = n / 4;
  const unsigned grid = static_cast<unsigned>((n4 + 255) / 256 > 148 * 8 ? 148 * 8 : (n4 + 255) / 256);
  f32_to_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(out), n4);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace al
