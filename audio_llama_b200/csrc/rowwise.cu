// HBM-bound row kernels: LayerNorm, mel repack for the conv-stem GEMM, the splice gather, casts.
#include <stdlib.h>
#include <string.h>

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

// d = 128 * VPL exactly (every shape of the path: 1280, 2048, 3072, 384, 256): persistent warps that walk the rows,
// gamma / beta staged once per CTA in shared memory, all arithmetic as packed fp32x2, y = x * (rstd * gamma) + (beta -
// mean * rstd * gamma). ~200 issue slots per row instead of ~620 (the generic kernel re-reads gamma / beta through L1
// for every row and spends three scalar operations per element): inside the power-capped step (SM clock ~1.4 GHz) the
// generic kernel was issue-limited at 0.78 of the HBM copy bandwidth while it reaches 0.94 at full clock.
template <int VPL, bool OUT_F32>
__global__ void __launch_bounds__(256, (VPL <= 12) ? 4 : 2)
layernorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      void* __restrict__ out, int rows, float eps, long long out_ld, int rows_per_group,
                      long long out_group_stride, long long out_row_offset) {
  constexpr int D = 128 * VPL;
  __shared__ float4 s_g[32 * VPL], s_b[32 * VPL];
  for (int i = threadIdx.x; i < 32 * VPL; i += 256) {
    s_g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i);
    s_b[i] = __ldg(reinterpret_cast<const float4*>(beta) + i);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * 8;
  constexpr float INV_D = 1.0f / D;
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += warps) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * D);
    unsigned long long v[2 * VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 t = xr[lane + 32 * i];
      v[2 * i] = pk2(t.x, t.y);
      v[2 * i + 1] = pk2(t.z, t.w);
    }
    unsigned long long s2 = v[0];
#pragma unroll
    for (int i = 1; i < 2 * VPL; ++i) s2 = fadd2(s2, v[i]);
    float sa, sb;
    unpk2(s2, sa, sb);
    const float mean = warp_sum(sa + sb) * INV_D;
    const unsigned long long nmean2 = pk2(-mean, -mean);
    unsigned long long q2 = pk2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 2 * VPL; ++i) {
      const unsigned long long dlt = fadd2(v[i], nmean2);
      q2 = ffma2(dlt, dlt, q2);
    }
    float qa, qb;
    unpk2(q2, qa, qb);
    const float rstd = rsqrtf(warp_sum(qa + qb) * INV_D + eps);
    const unsigned long long rstd2 = pk2(rstd, rstd);
    const long long orow = static_cast<long long>(row / rows_per_group) * out_group_stride + out_row_offset +
                           (row % rows_per_group);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float4 g = s_g[lane + 32 * i], bb = s_b[lane + 32 * i];
      const unsigned long long a0 = fmul2(pk2(g.x, g.y), rstd2), a1 = fmul2(pk2(g.z, g.w), rstd2);
      const unsigned long long y0 = ffma2(v[2 * i], a0, ffma2(a0, nmean2, pk2(bb.x, bb.y)));
      const unsigned long long y1 = ffma2(v[2 * i + 1], a1, ffma2(a1, nmean2, pk2(bb.z, bb.w)));
      float4 y;
      unpk2(y0, y.x, y.y);
      unpk2(y1, y.z, y.w);
      if constexpr (OUT_F32) {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + orow * out_ld)[lane + 32 * i] = y;
      } else {
        uint2 pk;
        pk.x = pack_bf16(y.x, y.y);
        pk.y = pack_bf16(y.z, y.w);
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + orow * out_ld)[lane + 32 * i] = pk;
      }
    }
  }
}

template <int VPL>
static int ln_rows_dispatch(const float* x, const float* g, const float* b, void* out, int rows, float eps, int out_dtype,
                            long long out_ld, int rpg, long long ogs, long long oro, cudaStream_t st) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int per_sm = VPL <= 12 ? 4 : 2;
  const int grid = min((rows + 7) / 8, sms * per_sm);
  if (out_dtype == 1)
    layernorm_rows_kernel<VPL, true><<<grid, 256, 0, st>>>(x, g, b, out, rows, eps, out_ld, rpg, ogs, oro);
  else
    layernorm_rows_kernel<VPL, false><<<grid, 256, 0, st>>>(x, g, b, out, rows, eps, out_ld, rpg, ogs, oro);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
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
#define AL_LN_ROWS(V) \
  if (d == 128 * V) return ln_rows_dispatch<V>(x, gamma, beta, out, rows, eps, out_dtype, out_ld, rows_per_group, out_group_stride, out_row_offset, stream)
  static const bool generic_only = [] { const char* e = getenv("AUDIOLLM_B200_LN"); return e && strcmp(e, "generic") == 0; }();
  if (!generic_only) {
    AL_LN_ROWS(1); AL_LN_ROWS(2); AL_LN_ROWS(3); AL_LN_ROWS(10); AL_LN_ROWS(16); AL_LN_ROWS(24);
  }
#undef AL_LN_ROWS
  if (d <= 128 * 4) return ln_dispatch<4>(x, gamma, beta, out, rows, d, eps, out_dtype, out_ld, rows_per_group, out_group_stride, out_row_offset, stream);
  if (d <= 128 * 12) return ln_dispatch<12>(x, gamma, beta, out, rows, d, eps, out_dtype, out_ld, rows_per_group, out_group_stride, out_row_offset, stream);
  if (d <= 128 * 24) return ln_dispatch<24>(x, gamma, beta, out, rows, d, eps, out_dtype, out_ld, rows_per_group, out_group_stride, out_row_offset, stream);
  return ln_dispatch<32>(x, gamma, beta, out, rows, d, eps, out_dtype, out_ld, rows_per_group, out_group_stride, out_row_offset, stream);
}

// ----------------------------------------------------------------------------- mel repack
// [B][n_mels][T] fp32 (what the reference hands the encoder, allm.py:214) -> [B][T+2][c_pad] bf16, time-major
// with one zero row before and after each clip: row t+1 holds frame t. In this layout the im2col row of
// conv1 (k=3, pad=1) for output frame t is the 3*c_pad contiguous elements starting at row t.
// clip_max_bits != null: `mel` holds the log10 values BEFORE the per-clip floor (al_mel_forward_ex, AL_MEL_RAW) and
// the floor + affine step of the feature extractor, (max(L, clipmax - 8) + 4) / 4 (HF :156-159), is applied here with
// the very operations of mel_finalize_kernel: the f32 mel is then written once and read once on the pipeline path.
__global__ void __launch_bounds__(256)
pack_mel_kernel(const float* __restrict__ mel, __nv_bfloat16* __restrict__ out, int n_mels, int T, int c_pad,
                const unsigned int* __restrict__ clip_max_bits) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  float floor_v = 0.f;
  if (clip_max_bits) {
    const unsigned int u = clip_max_bits[b];                                   // ordered-int encoding of the clip max
    floor_v = __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u) - 8.0f;
  }
  const int t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int m = m0 + i, t = t0 + tx;
    float v = 0.f;
    if (m < n_mels && t < T) {
      v = mel[(static_cast<long long>(b) * n_mels + m) * T + t];
      if (clip_max_bits) v = (fmaxf(v, floor_v) + 4.0f) / 4.0f;
    }
    tile[i][tx] = v;
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

int launch_pack_mel(const float* mel, void* out, int B, int n_mels, int T, int c_pad, const unsigned int* clip_max_bits,
                    cudaStream_t stream) {
  dim3 grid((T + 31) / 32, (c_pad + 31) / 32, B);
  pack_mel_kernel<<<grid, 256, 0, stream>>>(mel, reinterpret_cast<__nv_bfloat16*>(out), n_mels, T, c_pad, clip_max_bits);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------------------- splice (S1 / S2)
// One warp per output row of inputs_embeds. Row map of sample b (allm.py:165-170; bit-exact contract):
//   0 <- E[<audio>] ; 1..A <- audio rows (copied only if `audio_rows` != null — the projector's LayerNorm
//   normally writes them in place) ; A+1 <- E[</audio>] ; A+2+j <- E[input_ids[b][j]].
// mask_out = [1.0f x (A+2), attention_mask] (float32, allm.py:192-194); labels_out = [-100 x (A+2), labels].
// Four 16-byte loads in flight per lane before the first store (a 4 KB row is 8 of them per lane).
#ifndef AL_SPLICE_VARIANT
#define AL_SPLICE_VARIANT 2
#endif
// Streaming 16-byte accesses: every byte of the splice is read once and written once, so loads skip L1 allocation and
// stores carry the cache-streaming hint (evict-first in L2: the written rows start their way to HBM at once instead of
// lingering as dirty lines that a later kernel has to push out).
__device__ __forceinline__ uint4 ld_stream_16B(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_16B(uint4* p, const uint4& v) {
#if AL_SPLICE_VARIANT >= 2
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
  *p = v;
#endif
}
__device__ __forceinline__ void copy_row_16B(void* dst, const void* src, int n16, int lane) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  int i = lane;
#if AL_SPLICE_VARIANT >= 1
  // eight 16-byte loads in flight per lane before the first store (a 4 KB row = exactly one such batch per lane)
  for (; i + 224 < n16; i += 256) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ld_stream_16B(s + i + 32 * k);
#pragma unroll
    for (int k = 0; k < 8; ++k) st_stream_16B(d + i + 32 * k, v[k]);
  }
#endif
  for (; i + 96 < n16; i += 128) {
    const uint4 a = ld_stream_16B(s + i), b = ld_stream_16B(s + i + 32), c = ld_stream_16B(s + i + 64), e = ld_stream_16B(s + i + 96);
    st_stream_16B(d + i, a); st_stream_16B(d + i + 32, b); st_stream_16B(d + i + 64, c); st_stream_16B(d + i + 96, e);
  }
  for (; i < n16; i += 32) st_stream_16B(d + i, ld_stream_16B(s + i));
}

// COPY_AUDIO = true : one warp per output row, audio rows copied from `audio_rows`.
// COPY_AUDIO = false: the audio rows are already in place (the projector's LayerNorm wrote them), so only the
//   t_txt + 2 gathered rows of a sample get a warp; the mask / labels entries of the A audio rows are written by the
//   text-row warps, ceil(A / t_txt) each (or all by the <audio> warp when there is no text).
__device__ __forceinline__ void zero_row_16B(void* dst, int n16, int lane) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int i = lane; i < n16; i += 32) reinterpret_cast<uint4*>(dst)[i] = z;
}

template <bool COPY_AUDIO>
__global__ void __launch_bounds__(256)
splice_kernel(const uint8_t* __restrict__ table, long long row_bytes, const long long* __restrict__ input_ids,
              const long long* __restrict__ attn_mask, const long long* __restrict__ labels, int B, int t_txt,
              int n_audio, long long start_id, long long end_id, const uint8_t* __restrict__ audio_rows,
              uint8_t* __restrict__ out, float* __restrict__ mask_out, long long* __restrict__ labels_out,
              long long vocab, int* __restrict__ bad_id_flag) {
  const int S = n_audio + 2 + t_txt;
  const int per_sample = COPY_AUDIO ? S : t_txt + 2;
  const long long gw = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (gw >= static_cast<long long>(B) * per_sample) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(gw / per_sample);
  int r = static_cast<int>(gw % per_sample);
  if (!COPY_AUDIO && r >= 1) r += n_audio;               // 0 -> 0, 1 -> A + 1, 2 + j -> A + 2 + j
  const long long orow = static_cast<long long>(b) * S + r;
  uint8_t* dst = out + orow * row_bytes;
  const int n16 = static_cast<int>(row_bytes >> 4);
  float mk = 1.0f;
  long long lb = -100;
  if (r == 0) {
    copy_row_16B(dst, table + start_id * row_bytes, n16, lane);
  } else if (r <= n_audio) {
    if (COPY_AUDIO && audio_rows)
      copy_row_16B(dst, audio_rows + (static_cast<long long>(b) * n_audio + (r - 1)) * row_bytes, n16, lane);
  } else if (r == n_audio + 1) {
    copy_row_16B(dst, table + end_id * row_bytes, n16, lane);
  } else {
    const int j = r - n_audio - 2;
    const long long id = input_ids[static_cast<long long>(b) * t_txt + j];
    if (id >= 0 && id < vocab) {
      copy_row_16B(dst, table + id * row_bytes, n16, lane);
    } else {
      // the reference's embed_tokens(input_ids) raises on such an id (allm.py:64); here nothing outside the table is
      // read: the row is zeroed and the flag makes the host side raise (ops.splice / AudioConditioner)
      zero_row_16B(dst, n16, lane);
      if (lane == 0 && bad_id_flag) atomicOr(bad_id_flag, 1);
    }
    if (attn_mask) mk = static_cast<float>(attn_mask[static_cast<long long>(b) * t_txt + j]);
    if (labels) lb = labels[static_cast<long long>(b) * t_txt + j];
  }
  if (lane == 0) {
    if (mask_out) mask_out[orow] = mk;
    if (labels_out) labels_out[orow] = lb;
  }
  if (!COPY_AUDIO) {
    // mask 1.0 / labels -100 of the audio rows 1..A (allm.py:81-89, 184-196)
    int a0, a1;
    if (t_txt > 0) {
      const int per = (n_audio + t_txt - 1) / t_txt;
      const int j = r - n_audio - 2;
      a0 = j >= 0 ? 1 + j * per : 0;
      a1 = j >= 0 ? min(a0 + per, n_audio + 1) : 0;
    } else {
      a0 = r == 0 ? 1 : 0;
      a1 = r == 0 ? n_audio + 1 : 0;
    }
    for (int a = a0 + lane; a < a1; a += 32) {
      if (mask_out) mask_out[static_cast<long long>(b) * S + a] = 1.0f;
      if (labels_out) labels_out[static_cast<long long>(b) * S + a] = -100;
    }
  }
}


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
                     long long* __restrict__ labels_out, int* __restrict__ span_start_out, long long vocab,
                     int* __restrict__ bad_id_flag) {
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
    if (id >= 0 && id < vocab) {
      copy_row_16B(dst, table + id * row_bytes, n16, lane);
    } else {
      zero_row_16B(dst, n16, lane);
      if (lane == 0 && bad_id_flag) atomicOr(bad_id_flag, 1);
    }
    mk = attn_mask ? static_cast<float>(attn_mask[static_cast<long long>(b) * t_txt + j]) : 1.0f;
    if (labels) lb = labels[static_cast<long long>(b) * t_txt + j];
  } else {
    zero_row_16B(dst, n16, lane);
    mk = 0.f;
  }
  if (lane == 0) {
    if (mask_out) mask_out[gw] = mk;
    if (labels_out) labels_out[gw] = lb;
  }
}

int launch_splice(const void* table, int elem_bytes, int d, const long long* input_ids, const long long* attn_mask,
                  const long long* labels, int B, int t_txt, int n_audio, long long start_id, long long end_id,
                  const void* audio_rows, void* out, float* mask_out, long long* labels_out, long long vocab,
                  int* bad_id_flag, cudaStream_t stream) {
  const long long row_bytes = static_cast<long long>(d) * elem_bytes;
  AL_REQUIRE(row_bytes % 16 == 0, "splice: row of %lld bytes is not a multiple of 16", row_bytes);
  const bool copy_audio = audio_rows != nullptr;
  const long long rows = static_cast<long long>(B) * (copy_audio ? n_audio + 2 + t_txt : t_txt + 2);
  if (rows == 0) return 0;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (copy_audio)
    splice_kernel<true><<<grid, 256, 0, stream>>>(
        reinterpret_cast<const uint8_t*>(table), row_bytes, input_ids, attn_mask, labels, B, t_txt, n_audio, start_id,
        end_id, reinterpret_cast<const uint8_t*>(audio_rows), reinterpret_cast<uint8_t*>(out), mask_out, labels_out, vocab,
        bad_id_flag);
  else
    splice_kernel<false><<<grid, 256, 0, stream>>>(
        reinterpret_cast<const uint8_t*>(table), row_bytes, input_ids, attn_mask, labels, B, t_txt, n_audio, start_id,
        end_id, nullptr, reinterpret_cast<uint8_t*>(out), mask_out, labels_out, vocab, bad_id_flag);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_splice_ragged(const void* table, int elem_bytes, int d, const long long* input_ids,
                         const long long* attn_mask, const long long* labels, int B, int t_txt, int S_out,
                         const int* span_rows, const int* span_src_row, const int* n_spans, int max_spans,
                         const void* audio_rows, long long start_id, long long end_id, void* out, float* mask_out,
                         long long* labels_out, int* span_start_out, long long vocab, int* bad_id_flag,
                         cudaStream_t stream) {
  const long long row_bytes = static_cast<long long>(d) * elem_bytes;
  AL_REQUIRE(row_bytes % 16 == 0, "splice: row of %lld bytes is not a multiple of 16", row_bytes);
  const long long rows = static_cast<long long>(B) * S_out;
  if (rows == 0) return 0;
  splice_ragged_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const uint8_t*>(table), row_bytes, input_ids, attn_mask, labels, B, t_txt, S_out, span_rows,
      span_src_row, n_spans, max_spans, reinterpret_cast<const uint8_t*>(audio_rows), start_id, end_id,
      reinterpret_cast<uint8_t*>(out), mask_out, labels_out, span_start_out, vocab, bad_id_flag);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------------------- training-side row kernels
// bf16 [R][C] -> [C][R] (weight-gradient GEMMs contract over the row dimension; the tcgen05 GEMM wants both operands
// K-major, so dy / h / x are transposed once per step: ~0.1 ms per 100 MB, noise next to the LLaMA backward).
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C, int out_ld) {
  // 64 x 64 tile through shared memory (row pitch 66 elements = 33 words: the column reads below are conflict-free).
  // Global accesses are 8 bytes per thread when the tile is interior and the row pitches allow it: 16 lanes cover one
  // 128-byte tile row, so a warp instruction moves 2 full rows instead of half of one.
  __shared__ __nv_bfloat16 tile[64][66];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const bool vec_in = (C % 4 == 0) && r0 + 64 <= R && c0 + 64 <= C;
  const bool vec_out = (out_ld % 4 == 0) && c0 + 64 <= C && r0 + 64 <= out_ld;
  if (vec_in) {
    const int q = threadIdx.x & 15, rr = threadIdx.x >> 4;          // 16 x 8-byte columns, 16 rows per pass
#pragma unroll
    for (int i = 0; i < 64; i += 16) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(in + static_cast<long long>(r0 + i + rr) * C + c0 + 4 * q));
      uint32_t* t = reinterpret_cast<uint32_t*>(&tile[i + rr][4 * q]);
      t[0] = v.x;
      t[1] = v.y;
    }
  } else {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int i = ty; i < 64; i += 4) {
      const int r = r0 + i, c = c0 + tx;
      tile[i][tx] = (r < R && c < C) ? in[static_cast<long long>(r) * C + c] : __float2bfloat16_rn(0.f);
    }
  }
  __syncthreads();
  if (vec_out) {
    const int q = threadIdx.x & 15, cc = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 64; i += 16) {
      const int c = i + cc;
      const unsigned short a0 = *reinterpret_cast<const unsigned short*>(&tile[4 * q][c]);
      const unsigned short a1 = *reinterpret_cast<const unsigned short*>(&tile[4 * q + 1][c]);
      const unsigned short a2 = *reinterpret_cast<const unsigned short*>(&tile[4 * q + 2][c]);
      const unsigned short a3 = *reinterpret_cast<const unsigned short*>(&tile[4 * q + 3][c]);
      uint2 v;
      v.x = static_cast<uint32_t>(a0) | (static_cast<uint32_t>(a1) << 16);
      v.y = static_cast<uint32_t>(a2) | (static_cast<uint32_t>(a3) << 16);
      *reinterpret_cast<uint2*>(out + static_cast<long long>(c0 + c) * out_ld + r0 + 4 * q) = v;
    }
  } else {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int i = ty; i < 64; i += 4) {
      const int c = c0 + i, r = r0 + tx;
      if (c < C && r < out_ld) out[static_cast<long long>(c) * out_ld + r] = tile[tx][i];   // r in [R, out_ld): zero padding
    }
  }
}

int launch_transpose_bf16(const void* in, void* out, int R, int C, int out_ld, cudaStream_t stream) {
  if (R == 0 || C == 0) return 0;
  AL_REQUIRE(out_ld >= R, "transpose: out_ld=%d < rows=%d", out_ld, R);
  dim3 grid((C + 63) / 64, (out_ld + 63) / 64);
  transpose_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in),
                                                  reinterpret_cast<__nv_bfloat16*>(out), R, C, out_ld);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// LayerNorm backward (AudioProjector.layers[3]): y fp32 [rows][d] is the saved LN input, dout the upstream
// gradient (fp32). Per row: yh = (y - mu) rstd ; g = dout * gamma ; dy = rstd (g - mean(g) - yh mean(g yh)).
// Writes dy as bf16 (the A operand of the two backward GEMMs) and accumulates dgamma += dout*yh, dbeta += dout,
// dbias2 += dy (column sums) with one atomicAdd per column per CTA.
template <int MAXV>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dout, const float* __restrict__ gamma,
                     __nv_bfloat16* __restrict__ dy, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dbias, int rows, int d, float eps, int rows_per_cta) {
  extern __shared__ float sacc[];                 // [3][d] per-CTA column partial sums
  for (int i = threadIdx.x; i < 3 * d; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = d >> 2;
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(row_begin + rows_per_cta, rows);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  for (int row = row_begin + warp; row < row_end; row += 8) {
    const float4* yr = reinterpret_cast<const float4*>(y + static_cast<long long>(row) * d);
    const float4* dr = reinterpret_cast<const float4*>(dout + static_cast<long long>(row) * d);
    float4 v[MAXV], g[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        v[i] = yr[idx];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
    }
    const float mean = warp_sum(s) / d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / d + eps);
    float sg = 0.f, sgy = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        const float4 dd = dr[idx], gm = __ldg(g4 + idx);
        v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;            // yh
        // column partial sums: dgamma, dbeta
        atomicAdd(&sacc[4 * idx + 0], dd.x * v[i].x); atomicAdd(&sacc[4 * idx + 1], dd.y * v[i].y);
        atomicAdd(&sacc[4 * idx + 2], dd.z * v[i].z); atomicAdd(&sacc[4 * idx + 3], dd.w * v[i].w);
        atomicAdd(&sacc[d + 4 * idx + 0], dd.x); atomicAdd(&sacc[d + 4 * idx + 1], dd.y);
        atomicAdd(&sacc[d + 4 * idx + 2], dd.z); atomicAdd(&sacc[d + 4 * idx + 3], dd.w);
        g[i] = make_float4(dd.x * gm.x, dd.y * gm.y, dd.z * gm.z, dd.w * gm.w);
        sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        sgy += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
      }
    }
    const float mg = warp_sum(sg) / d, mgy = warp_sum(sgy) / d;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int idx = lane + 32 * i;
      if (idx < nv) {
        float4 o;
        o.x = rstd * (g[i].x - mg - v[i].x * mgy);
        o.y = rstd * (g[i].y - mg - v[i].y * mgy);
        o.z = rstd * (g[i].z - mg - v[i].z * mgy);
        o.w = rstd * (g[i].w - mg - v[i].w * mgy);
        atomicAdd(&sacc[2 * d + 4 * idx + 0], o.x); atomicAdd(&sacc[2 * d + 4 * idx + 1], o.y);
        atomicAdd(&sacc[2 * d + 4 * idx + 2], o.z); atomicAdd(&sacc[2 * d + 4 * idx + 3], o.w);
        reinterpret_cast<uint2*>(dy + static_cast<long long>(row) * d)[idx] =
            make_uint2(pack_bf16(o.x, o.y), pack_bf16(o.z, o.w));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    atomicAdd(dgamma + i, sacc[i]);
    atomicAdd(dbeta + i, sacc[d + i]);
    atomicAdd(dbias + i, sacc[2 * d + i]);
  }
}

int launch_layernorm_bwd(const float* y, const float* dout, const float* gamma, void* dy_bf16, float* dgamma,
                         float* dbeta, float* dbias, int rows, int d, float eps, cudaStream_t stream) {
  AL_REQUIRE(d % 4 == 0 && d <= 128 * 24, "layernorm_bwd: d=%d must be a multiple of 4 and <= 3072", d);
  if (rows == 0) return 0;
  const int rows_per_cta = 64;
  const int grid = (rows + rows_per_cta - 1) / rows_per_cta;
  const size_t smem = static_cast<size_t>(3) * d * sizeof(float);
  auto* dyp = reinterpret_cast<__nv_bfloat16*>(dy_bf16);
  if (d <= 128 * 4) layernorm_bwd_kernel<4><<<grid, 256, smem, stream>>>(y, dout, gamma, dyp, dgamma, dbeta, dbias, rows, d, eps, rows_per_cta);
  else if (d <= 128 * 12) layernorm_bwd_kernel<12><<<grid, 256, smem, stream>>>(y, dout, gamma, dyp, dgamma, dbeta, dbias, rows, d, eps, rows_per_cta);
  else layernorm_bwd_kernel<24><<<grid, 256, smem, stream>>>(y, dout, gamma, dyp, dgamma, dbeta, dbias, rows, d, eps, rows_per_cta);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Column sums of a bf16 [rows][n] matrix into fp32 [n] (bias gradients): out += sum over rows.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int rows, int n, int rows_per_cta) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= n) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += __bfloat162float(x[static_cast<long long>(r) * n + col]);
  atomicAdd(out + col, acc);
}

int launch_colsum_bf16(const void* x, float* out, int rows, int n, cudaStream_t stream) {
  if (rows == 0 || n == 0) return 0;
  const int rows_per_cta = 256;
  dim3 grid((n + 255) / 256, (rows + rows_per_cta - 1) / rows_per_cta);
  colsum_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), out, rows, n, rows_per_cta);
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

// The fused GEMM's LoRA operands from the fp32 parameters in ONE launch: a_pad [r_pad][in] = bf16(A) (rows >= rank zero),
// b_pad [out][r_pad] = bf16(scaling * B) (columns >= rank zero). (As separate torch ops this is seven tiny launches per
// LoRA-carrying linear per optimizer step: ~1200 launches in the README training step.)
__global__ void lora_pack_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int rank, int r_pad, int in_dim,
                                 int out_dim, float scaling, __nv_bfloat16* __restrict__ a_pad, __nv_bfloat16* __restrict__ b_pad) {
  const long long na = static_cast<long long>(r_pad) * in_dim, nb = static_cast<long long>(out_dim) * r_pad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < na + nb;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (i < na) {
      const int r = static_cast<int>(i / in_dim);
      a_pad[i] = __float2bfloat16(r < rank ? A[i] : 0.f);
    } else {
      const long long j = i - na;
      const int o = static_cast<int>(j / r_pad), r = static_cast<int>(j % r_pad);
      b_pad[j] = __float2bfloat16(r < rank ? Bm[static_cast<long long>(o) * rank + r] * scaling : 0.f);
    }
  }
}
int launch_lora_pack(const float* A, const float* Bm, int rank, int r_pad, int in_dim, int out_dim, float scaling, void* a_pad,
                     void* b_pad, cudaStream_t stream) {
  const long long n = static_cast<long long>(r_pad) * in_dim + static_cast<long long>(out_dim) * r_pad;
  const unsigned grid = static_cast<unsigned>((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256);
  lora_pack_kernel<<<grid, 256, 0, stream>>>(A, Bm, rank, r_pad, in_dim, out_dim, scaling, reinterpret_cast<__nv_bfloat16*>(a_pad),
                                             reinterpret_cast<__nv_bfloat16*>(b_pad));
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace al
