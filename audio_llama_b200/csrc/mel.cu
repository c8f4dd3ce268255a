// Fused log-mel front end (fp32): framing (reflect pad) + Hann window + real FFT-400 + |.|^2 + sparse mel
// projection + log, one pass over the waveform.
//
// Replaces HF WhisperFeatureExtractor._torch_extract_fbank_features
// (feature_extraction_whisper.py:135-164: torch.stft -> abs()**2 -> mel_filters.T @ . -> clamp/log10 ->
// per-clip max-8 floor -> (x+4)/4) as called by /root/reference/src/inference.py:100-105, and, in mode 1,
// the torchaudio MelSpectrogram + log(x+1e-9) of /root/reference/src/dataset.py:125-133.
//
// One CTA = 8 consecutive frames of one clip (3000 = 375 x 8; 320 threads = one radix-5 butterfly per thread per
// stage; ~32 KB of shared memory so 6 CTAs share an SM — measured sweep in DESIGN.md: 24 frames / 256 threads
// 299 us, 8 / 320 162 us for 32 clips). The 1520 samples those frames touch are read from HBM once into shared
// memory (the 2.5x frame overlap is served from there). The real FFT of 400
// points is a 200-point complex FFT of (even + i*odd) samples — Stockham, radices 5,5,8, twiddles from an
// fp64-built table — plus the split post-pass. The mel filter bank is applied sparse (<= 9 non-zeros per mel
// bin, 394 in total at 128 bins; CSC built on the host in fp64 exactly as HF audio_utils.mel_filter_bank).
// The whisper mode's per-clip max goes through one atomicMax per CTA; the floor + affine is the light second
// kernel below (it has to wait for the whole clip).
#include "common.cuh"
#include "kernels.h"

namespace al {

#ifndef MEL_FR_CFG
#define MEL_FR_CFG 8
#endif
#ifndef MEL_THREADS_CFG
#define MEL_THREADS_CFG 320
#endif
constexpr int MEL_FR = MEL_FR_CFG;                          // frames per CTA (must divide 3000)
constexpr int MEL_NS = (MEL_FR - 1) * 160 + 400;            // 4080 samples staged per CTA
constexpr int MEL_THREADS = MEL_THREADS_CFG;
static_assert(3000 % MEL_FR == 0 && MEL_NS % 4 == 0, "frames per CTA must divide 3000");
constexpr int MEL_MAX_MELS = 256;
constexpr int MEL_MAX_NNZ = 1024;
constexpr int N_CLIP = 480000;
constexpr int N_FRAMES = 3000;

struct cf { float x, y; };
__device__ __forceinline__ cf cadd(cf a, cf b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cf csub(cf a, cf b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cf cmul(cf a, cf b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cf mul_neg_i(cf a) { return {a.y, -a.x}; }     // a * (-i)

__device__ __forceinline__ void dft5(cf* v) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  cf t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
  cf a1 = {v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y};
  cf a2 = {v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y};
  cf b1 = {s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y};
  cf b2 = {s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y};
  v[0] = {v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y};
  v[1] = {a1.x + b1.y, a1.y - b1.x};     // a1 - i b1
  v[4] = {a1.x - b1.y, a1.y + b1.x};     // a1 + i b1
  v[2] = {a2.x + b2.y, a2.y - b2.x};
  v[3] = {a2.x - b2.y, a2.y + b2.x};
}

__device__ __forceinline__ void dft8(cf* v) {
  const float r = 0.70710678118654752f;
  cf a0 = cadd(v[0], v[4]), a1 = csub(v[0], v[4]), a2 = cadd(v[2], v[6]), a3 = mul_neg_i(csub(v[2], v[6]));
  cf b0 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]), b2 = cadd(v[3], v[7]), b3 = mul_neg_i(csub(v[3], v[7]));
  cf e0 = cadd(a0, a2), e1 = cadd(a1, a3), e2 = csub(a0, a2), e3 = csub(a1, a3);
  cf o0 = cadd(b0, b2), o1 = cadd(b1, b3), o2 = csub(b0, b2), o3 = csub(b1, b3);
  o1 = {r * (o1.x + o1.y), r * (o1.y - o1.x)};          // * (1 - i)/sqrt2
  o2 = mul_neg_i(o2);                                   // * -i
  o3 = {r * (o3.y - o3.x), -r * (o3.x + o3.y)};         // * (-1 - i)/sqrt2
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

__device__ __forceinline__ unsigned int float_to_ordered(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void __launch_bounds__(MEL_THREADS)
mel_kernel(const float* __restrict__ wave, const int* __restrict__ n_samples, long long wave_stride, MelTables tb,
           int mode, float* __restrict__ out, unsigned int* __restrict__ clip_max_bits) {
  extern __shared__ uint8_t smem_raw[];
  float* sW = reinterpret_cast<float*>(smem_raw);                       // [MEL_NS] samples
  cf* bufA = reinterpret_cast<cf*>(sW + MEL_NS);                        // [MEL_FR][200]
  cf* bufB = bufA + MEL_FR * 200;                                       // [MEL_FR][200]
  float* sP = reinterpret_cast<float*>(bufA);                           // [MEL_FR][204] power, aliases bufA
  int* sColStart = reinterpret_cast<int*>(bufB + MEL_FR * 200);         // [n_mels + 1]
  int* sNzFreq = sColStart + MEL_MAX_MELS + 1;                          // [nnz]
  float* sNzW = reinterpret_cast<float*>(sNzFreq + MEL_MAX_NNZ);        // [nnz]
  __shared__ float s_red[MEL_THREADS / 32];

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * MEL_FR;
  const int tid = threadIdx.x;
  // samples beyond nv are the zero padding; never past the clip's own row (wave_stride) whatever n_samples says
  const int nv = max(0, static_cast<int>(min(static_cast<long long>(n_samples ? min(n_samples[b], N_CLIP) : N_CLIP), wave_stride)));
  const float* w = wave + static_cast<long long>(b) * wave_stride;

  // 1. stage the samples of frames f0..f0+FR-1: padded index p = 160*f0 + i  <->  clip index p - 200, reflected
  //    at both clip ends (torch.stft center=True, pad_mode="reflect") and zero beyond the clip's own samples.
  const int s_first = f0 * 160 - 200;
  if (s_first >= 0 && s_first + MEL_NS <= nv && (reinterpret_cast<uintptr_t>(w + s_first) & 15) == 0) {
    // interior CTA (all but the first and last two per clip): no reflection, no padding -> 128-bit loads
    // (s_first is a multiple of 8 samples: f0 is a multiple of MEL_FR = 8 frames of 160, minus 200)
    const float4* src = reinterpret_cast<const float4*>(w + s_first);
    float4* dst = reinterpret_cast<float4*>(sW);
    for (int i = tid; i < MEL_NS / 4; i += MEL_THREADS) dst[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < MEL_NS; i += MEL_THREADS) {
      int s = s_first + i;
      if (s < 0) s = -s;
      if (s >= N_CLIP) s = 2 * (N_CLIP - 1) - s;
      sW[i] = (s < nv) ? __ldg(w + s) : 0.f;
    }
  }
  // the sparse filter bank (CSC) moves to shared memory once per CTA: the mel stage below reads it ~3x per output
  for (int i = tid; i <= tb.n_mels; i += MEL_THREADS) sColStart[i] = __ldg(tb.col_start + i);
  for (int i = tid; i < tb.nnz; i += MEL_THREADS) {
    sNzFreq[i] = __ldg(tb.nz_freq + i);
    sNzW[i] = __ldg(tb.nz_w + i);
  }
  __syncthreads();

  // 2. stage 1 (radix 5, Ns = 1) straight from the windowed samples: z[n] = w[2n] x[2n] + i w[2n+1] x[2n+1]
  for (int t = tid; t < MEL_FR * 40; t += MEL_THREADS) {
    const int f = t / 40, j = t - f * 40;
    const float* x = sW + f * 160;
    cf v[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const int n = j + r * 40;
      const float2 win = *reinterpret_cast<const float2*>(tb.window + 2 * n);
      const float2 xs = *reinterpret_cast<const float2*>(x + 2 * n);
      v[r] = {win.x * xs.x, win.y * xs.y};
    }
    dft5(v);
    cf* o = bufB + f * 200 + j * 5;
#pragma unroll
    for (int r = 0; r < 5; ++r) o[r] = v[r];
  }
  __syncthreads();
  // stage 2 (radix 5, Ns = 5): twiddle exp(-2 pi i r k / 25) = tw200[8 r k]
  for (int t = tid; t < MEL_FR * 40; t += MEL_THREADS) {
    const int f = t / 40, j = t - f * 40;
    const int k = j % 5;
    const cf* in = bufB + f * 200;
    cf v[5];
    v[0] = in[j];
#pragma unroll
    for (int r = 1; r < 5; ++r) {
      const float2 tw = __ldg(tb.tw200 + 8 * r * k);
      v[r] = cmul(in[j + r * 40], cf{tw.x, tw.y});
    }
    dft5(v);
    cf* o = bufA + f * 200 + (j / 5) * 25 + k;
#pragma unroll
    for (int r = 0; r < 5; ++r) o[r * 5] = v[r];
  }
  __syncthreads();
  // stage 3 (radix 8, Ns = 25): twiddle exp(-2 pi i r k / 200) = tw200[r k]; output in natural order
  for (int t = tid; t < MEL_FR * 25; t += MEL_THREADS) {
    const int f = t / 25, k = t - f * 25;
    const cf* in = bufA + f * 200;
    cf v[8];
    v[0] = in[k];
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      const float2 tw = __ldg(tb.tw200 + r * k);
      v[r] = cmul(in[k + r * 25], cf{tw.x, tw.y});
    }
    dft8(v);
    cf* o = bufB + f * 200 + k;
#pragma unroll
    for (int r = 0; r < 8; ++r) o[r * 25] = v[r];
  }
  __syncthreads();
  // 3. split post-pass: X[k] = Xe[k] + W400^k Xo[k], X[200-k] = conj(Xe[k] - W400^k Xo[k]); power into sP
  for (int t = tid; t < MEL_FR * 101; t += MEL_THREADS) {
    const int f = t / 101, k = t - f * 101;
    const cf* Z = bufB + f * 200;
    const cf zk = Z[k];
    const cf zn = Z[k == 0 ? 0 : 200 - k];
    const cf xe = {0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};          // (Z[k] + conj Z[N-k]) / 2
    const cf dd = {0.5f * (zk.x - zn.x), 0.5f * (zk.y + zn.y)};          // (Z[k] - conj Z[N-k]) / 2
    const cf xo = mul_neg_i(dd);
    const float2 tw = __ldg(tb.tw400 + k);
    const cf tt = cmul(xo, cf{tw.x, tw.y});
    const cf p = cadd(xe, tt), q = csub(xe, tt);
    float* P = sP + f * 204;
    // sP aliases bufA (not bufB, which is still being read) — safe.
    P[k] = p.x * p.x + p.y * p.y;
    if (k != 100) P[200 - k] = q.x * q.x + q.y * q.y;
  }
  __syncthreads();
  // 4. sparse mel + log: one thread per mel bin for all MEL_FR frames — each filter weight / frequency index is
  //    read once and used for 8 frames, and the 8 results are one contiguous 32-byte store per mel row.
  float lmax = -INFINITY;
  const int n_mels = tb.n_mels;
  for (int m = tid; m < n_mels; m += MEL_THREADS) {
    float acc[MEL_FR];
#pragma unroll
    for (int f = 0; f < MEL_FR; ++f) acc[f] = 0.f;
    const int e = sColStart[m + 1];
    for (int i = sColStart[m]; i < e; ++i) {
      const float wgt = sNzW[i];
      const float* P = sP + sNzFreq[i];
#pragma unroll
      for (int f = 0; f < MEL_FR; ++f) acc[f] = fmaf(wgt, P[f * 204], acc[f]);
    }
    float v[MEL_FR];
#pragma unroll
    for (int f = 0; f < MEL_FR; ++f) {
      if (mode == 0) {
        // log10 via MUFU lg2 (abs. error ~1e-7 on values in [-10, 5]; the parity budget is 1e-5 * max)
        v[f] = __log2f(fmaxf(acc[f], 1e-10f)) * 0.30102999566398120f;
        lmax = fmaxf(lmax, v[f]);
      } else {
        v[f] = logf(acc[f] + 1e-9f);
      }
    }
    float* orow = out + (static_cast<long long>(b) * n_mels + m) * N_FRAMES + f0;   // f0 % 8 == 0: 32 B aligned
    static_assert(MEL_FR % 4 == 0, "frames per CTA must be a multiple of 4 for the vector stores");
#pragma unroll
    for (int f = 0; f < MEL_FR; f += 4) *reinterpret_cast<float4*>(orow + f) = make_float4(v[f], v[f + 1], v[f + 2], v[f + 3]);
  }
  if (mode == 0) {
    lmax = warp_max(lmax);
    if ((tid & 31) == 0) s_red[tid >> 5] = lmax;
    __syncthreads();
    if (tid == 0) {
      float mx = s_red[0];
      for (int i = 1; i < MEL_THREADS / 32; ++i) mx = fmaxf(mx, s_red[i]);
      atomicMax(clip_max_bits + b, float_to_ordered(mx));
    }
  }
}

// (max(L, clipmax - 8) + 4) / 4, in place (HF :156-159).
__global__ void mel_finalize_kernel(float* __restrict__ out, const unsigned int* __restrict__ clip_max_bits,
                                    long long per_clip4) {
  const int b = blockIdx.y;
  const float floor_v = ordered_to_float(clip_max_bits[b]) - 8.0f;
  float4* p = reinterpret_cast<float4*>(out) + static_cast<long long>(b) * per_clip4;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < per_clip4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 v = p[i];
    v.x = (fmaxf(v.x, floor_v) + 4.0f) / 4.0f;
    v.y = (fmaxf(v.y, floor_v) + 4.0f) / 4.0f;
    v.z = (fmaxf(v.z, floor_v) + 4.0f) / 4.0f;
    v.w = (fmaxf(v.w, floor_v) + 4.0f) / 4.0f;
    p[i] = v;
  }
}

int launch_mel(const float* wave, const int* n_samples, int B, long long wave_stride, const MelTables& tb, int mode,
               float* out, unsigned int* clip_max_bits, cudaStream_t stream) {
  constexpr int smem = MEL_NS * 4 + 2 * MEL_FR * 200 * 8 + (MEL_MAX_MELS + 1) * 4 + MEL_MAX_NNZ * 8;
  static bool attr_set = false;
  if (!attr_set) {
    AL_CHECK_CUDA(cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  if (mode == 0) AL_CHECK_CUDA(cudaMemsetAsync(clip_max_bits, 0, sizeof(unsigned int) * B, stream));
  dim3 grid(N_FRAMES / MEL_FR, B);
  mel_kernel<<<grid, MEL_THREADS, smem, stream>>>(wave, n_samples, wave_stride, tb, mode, out, clip_max_bits);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_mel_finalize(float* out, const unsigned int* clip_max_bits, int B, int n_mels, cudaStream_t stream) {
  const long long per_clip4 = static_cast<long long>(n_mels) * N_FRAMES / 4;
  dim3 grid(static_cast<unsigned>((per_clip4 + 255) / 256 > 148 ? 148 : (per_clip4 + 255) / 256), B);
  mel_finalize_kernel<<<grid, 256, 0, stream>>>(out, clip_max_bits, per_clip4);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace al
