// Waveform ingest (§8f row 2): channel mean + polyphase windowed-sinc resampling to 16 kHz in one pass, output
// already in the [clips][480000] zero-padded layout the mel kernel reads.
//
// Replaces torch.mean(waveform, dim=0) + torchaudio.transforms.Resample of
// /root/reference/src/inference.py:87-93 and /root/reference/src/dataset.py:115-123 (torchaudio functional
// _get_sinc_resample_kernel / _apply_sinc_resample_kernel: sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99).
//
// torchaudio convolves every output phase with a dense filter of 2*width + orig taps, of which only the <= 2*width+2
// inside the Hann window's support are non-zero (34 of 475 at 44.1 -> 16 kHz). The host keeps, per phase, the first
// non-zero tap and the packed run of weights (built in fp64 like torchaudio, cast to fp32); one thread produces one
// output sample: y[n*new + p] = sum_i w[p][i] * x[n*orig + first[p] + i - width].
#include "common.cuh"
#include "kernels.h"

namespace al {

__global__ void __launch_bounds__(256)
ingest_resample_kernel(const float* __restrict__ in, long long clip_stride, long long chan_stride, int n_chan,
                       const int* __restrict__ n_in, int n_in_cap, ResampleTable tb, float* __restrict__ out,
                       long long out_stride, int out_cap, int* __restrict__ n_out) {
  const int b = blockIdx.y;
  const int len = min(n_in ? n_in[b] : n_in_cap, n_in_cap);           // valid input samples (train mode caps first)
  // ceil(new * len / orig) like torchaudio, then the 30 s cap of the caller
  const long long full = (static_cast<long long>(tb.new_f) * len + tb.orig_f - 1) / tb.orig_f;
  const int n_valid = static_cast<int>(full < out_cap ? full : out_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0 && n_out) n_out[b] = n_valid;
  const float* x = in + b * clip_stride;
  const float inv_c = 1.0f / n_chan;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < out_cap; j += gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (j < n_valid) {
      const int blk = j / tb.new_f, p = j - blk * tb.new_f;
      const int first = __ldg(tb.first + p);
      const float* w = tb.weights + static_cast<long long>(p) * tb.max_taps;
      const int s0 = blk * tb.orig_f + first - tb.width;               // input index of tap 0
      for (int i = 0; i < tb.max_taps; ++i) {
        const int s = s0 + i;
        if (s >= 0 && s < len) {
          float v = __ldg(x + s);
          for (int c = 1; c < n_chan; ++c) v += __ldg(x + c * chan_stride + s);
          acc = fmaf(__ldg(w + i), n_chan > 1 ? v * inv_c : v, acc);
        }
      }
    }
    out[b * out_stride + j] = acc;                                     // zero padding beyond n_valid
  }
}

// Same rate: channel mean + pad / truncate only.
__global__ void __launch_bounds__(256)
ingest_copy_kernel(const float* __restrict__ in, long long clip_stride, long long chan_stride, int n_chan,
                   const int* __restrict__ n_in, int n_in_cap, float* __restrict__ out, long long out_stride,
                   int out_cap, int* __restrict__ n_out) {
  const int b = blockIdx.y;
  const int len = min(min(n_in ? n_in[b] : n_in_cap, n_in_cap), out_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0 && n_out) n_out[b] = len;
  const float* x = in + b * clip_stride;
  const float inv_c = 1.0f / n_chan;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < out_cap; j += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (j < len) {
      v = __ldg(x + j);
      for (int c = 1; c < n_chan; ++c) v += __ldg(x + c * chan_stride + j);
      if (n_chan > 1) v *= inv_c;
    }
    out[b * out_stride + j] = v;
  }
}

int launch_ingest(const float* in, long long clip_stride, long long chan_stride, int n_chan, const int* n_in,
                  int n_in_cap, const ResampleTable* tb, float* out, long long out_stride, int out_cap, int* n_out,
                  int B, cudaStream_t stream) {
  if (B == 0) return 0;
  dim3 grid((out_cap + 255) / 256 > 592 ? 592 : (out_cap + 255) / 256, B);
  if (tb)
    ingest_resample_kernel<<<grid, 256, 0, stream>>>(in, clip_stride, chan_stride, n_chan, n_in, n_in_cap, *tb, out,
                                                     out_stride, out_cap, n_out);
  else
    ingest_copy_kernel<<<grid, 256, 0, stream>>>(in, clip_stride, chan_stride, n_chan, n_in, n_in_cap, out, out_stride,
                                                 out_cap, n_out);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace al
