// Row / element kernels of the LLaMA side (first slice): RMSNorm, SwiGLU, rotary embedding, and the
// cross-entropy that follows lm_head — forward and backward, bf16 in / out, fp32 arithmetic. All HBM-bound: one pass
// over the data with 16-byte accesses, one warp per row for the norms, grids sized by the data.
//
// They replace, inside the HF LlamaForCausalLM the reference drives (/root/reference/src/models/allm.py:99-104 ->
// HF models/llama/modeling_llama.py): LlamaRMSNorm.forward (x.float() -> pow -> mean -> rsqrt -> mul -> cast -> mul:
// 7 elementwise kernels + their autograd), `act_fn(gate_proj(x)) * up_proj(x)` of LlamaMLP.forward,
// apply_rotary_pos_emb / rotate_half (mul, slice, neg, cat, mul, add per tensor), and the loss of
// LlamaForCausalLM.forward (logits.float(), shift, cross_entropy with ignore_index = -100, mean over valid tokens).
#include "common.cuh"
#include "kernels.h"

namespace al {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ----------------------------------------------------------------------------- RMSNorm
// y = w * bf16(x * rsqrt(mean(x^2) + eps))   (HF rounds the normalised value to the input dtype before the weight)
// One warp per row, the row kept in registers (MAXV 16-byte vectors per lane; d <= 256 * MAXV).
template <int MAXV>
__global__ void __launch_bounds__(256, (MAXV <= 12) ? 4 : 2)
rmsnorm_fwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ w, uint4* __restrict__ y, float* __restrict__ rstd_out,
                   int rows, int nvec, float inv_d, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const uint4* xr = x + static_cast<long long>(row) * nvec;
  uint4 v[MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      v[i] = __ldg(xr + c);
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) ss = fmaf(f[k], f[k], ss);
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss * inv_d + eps);
  if (lane == 0 && rstd_out) rstd_out[row] = rstd;
  uint4* yr = y + static_cast<long long>(row) * nvec;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      float f[8], g[8];
      unpack8(v[i], f);
      unpack8(__ldg(w + c), g);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = g[k] * bf16_round(f[k] * rstd);
      yr[c] = pack8(f);
    }
  }
}

// dx = rstd * (g - xhat * mean(g * xhat)), g = dy * w, xhat = x * rstd (the weight is frozen: no dw), + addend when given
// (the gradient that reaches x through the residual connection: one pass instead of autograd's separate add).
template <int MAXV>
__global__ void __launch_bounds__(256, (MAXV <= 12) ? 3 : 2)
rmsnorm_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ w, const float* __restrict__ rstd_in,
                   const uint4* __restrict__ dy, const uint4* addend, uint4* dx, int rows, int nvec, float inv_d) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long base = static_cast<long long>(row) * nvec;
  const float rstd = rstd_in[row];
  uint4 xv[MAXV];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      xv[i] = __ldg(x + base + c);
      float xf[8], df[8], wf[8];
      unpack8(xv[i], xf);
      unpack8(__ldg(dy + base + c), df);
      unpack8(__ldg(w + c), wf);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        df[k] *= wf[k];
        dot = fmaf(df[k], xf[k] * rstd, dot);
      }
    }
  }
  dot = warp_sum(dot) * inv_d;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      float xf[8], df[8], wf[8], o[8];
      unpack8(xv[i], xf);
      unpack8(__ldg(dy + base + c), df);       // L1 / L2 hit: read a few hundred cycles ago by this very lane (keeping
                                               // dy * w in registers as well would double the register file)
      unpack8(__ldg(w + c), wf);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = rstd * (df[k] * wf[k] - xf[k] * rstd * dot);
      if (addend != nullptr) {                 // (may be dx itself: each element is read and written by this lane only)
        float af[8];
        unpack8(addend[base + c], af);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += af[k];
      }
      dx[base + c] = pack8(o);
    }
  }
}

template <int MAXV>
static int rmsnorm_launch(const void* x, const void* w, void* y, float* rstd, const void* dy, const void* addend, void* dx, int rows,
                          int d, float eps, bool backward, cudaStream_t st) {
  const int nvec = d / 8;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (!backward)
    rmsnorm_fwd_kernel<MAXV><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(w),
                                                    reinterpret_cast<uint4*>(y), rstd, rows, nvec, 1.0f / d, eps);
  else
    rmsnorm_bwd_kernel<MAXV><<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(w), rstd,
                                                    reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(addend),
                                                    reinterpret_cast<uint4*>(dx), rows, nvec, 1.0f / d);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_rmsnorm(const void* x, const void* w, void* y, float* rstd, const void* dy, const void* addend, void* dx, int rows, int d,
                   float eps, bool backward, cudaStream_t st) {
  AL_REQUIRE(d % 8 == 0 && d <= 8192, "rmsnorm: d=%d must be a multiple of 8 and <= 8192", d);
  if (rows == 0) return 0;
  if (d <= 2048) return rmsnorm_launch<8>(x, w, y, rstd, dy, addend, dx, rows, d, eps, backward, st);
  if (d <= 3072) return rmsnorm_launch<12>(x, w, y, rstd, dy, addend, dx, rows, d, eps, backward, st);   // fewer registers: 4 CTAs / SM
  if (d <= 4096) return rmsnorm_launch<16>(x, w, y, rstd, dy, addend, dx, rows, d, eps, backward, st);
  return rmsnorm_launch<32>(x, w, y, rstd, dy, addend, dx, rows, d, eps, backward, st);
}

// ----------------------------------------------------------------------------- SwiGLU
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const uint4* __restrict__ gate, const uint4* __restrict__ up, uint4* __restrict__ h, long long nvec) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < nvec; i += gridDim.x * 256LL) {
    float g[8], u[8];
    unpack8(__ldg(gate + i), g);
    unpack8(__ldg(up + i), u);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = g[k] / (1.0f + __expf(-g[k])) * u[k];
    h[i] = pack8(g);
  }
}
// h = silu(g) u: dg = dh u s (1 + g (1 - s)), du = dh g s, s = sigmoid(g)
__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const uint4* __restrict__ gate, const uint4* __restrict__ up, const uint4* __restrict__ dh,
                  uint4* __restrict__ dgate, uint4* __restrict__ dup, long long nvec) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < nvec; i += gridDim.x * 256LL) {
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8(__ldg(gate + i), g);
    unpack8(__ldg(up + i), u);
    unpack8(__ldg(dh + i), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float s = 1.0f / (1.0f + __expf(-g[k]));
      dg[k] = d[k] * u[k] * s * (1.0f + g[k] * (1.0f - s));
      du[k] = d[k] * g[k] * s;
    }
    dgate[i] = pack8(dg);
    dup[i] = pack8(du);
  }
}
int launch_swiglu(const void* gate, const void* up, const void* dh, void* out0, void* out1, long long n, bool backward,
                  int num_sms, cudaStream_t st) {
  AL_REQUIRE(n % 8 == 0, "swiglu: element count %lld must be a multiple of 8", n);
  if (n == 0) return 0;
  const long long nvec = n / 8;
  const long long want = (nvec + 255) / 256;
  const unsigned grid = static_cast<unsigned>(want < 16LL * num_sms ? want : 16LL * num_sms);
  if (!backward)
    swiglu_fwd_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(gate), reinterpret_cast<const uint4*>(up),
                                            reinterpret_cast<uint4*>(out0), nvec);
  else
    swiglu_bwd_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(gate), reinterpret_cast<const uint4*>(up),
                                            reinterpret_cast<const uint4*>(dh), reinterpret_cast<uint4*>(out0),
                                            reinterpret_cast<uint4*>(out1), nvec);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------------------- rotary embedding
// x [B][S][H][hd] (the projection output viewed per head), cos / sin [Bc][S][hd] (Bc = 1 broadcasts).
// forward : o1 = x1 c1 - x2 s1 ; o2 = x2 c2 + x1 s2        (q cos + rotate_half(q) sin)
// backward: d1 = g1 c1 + g2 s2 ; d2 = g2 c2 - g1 s1
// One thread per 8 + 8 elements (a 16-byte vector of each half).
__global__ void __launch_bounds__(256)
rope_kernel(const uint4* __restrict__ x, const uint4* __restrict__ cs, const uint4* __restrict__ sn, uint4* __restrict__ out,
            long long n_items, int S, int H, int hv /* hd / 16 */, long long cos_batch_stride_vec, int backward) {
  for (long long it = blockIdx.x * 256LL + threadIdx.x; it < n_items; it += gridDim.x * 256LL) {
    const int j = static_cast<int>(it % hv);
    const long long bsh = it / hv;                    // (b * S + s) * H + h
    const long long bs = bsh / H;
    const int s = static_cast<int>(bs % S);
    const long long b = bs / S;
    const long long xo = bsh * (2 * hv) + j;
    const long long co = b * cos_batch_stride_vec + static_cast<long long>(s) * (2 * hv) + j;
    float x1[8], x2[8], c1[8], c2[8], s1[8], s2[8], o1[8], o2[8];
    unpack8(__ldg(x + xo), x1);
    unpack8(__ldg(x + xo + hv), x2);
    unpack8(__ldg(cs + co), c1);
    unpack8(__ldg(cs + co + hv), c2);
    unpack8(__ldg(sn + co), s1);
    unpack8(__ldg(sn + co + hv), s2);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (!backward) {
        o1[k] = x1[k] * c1[k] - x2[k] * s1[k];
        o2[k] = x2[k] * c2[k] + x1[k] * s2[k];
      } else {
        o1[k] = x1[k] * c1[k] + x2[k] * s2[k];
        o2[k] = x2[k] * c2[k] - x1[k] * s1[k];
      }
    }
    out[xo] = pack8(o1);
    out[xo + hv] = pack8(o2);
  }
}
int launch_rope(const void* x, const void* cs, const void* sn, void* out, int B, int S, int H, int hd, int cos_batch,
                int backward, int num_sms, cudaStream_t st) {
  AL_REQUIRE(hd % 16 == 0 && hd > 0, "rope: head_dim=%d must be a multiple of 16", hd);
  AL_REQUIRE(cos_batch == 1 || cos_batch == B, "rope: cos/sin batch %d must be 1 or %d", cos_batch, B);
  const int hv = hd / 16;
  const long long n_items = static_cast<long long>(B) * S * H * hv;
  if (n_items == 0) return 0;
  const long long want = (n_items + 255) / 256;
  const unsigned grid = static_cast<unsigned>(want < 16LL * num_sms ? want : 16LL * num_sms);
  rope_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(cs),
                                    reinterpret_cast<const uint4*>(sn), reinterpret_cast<uint4*>(out), n_items, S, H, hv,
                                    cos_batch == 1 ? 0 : static_cast<long long>(S) * 2 * hv, backward);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------------------- cross-entropy on a chunk of logits
// logits [rows][ld] bf16 (vocab valid columns), labels [rows] (already shifted; -100 = ignored). Per row:
// lse = log sum exp; loss_sum += lse - logit[label]; the row is overwritten with its gradient
// (softmax - onehot) * grad_scale (zero for ignored rows). One CTA per row: an online (max, sum) pass, then the
// write pass re-reads the row (256 KB at LLaMA's vocabulary: an L2 hit).
__global__ void __launch_bounds__(512)
ce_inplace_kernel(__nv_bfloat16* __restrict__ logits, const long long* __restrict__ labels, int vocab, long long ld,
                  float grad_scale, float* __restrict__ loss_sum) {
  __shared__ float s_m[16], s_s[16];
  __shared__ float s_lse;
  const int row = blockIdx.x;
  __nv_bfloat16* lr = logits + static_cast<long long>(row) * ld;
  const long long lab = labels[row];
  const int nvec = vocab / 8;
  uint4* lv = reinterpret_cast<uint4*>(lr);
  if (lab < 0 || lab >= vocab) {                   // ignored position: gradient 0, no loss term
    for (int i = threadIdx.x; i < nvec; i += 512) lv[i] = make_uint4(0, 0, 0, 0);
    for (int i = nvec * 8 + threadIdx.x; i < vocab; i += 512) lr[i] = __float2bfloat16_rn(0.f);
    // a label >= vocab is an error (torch's cross_entropy asserts on it): nothing outside the row is read, and the
    // loss is poisoned with NaN so that the step fails loudly instead of training on garbage
    if (lab >= vocab && threadIdx.x == 0) atomicAdd(loss_sum, __int_as_float(0x7fc00000));
    return;
  }
  float m = -INFINITY, s = 0.f;
  auto push = [&](float v) {
    if (v > m) {
      s = s * __expf(m - v) + 1.0f;
      m = v;
    } else {
      s += __expf(v - m);
    }
  };
  for (int i = threadIdx.x; i < nvec; i += 512) {
    float f[8];
    unpack8(lv[i], f);
    float vm = f[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) vm = fmaxf(vm, f[k]);
    if (vm > m) {
      s *= __expf(m - vm);
      m = vm;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s += __expf(f[k] - m);
  }
  for (int i = nvec * 8 + threadIdx.x; i < vocab; i += 512) push(__bfloat162float(lr[i]));
  // combine (m, s) across the CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mn)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mn));
    m = mn;
  }
  if ((threadIdx.x & 31) == 0) {
    s_m[threadIdx.x >> 5] = m;
    s_s[threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = s_m[0], ss = s_s[0];
    for (int i = 1; i < 16; ++i) {
      const float mn = fmaxf(mm, s_m[i]);
      ss = (mm == -INFINITY ? 0.f : ss * __expf(mm - mn)) + (s_m[i] == -INFINITY ? 0.f : s_s[i] * __expf(s_m[i] - mn));
      mm = mn;
    }
    const float lse = mm + logf(ss);
    s_lse = lse;
    atomicAdd(loss_sum, lse - __bfloat162float(lr[lab]));
  }
  __syncthreads();
  const float lse = s_lse;
  for (int i = threadIdx.x; i < nvec; i += 512) {
    float f[8];
    unpack8(lv[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = __expf(f[k] - lse) * grad_scale;
    const long long c0 = static_cast<long long>(i) * 8;
    if (lab >= c0 && lab < c0 + 8) f[lab - c0] -= grad_scale;
    lv[i] = pack8(f);
  }
  for (int i = nvec * 8 + threadIdx.x; i < vocab; i += 512) {
    float p = __expf(__bfloat162float(lr[i]) - lse) * grad_scale;
    if (i == lab) p -= grad_scale;
    lr[i] = __float2bfloat16_rn(p);
  }
}
int launch_ce_inplace(void* logits, const long long* labels, int rows, int vocab, long long ld, float grad_scale,
                      float* loss_sum, cudaStream_t st) {
  AL_REQUIRE(ld % 8 == 0 && ld >= vocab, "cross-entropy: row pitch %lld must be a multiple of 8 and >= vocab %d", ld, vocab);
  if (rows == 0) return 0;
  ce_inplace_kernel<<<rows, 512, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(logits), labels, vocab, ld, grad_scale, loss_sum);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace al
