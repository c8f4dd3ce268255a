// Causal grouped-query attention for the LLaMA side of the training step, head_dim 128, forward and
// backward, on sm_100a (tcgen05 / TMEM / TMA). Replaces the scaled_dot_product_attention call that HF's LlamaAttention
// makes on the tensors the reference drives through it (/root/reference/src/models/allm.py:99-104 ->
// transformers/models/llama/modeling_llama.py LlamaAttention.forward -> attention_interface), forward AND autograd.
//
//   q  [B][S][Hq ][128] bf16 (after RoPE)      out  [B][S][Hq][128] bf16
//   k  [B][S][Hkv][128] bf16 (after RoPE)      lse  [B][Hq][S] f32   (log2 domain: log2 sum_j 2^(s_ij * scale * log2 e))
//   v  [B][S][Hkv][128] bf16                   kv_len [B] int32 or NULL: keys >= kv_len[b] are masked (right padding)
// Mask = causal AND key < kv_len[b]; query head h reads kv head h / (Hq / Hkv).
//
// Three kernels, one (128-row tile, head, batch) work item per CTA, one CTA per SM (192 KB of shared memory):
//   gqa_fwd_kernel      S = Q K_j^T (SS MMA, double-buffered in TMEM) -> online softmax, two threads per query row
//                       (exact row maximum exchanged through shared memory, lazy rescale of O) -> P bf16 in TMEM ->
//                       O += P V_j (A from TMEM, V MN-major). Writes O and the row log-sum-exp.
//   gqa_bwd_dq_kernel   per query tile, over kv tiles j <= i: S and dP = dO V_j^T (two SS MMAs), P = 2^(S c - lse),
//                       dS = P (dP - D) scale -> bf16 in TMEM -> dQ += dS K_j (A from TMEM, K MN-major).
//   gqa_bwd_dkv_kernel  per kv tile, over the group's query heads and query tiles i >= j, everything TRANSPOSED so that
//                       the products that contract over queries take their A operand from TMEM: S^T = K_j Q_i^T,
//                       dP^T = V_j dO_i^T, then dV += P^T dO_i and dK += dS^T Q_i (dO / Q as MN-major B operands).
//                       P^T / dS^T overwrite the first 64 columns of S^T / dP^T (the next tile's S^T is issued after the
//                       products that read them: tensor-core operations of one thread execute in issue order).
// D = rowsum(dO * O) comes from gqa_rowdot_kernel. Scores are recomputed in the backward (nothing but O and lse is saved).
#include "common.cuh"
#include "kernels.h"

namespace al {

constexpr int GQ_T = 128;                      // tile rows (queries or keys) and head_dim
constexpr int GQ_TILE_BYTES = GQ_T * GQ_T * 2; // 32 KB: one 128 x 128 bf16 tile = two [128][64] SW128 boxes
constexpr int GQ_BOX_BYTES = GQ_T * 64 * 2;    // 16 KB
constexpr int GQ_THREADS = 384;                // warps 0-3 control (TMA, MMA, TMEM allocator, idle), warps 4-11 compute
constexpr float GQ_TAU = 8.0f;                 // lazy-rescale threshold of the forward, log2 units
constexpr int GQ_BAR_EXCH = 1;                 // named barrier of the 256 compute threads
constexpr float LOG2E_GQ = 1.4426950408889634f;

__device__ __forceinline__ void tma_load_4d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// One [128 rows][128 d] tile of a [B][S][H][128] tensor -> two SW128 boxes (columns 0..63 | 64..127) on one barrier.
__device__ __forceinline__ void gq_load_tile(uint32_t dst, const CUtensorMap* m, uint32_t bar, int h, int s0, int b) {
  mbar_arrive_expect_tx_a(bar, GQ_TILE_BYTES);
  tma_load_4d_a(dst, m, bar, 0, h, s0, b);
  tma_load_4d_a(dst + GQ_BOX_BYTES, m, bar, 64, h, s0, b);
}
// D[tmem 128 x 128] (+)= A[smem tile, K-major over d] * B[smem tile, K-major over d]^T   (contraction over head_dim)
__device__ __forceinline__ void gq_mma_kk(uint32_t d_tmem, uint32_t a_tile, uint32_t b_tile, bool accumulate) {
  constexpr uint32_t IDESC = umma_idesc_bf16(GQ_T, GQ_T);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t off = (k >> 2) * GQ_BOX_BYTES + (k & 3) * 32;
    umma_ss(d_tmem, umma_desc_sw128(a_tile + off, 16, 1024), umma_desc_sw128(b_tile + off, 16, 1024), IDESC,
            (k != 0) || accumulate);
  }
}
// D[tmem 128 x 128] (+)= A[tmem: 128 lanes x 128 bf16 = 64 columns] * B[smem tile [rows = contraction][128 cols], MN-major]
__device__ __forceinline__ void gq_mma_tm(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_tile, bool accumulate) {
  constexpr uint32_t IDESC = umma_idesc_bf16(GQ_T, GQ_T, 0, 1);
#pragma unroll
  for (int k = 0; k < 8; ++k)   // 16 contraction rows = 2048 B per step; the two 64-column atoms are 16 KB apart (LBO)
    umma_ts(d_tmem, a_tmem + k * 8, umma_desc_sw128(b_tile + 2048 * k, GQ_BOX_BYTES, 1024), IDESC, (k != 0) || accumulate);
}

// ============================================================================ forward
constexpr uint32_t GF_OFF_Q = 0;
constexpr uint32_t GF_OFF_K = GQ_TILE_BYTES;
constexpr uint32_t GF_OFF_V = GF_OFF_K + 2 * GQ_TILE_BYTES;
constexpr uint32_t GF_OFF_EXCH = GF_OFF_V + 2 * GQ_TILE_BYTES;       // [2 parities][2 halves][128] f32 row maxima, then [2][128] sums
constexpr uint32_t GF_OFF_BARS = GF_OFF_EXCH + 3 * 2 * GQ_T * 4;
constexpr uint32_t GF_Q_FULL = GF_OFF_BARS, GF_K_FULL = GF_Q_FULL + 8, GF_K_EMPTY = GF_K_FULL + 16, GF_V_FULL = GF_K_EMPTY + 16,
                   GF_V_EMPTY = GF_V_FULL + 16, GF_S_FULL = GF_V_EMPTY + 16, GF_S_EMPTY = GF_S_FULL + 16, GF_P_FULL = GF_S_EMPTY + 16,
                   GF_O_FULL = GF_P_FULL + 8, GF_TMEM_PTR = GF_O_FULL + 8;
constexpr int GF_SMEM = GF_TMEM_PTR + 16 + 1024;

__global__ void __launch_bounds__(GQ_THREADS, 1)
gqa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
               const int* __restrict__ kv_len, int S, int Hq, int Hkv, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = opaque_u32((smem_u32(smem_raw) + 1023u) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heavy tiles first: query tile qt visits qt + 1 kv tiles (causal), so the block order runs from the last tile down
  const int qt = gridDim.x - 1 - blockIdx.x, hq = blockIdx.y, b = blockIdx.z;
  const int hkv = hq / (Hq / Hkv);
  const int kvl = max(1, min(S, kv_len ? kv_len[b] : S));
  const int n_tiles = min(qt + 1, (kvl + GQ_T - 1) / GQ_T);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init_a(sb + GF_Q_FULL, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init_a(sb + GF_K_FULL + 8 * s, 1);
      mbar_init_a(sb + GF_K_EMPTY + 8 * s, 1);
      mbar_init_a(sb + GF_V_FULL + 8 * s, 1);
      mbar_init_a(sb + GF_V_EMPTY + 8 * s, 1);
      mbar_init_a(sb + GF_S_FULL + 8 * s, 1);
      mbar_init_a(sb + GF_S_EMPTY + 8 * s, 256);
    }
    mbar_init_a(sb + GF_P_FULL, 256);
    mbar_init_a(sb + GF_O_FULL, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + GF_TMEM_PTR), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sb + GF_TMEM_PTR) : "memory");
  // TMEM columns: S0 0..127 | S1 128..255 | O 256..383 | P 384..447 (bf16 pairs)

  if (warp == 0) {
    if (elect_one()) {                                     // ---------------- TMA producer
      gq_load_tile(sb + GF_OFF_Q, &tmQ, sb + GF_Q_FULL, hq, qt * GQ_T, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait_a(sb + GF_K_EMPTY + 8 * s, ((j >> 1) - 1) & 1);
        gq_load_tile(sb + GF_OFF_K + s * GQ_TILE_BYTES, &tmK, sb + GF_K_FULL + 8 * s, hkv, j * GQ_T, b);
        if (j >= 2) mbar_wait_a(sb + GF_V_EMPTY + 8 * s, ((j >> 1) - 1) & 1);
        gq_load_tile(sb + GF_OFF_V + s * GQ_TILE_BYTES, &tmV, sb + GF_V_FULL + 8 * s, hkv, j * GQ_T, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: S_j one tile ahead of P V_{j-1}
    mbar_wait_a(sb + GF_Q_FULL, 0);
    auto issue_qk = [&](int j) {
      const int s = j & 1;
      mbar_wait_a(sb + GF_K_FULL + 8 * s, (j >> 1) & 1);
      if (j >= 2) mbar_wait_a(sb + GF_S_EMPTY + 8 * s, ((j >> 1) - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        gq_mma_kk(tmem_base + s * 128, sb + GF_OFF_Q, sb + GF_OFF_K + s * GQ_TILE_BYTES, false);
        umma_commit_a(sb + GF_K_EMPTY + 8 * s);
        umma_commit_a(sb + GF_S_FULL + 8 * s);
      }
      __syncwarp();
    };
    issue_qk(0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) issue_qk(j + 1);
      const int s = j & 1;
      mbar_wait_a(sb + GF_V_FULL + 8 * s, (j >> 1) & 1);
      mbar_wait_a(sb + GF_P_FULL, j & 1);
      tc_fence_after();
      if (elect_one()) {
        gq_mma_tm(tmem_base + 256, tmem_base + 384, sb + GF_OFF_V + s * GQ_TILE_BYTES, j != 0);
        umma_commit_a(sb + GF_V_EMPTY + 8 * s);
        umma_commit_a(sb + GF_O_FULL);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax: two threads per query row
    const int half = (warp - 4) >> 2;
    const uint32_t row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((row & ~31u) << 16);
    const int q_glob = qt * GQ_T + static_cast<int>(row);
    const uint32_t ex_mine = sb + GF_OFF_EXCH + half * (GQ_T * 4) + row * 4;
    const uint32_t ex_other = ex_mine ^ (GQ_T * 4);
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int sbuf = j & 1;
      mbar_wait_a(sb + GF_S_FULL + 8 * sbuf, (j >> 1) & 1);
      tc_fence_after();
      uint32_t s[64];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        tmem_ld_32x32(tlane + sbuf * 128 + half * 64, s0);
        tmem_ld_32x32(tlane + sbuf * 128 + half * 64 + 32, s1);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive_a(sb + GF_S_EMPTY + 8 * sbuf);
      const int col0 = j * GQ_T + half * 64;
      if (j == qt || col0 + 64 > kvl) {                    // diagonal tile or the tile holding the padding boundary
#pragma unroll
        for (int k = 0; k < 64; ++k)
          if (col0 + k > q_glob || col0 + k >= kvl) s[k] = 0xff800000u;
      }
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
#pragma unroll
      for (int k = 2; k < 64; k += 4) {
        mx0 = fmax3(mx0, __uint_as_float(s[k]), __uint_as_float(s[k + 1]));
        if (k + 2 < 64) mx1 = fmax3(mx1, __uint_as_float(s[k + 2]), __uint_as_float(s[k + 3]));
      }
      sts_f32(ex_mine + sbuf * (2 * GQ_T * 4), fmaxf(mx0, mx1));
      named_bar_sync(GQ_BAR_EXCH, 256);
      const float tile_max = fmaxf(fmaxf(mx0, mx1), lds_f32(ex_other + sbuf * (2 * GQ_T * 4))) * scale_log2;
      float alpha = 1.0f;
      const bool rescale = __any_sync(0xffffffffu, tile_max > m_ref + GQ_TAU);
      if (rescale && tile_max > m_ref + GQ_TAU) {          // (per row; the TMEM traffic below stays warp-uniform)
        alpha = ex2f(m_ref - tile_max);                    // 0 on the first tile
        m_ref = tile_max;
        l *= alpha;
      }
      const unsigned long long c2 = pk2(scale_log2, scale_log2), nm2 = pk2(-m_ref, -m_ref);
      unsigned long long l2a = pk2(0.f, 0.f), l2b = pk2(0.f, 0.f);
      uint32_t pk[32];
#pragma unroll
      for (int k = 0; k < 64; k += 2) {
        const unsigned long long x2 = ffma2(pk2(__uint_as_float(s[k]), __uint_as_float(s[k + 1])), c2, nm2);
        float x0, x1;
        unpk2(x2, x0, x1);
        const float p0 = ex2f(x0), p1 = ex2f(x1);
        if ((k >> 1) & 1) l2b = fadd2(l2b, pk2(p0, p1));
        else l2a = fadd2(l2a, pk2(p0, p1));
        pk[k >> 1] = pack_bf16(p0, p1);
      }
      {
        float a, c, e, f;
        unpk2(l2a, a, c);
        unpk2(l2b, e, f);
        l += (a + c) + (e + f);
      }
      if (j > 0) {                                         // the previous P V has read P and finished its part of O
        mbar_wait_a(sb + GF_O_FULL, (j - 1) & 1);
        tc_fence_after();
        if (rescale) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tlane + 256 + half * 64 + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * alpha);
            tmem_st_32x32(tlane + 256 + half * 64 + c * 32, r);
          }
        }
      }
      {
        uint32_t(&p0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&pk[0]);
        uint32_t(&p1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&pk[16]);
        tmem_st_32x16(tlane + 384 + half * 32, p0);
        tmem_st_32x16(tlane + 384 + half * 32 + 16, p1);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(sb + GF_P_FULL);
    }
    // row sum of both halves, normalise, store O (this half's 64 columns) and the log-sum-exp
    sts_f32(ex_mine + 2 * (2 * GQ_T * 4), l);
    named_bar_sync(GQ_BAR_EXCH, 256);
    const float l_tot = l + lds_f32(ex_other + 2 * (2 * GQ_T * 4));
    const float inv_l = 1.0f / l_tot;
    mbar_wait_a(sb + GF_O_FULL, (n_tiles - 1) & 1);
    tc_fence_after();
    if (q_glob < S) {
      if (half == 0) lse[(static_cast<size_t>(b) * Hq + hq) * S + q_glob] = m_ref + log2f(l_tot);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(b) * S + q_glob) * Hq + hq) * GQ_T + half * 64);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tlane + 256 + half * 64 + c * 32, r);
      tmem_ld_wait();
      if (q_glob < S) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(r[8 * u]) * inv_l, __uint_as_float(r[8 * u + 1]) * inv_l);
          v.y = pack_bf16(__uint_as_float(r[8 * u + 2]) * inv_l, __uint_as_float(r[8 * u + 3]) * inv_l);
          v.z = pack_bf16(__uint_as_float(r[8 * u + 4]) * inv_l, __uint_as_float(r[8 * u + 5]) * inv_l);
          v.w = pack_bf16(__uint_as_float(r[8 * u + 6]) * inv_l, __uint_as_float(r[8 * u + 7]) * inv_l);
          dst[c * 4 + u] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

int launch_gqa_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, void* out, float* lse,
                   const int* kv_len, int B, int S, int Hq, int Hkv, float scale, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    AL_CHECK_CUDA(cudaFuncSetAttribute(gqa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM));
    attr_set = true;
  }
  dim3 grid((S + GQ_T - 1) / GQ_T, Hq, B);
  gqa_fwd_kernel<<<grid, GQ_THREADS, GF_SMEM, stream>>>(tq, tk, tv, reinterpret_cast<__nv_bfloat16*>(out), lse, kv_len, S, Hq,
                                                        Hkv, scale * 1.4426950408889634f);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// ============================================================================ backward
__device__ __forceinline__ void gq_issue_tile(uint32_t dst, const CUtensorMap* m, uint32_t bar, int h, int s0, int b) {
  tma_load_4d_a(dst, m, bar, 0, h, s0, b);
  tma_load_4d_a(dst + GQ_BOX_BYTES, m, bar, 64, h, s0, b);
}

// D[b][h][s] = sum_d dO[b][s][h][d] * O[b][s][h][d]   (one warp per row)
__global__ void __launch_bounds__(256)
gqa_rowdot_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, float* __restrict__ dsum,
                  int B, int S, int H) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= static_cast<long long>(B) * S * H) return;
  const int lane = threadIdx.x & 31;
  const uint2 a = reinterpret_cast<const uint2*>(o + row * GQ_T)[lane];
  const uint2 g = reinterpret_cast<const uint2*>(d_o + row * GQ_T)[lane];
  float acc = __uint_as_float(a.x << 16) * __uint_as_float(g.x << 16);
  acc = fmaf(__uint_as_float(a.x & 0xffff0000u), __uint_as_float(g.x & 0xffff0000u), acc);
  acc = fmaf(__uint_as_float(a.y << 16), __uint_as_float(g.y << 16), acc);
  acc = fmaf(__uint_as_float(a.y & 0xffff0000u), __uint_as_float(g.y & 0xffff0000u), acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    const int h = static_cast<int>(row % H);
    const long long bs = row / H;
    const int sidx = static_cast<int>(bs % S);
    const int b = static_cast<int>(bs / S);
    dsum[(static_cast<long long>(b) * H + h) * S + sidx] = acc;
  }
}

// ---------------------------------------------------------------------------- dQ
constexpr uint32_t GD_OFF_Q = 0, GD_OFF_DO = GQ_TILE_BYTES, GD_OFF_K = 2 * GQ_TILE_BYTES, GD_OFF_V = 4 * GQ_TILE_BYTES;
constexpr uint32_t GD_OFF_BARS = 6 * GQ_TILE_BYTES;
constexpr uint32_t GD_QDO_FULL = GD_OFF_BARS, GD_KV_FULL = GD_QDO_FULL + 8, GD_KV_EMPTY = GD_KV_FULL + 16, GD_SDP_FULL = GD_KV_EMPTY + 16,
                   GD_SDP_EMPTY = GD_SDP_FULL + 8, GD_DS_FULL = GD_SDP_EMPTY + 8, GD_DQ_DONE = GD_DS_FULL + 8, GD_TMEM_PTR = GD_DQ_DONE + 8;
constexpr int GD_SMEM = GD_TMEM_PTR + 16 + 1024;

__global__ void __launch_bounds__(GQ_THREADS, 1)
gqa_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                  const float* __restrict__ lse, const float* __restrict__ dsum, const int* __restrict__ kv_len,
                  __nv_bfloat16* __restrict__ dq, int S, int Hq, int Hkv, float scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = opaque_u32((smem_u32(smem_raw) + 1023u) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heavy tiles first: query tile qt visits qt + 1 kv tiles (causal), so the block order runs from the last tile down
  const int qt = gridDim.x - 1 - blockIdx.x, hq = blockIdx.y, b = blockIdx.z;
  const int hkv = hq / (Hq / Hkv);
  const int kvl = max(1, min(S, kv_len ? kv_len[b] : S));
  const int n_tiles = min(qt + 1, (kvl + GQ_T - 1) / GQ_T);
  const float scale_log2 = scale * LOG2E_GQ;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init_a(sb + GD_QDO_FULL, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init_a(sb + GD_KV_FULL + 8 * s, 1);
      mbar_init_a(sb + GD_KV_EMPTY + 8 * s, 1);
    }
    mbar_init_a(sb + GD_SDP_FULL, 1);
    mbar_init_a(sb + GD_SDP_EMPTY, 256);
    mbar_init_a(sb + GD_DS_FULL, 256);
    mbar_init_a(sb + GD_DQ_DONE, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + GD_TMEM_PTR), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sb + GD_TMEM_PTR) : "memory");
  // TMEM columns: S 0..127 | dP 128..255 | dQ 256..383 | dS 384..447 (bf16 pairs)

  if (warp == 0) {
    if (elect_one()) {                                     // ---------------- TMA producer
      mbar_arrive_expect_tx_a(sb + GD_QDO_FULL, 2 * GQ_TILE_BYTES);
      gq_issue_tile(sb + GD_OFF_Q, &tmQ, sb + GD_QDO_FULL, hq, qt * GQ_T, b);
      gq_issue_tile(sb + GD_OFF_DO, &tmDO, sb + GD_QDO_FULL, hq, qt * GQ_T, b);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        if (j >= 2) mbar_wait_a(sb + GD_KV_EMPTY + 8 * s, ((j >> 1) - 1) & 1);
        mbar_arrive_expect_tx_a(sb + GD_KV_FULL + 8 * s, 2 * GQ_TILE_BYTES);
        gq_issue_tile(sb + GD_OFF_K + s * GQ_TILE_BYTES, &tmK, sb + GD_KV_FULL + 8 * s, hkv, j * GQ_T, b);
        gq_issue_tile(sb + GD_OFF_V + s * GQ_TILE_BYTES, &tmV, sb + GD_KV_FULL + 8 * s, hkv, j * GQ_T, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    mbar_wait_a(sb + GD_QDO_FULL, 0);
    auto issue_sdp = [&](int j) {                          // S = Q K_j^T and dP = dO V_j^T
      const int s = j & 1;
      mbar_wait_a(sb + GD_KV_FULL + 8 * s, (j >> 1) & 1);
      if (j >= 1) mbar_wait_a(sb + GD_SDP_EMPTY, (j - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        gq_mma_kk(tmem_base, sb + GD_OFF_Q, sb + GD_OFF_K + s * GQ_TILE_BYTES, false);
        gq_mma_kk(tmem_base + 128, sb + GD_OFF_DO, sb + GD_OFF_V + s * GQ_TILE_BYTES, false);
        umma_commit_a(sb + GD_SDP_FULL);
      }
      __syncwarp();
    };
    issue_sdp(0);
    for (int j = 0; j < n_tiles; ++j) {
      if (j + 1 < n_tiles) issue_sdp(j + 1);
      mbar_wait_a(sb + GD_DS_FULL, j & 1);
      tc_fence_after();
      if (elect_one()) {                                   // dQ += dS K_j
        gq_mma_tm(tmem_base + 256, tmem_base + 384, sb + GD_OFF_K + (j & 1) * GQ_TILE_BYTES, j != 0);
        umma_commit_a(sb + GD_KV_EMPTY + 8 * (j & 1));
        umma_commit_a(sb + GD_DQ_DONE);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ dS: two threads per query row
    const int half = (warp - 4) >> 2;
    const uint32_t row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((row & ~31u) << 16);
    const int q_glob = qt * GQ_T + static_cast<int>(row);
    const bool live = q_glob < S;
    const size_t stat = (static_cast<size_t>(b) * Hq + hq) * S + (live ? q_glob : 0);
    const float neg_lse = live ? -lse[stat] : -INFINITY;   // rows past S: p = 0
    const float dsum_row = live ? dsum[stat] : 0.f;
    const unsigned long long c2 = pk2(scale_log2, scale_log2), nl2 = pk2(neg_lse, neg_lse);
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait_a(sb + GD_SDP_FULL, j & 1);
      tc_fence_after();
      uint32_t pk[32];
      const int col0 = j * GQ_T + half * 64;
      const bool edge = (j == qt) || (col0 + 64 > kvl);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], dp[32];
        tmem_ld_32x32(tlane + half * 64 + c * 32, sv);
        tmem_ld_32x32(tlane + 128 + half * 64 + c * 32, dp);
        tmem_ld_wait();
        if (c == 1) {
          tc_fence_before();
          mbar_arrive_a(sb + GD_SDP_EMPTY);                // S and dP of this tile are in registers
        }
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          const unsigned long long x2 = ffma2(pk2(__uint_as_float(sv[k]), __uint_as_float(sv[k + 1])), c2, nl2);
          float x0, x1;
          unpk2(x2, x0, x1);
          float p0 = ex2f(x0), p1 = ex2f(x1);
          if (edge) {
            const int cg = col0 + c * 32 + k;
            if (cg > q_glob || cg >= kvl) p0 = 0.f;
            if (cg + 1 > q_glob || cg + 1 >= kvl) p1 = 0.f;
          }
          const float d0 = p0 * (__uint_as_float(dp[k]) - dsum_row) * scale;
          const float d1 = p1 * (__uint_as_float(dp[k + 1]) - dsum_row) * scale;
          pk[c * 16 + (k >> 1)] = pack_bf16(d0, d1);
        }
      }
      if (j > 0) {                                         // the previous dQ product has read dS
        mbar_wait_a(sb + GD_DQ_DONE, (j - 1) & 1);
        tc_fence_after();
      }
      {
        uint32_t(&p0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&pk[0]);
        uint32_t(&p1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&pk[16]);
        tmem_st_32x16(tlane + 384 + half * 32, p0);
        tmem_st_32x16(tlane + 384 + half * 32 + 16, p1);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(sb + GD_DS_FULL);
    }
    mbar_wait_a(sb + GD_DQ_DONE, (n_tiles - 1) & 1);
    tc_fence_after();
    uint4* dst = reinterpret_cast<uint4*>(dq + ((static_cast<size_t>(b) * S + q_glob) * Hq + hq) * GQ_T + half * 64);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tlane + 256 + half * 64 + c * 32, r);
      tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(r[8 * u]), __uint_as_float(r[8 * u + 1]));
          v.y = pack_bf16(__uint_as_float(r[8 * u + 2]), __uint_as_float(r[8 * u + 3]));
          v.z = pack_bf16(__uint_as_float(r[8 * u + 4]), __uint_as_float(r[8 * u + 5]));
          v.w = pack_bf16(__uint_as_float(r[8 * u + 6]), __uint_as_float(r[8 * u + 7]));
          dst[c * 4 + u] = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------- dK, dV
constexpr uint32_t GK_OFF_K = 0, GK_OFF_V = GQ_TILE_BYTES, GK_OFF_Q = 2 * GQ_TILE_BYTES, GK_OFF_DO = 4 * GQ_TILE_BYTES;
constexpr uint32_t GK_OFF_STAT = 6 * GQ_TILE_BYTES;                 // [2 stages][lse | D][128] f32
constexpr uint32_t GK_OFF_BARS = GK_OFF_STAT + 2 * 2 * GQ_T * 4;
constexpr uint32_t GK_KV_FULL = GK_OFF_BARS, GK_QDO_FULL = GK_KV_FULL + 8, GK_QDO_EMPTY = GK_QDO_FULL + 16, GK_ST_FULL = GK_QDO_EMPTY + 16,
                   GK_PT_FULL = GK_ST_FULL + 8, GK_ACC_DONE = GK_PT_FULL + 8, GK_TMEM_PTR = GK_ACC_DONE + 8;
constexpr int GK_SMEM = GK_TMEM_PTR + 16 + 1024;

__global__ void __launch_bounds__(GQ_THREADS, 1)
gqa_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const float* __restrict__ lse, const float* __restrict__ dsum, const int* __restrict__ kv_len,
                   __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int S, int Hq, int Hkv, float scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = opaque_u32((smem_u32(smem_raw) + 1023u) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, hkv = blockIdx.y, b = blockIdx.z;
  const int G = Hq / Hkv;
  const int kvl = max(1, min(S, kv_len ? kv_len[b] : S));
  const int nq = (S + GQ_T - 1) / GQ_T;
  const int n_q_local = nq - kt;                           // query tiles i >= kt (causal)
  const int n_it = G * n_q_local;
  const float scale_log2 = scale * LOG2E_GQ;

  if (kt * GQ_T >= kvl) {                                  // every key of this tile is padding: zero gradients
    if (warp >= 4) {
      const int half = (warp - 4) >> 2;
      const int kv_glob = kt * GQ_T + (warp & 3) * 32 + lane;
      if (kv_glob < S) {
        const size_t o = ((static_cast<size_t>(b) * S + kv_glob) * Hkv + hkv) * GQ_T + half * 64;
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          reinterpret_cast<uint4*>(dk + o)[u] = z;
          reinterpret_cast<uint4*>(dv + o)[u] = z;
        }
      }
    }
    return;
  }

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init_a(sb + GK_KV_FULL, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init_a(sb + GK_QDO_FULL + 8 * s, 1);
      mbar_init_a(sb + GK_QDO_EMPTY + 8 * s, 1);
    }
    mbar_init_a(sb + GK_ST_FULL, 1);
    mbar_init_a(sb + GK_PT_FULL, 256);
    mbar_init_a(sb + GK_ACC_DONE, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + GK_TMEM_PTR), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sb + GK_TMEM_PTR) : "memory");
  // TMEM columns: S^T 0..127 (P^T bf16 over 0..63) | dP^T 128..255 (dS^T bf16 over 128..191) | dV 256..383 | dK 384..511

  if (warp == 0) {
    if (elect_one()) {                                     // ---------------- TMA producer
      mbar_arrive_expect_tx_a(sb + GK_KV_FULL, 2 * GQ_TILE_BYTES);
      gq_issue_tile(sb + GK_OFF_K, &tmK, sb + GK_KV_FULL, hkv, kt * GQ_T, b);
      gq_issue_tile(sb + GK_OFF_V, &tmV, sb + GK_KV_FULL, hkv, kt * GQ_T, b);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const int hq = hkv * G + it / n_q_local, i = kt + it % n_q_local;
        if (it >= 2) mbar_wait_a(sb + GK_QDO_EMPTY + 8 * s, ((it >> 1) - 1) & 1);
        mbar_arrive_expect_tx_a(sb + GK_QDO_FULL + 8 * s, 2 * GQ_TILE_BYTES);
        gq_issue_tile(sb + GK_OFF_Q + s * GQ_TILE_BYTES, &tmQ, sb + GK_QDO_FULL + 8 * s, hq, i * GQ_T, b);
        gq_issue_tile(sb + GK_OFF_DO + s * GQ_TILE_BYTES, &tmDO, sb + GK_QDO_FULL + 8 * s, hq, i * GQ_T, b);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    mbar_wait_a(sb + GK_KV_FULL, 0);
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      mbar_wait_a(sb + GK_QDO_FULL + 8 * s, (it >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {                                   // S^T = K Q_i^T, dP^T = V dO_i^T (after the previous dV / dK: in order)
        gq_mma_kk(tmem_base, sb + GK_OFF_K, sb + GK_OFF_Q + s * GQ_TILE_BYTES, false);
        gq_mma_kk(tmem_base + 128, sb + GK_OFF_V, sb + GK_OFF_DO + s * GQ_TILE_BYTES, false);
        umma_commit_a(sb + GK_ST_FULL);
      }
      __syncwarp();
      mbar_wait_a(sb + GK_PT_FULL, it & 1);
      tc_fence_after();
      if (elect_one()) {                                   // dV += P^T dO_i, dK += dS^T Q_i
        gq_mma_tm(tmem_base + 256, tmem_base, sb + GK_OFF_DO + s * GQ_TILE_BYTES, it != 0);
        gq_mma_tm(tmem_base + 384, tmem_base + 128, sb + GK_OFF_Q + s * GQ_TILE_BYTES, it != 0);
        umma_commit_a(sb + GK_QDO_EMPTY + 8 * s);
        if (it == n_it - 1) umma_commit_a(sb + GK_ACC_DONE);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ P^T, dS^T: lane = key row, columns = queries
    const int half = (warp - 4) >> 2;
    const uint32_t row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((row & ~31u) << 16);
    const int kv_glob = kt * GQ_T + static_cast<int>(row);
    const bool key_ok = kv_glob < kvl;
    const int tid_c = threadIdx.x - 128;
    const unsigned long long c2 = pk2(scale_log2, scale_log2);
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      const int hq = hkv * G + it / n_q_local, i = kt + it % n_q_local;
      {  // this query tile's lse / D rows -> shared memory (under the S^T / dP^T products)
        const int qi = i * GQ_T + (tid_c & 127);
        const size_t stat = (static_cast<size_t>(b) * Hq + hq) * S + min(qi, S - 1);
        float v = (tid_c < 128) ? lse[stat] : dsum[stat];
        if (qi >= S) v = (tid_c < 128) ? INFINITY : 0.f;   // rows past S: p = 2^(-inf) = 0
        sts_f32(sb + GK_OFF_STAT + s * (2 * GQ_T * 4) + tid_c * 4, v);
      }
      named_bar_sync(GQ_BAR_EXCH, 256);
      mbar_wait_a(sb + GK_ST_FULL, it & 1);
      tc_fence_after();
      const int q0 = i * GQ_T + half * 64;                 // first query column of this thread
      const bool edge = (i == kt);                         // diagonal tile: query < key is masked
      const uint32_t st_lse = sb + GK_OFF_STAT + s * (2 * GQ_T * 4) + half * 256, st_d = st_lse + GQ_T * 4;
      uint32_t pt[32], ds[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], dp[32];
        tmem_ld_32x32(tlane + half * 64 + c * 32, sv);
        tmem_ld_32x32(tlane + 128 + half * 64 + c * 32, dp);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          float l0, l1, e0, e1;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(l0), "=f"(l1) : "r"(st_lse + (c * 32 + k) * 4));
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(e0), "=f"(e1) : "r"(st_d + (c * 32 + k) * 4));
          const unsigned long long x2 = ffma2(pk2(__uint_as_float(sv[k]), __uint_as_float(sv[k + 1])), c2, pk2(-l0, -l1));
          float x0, x1;
          unpk2(x2, x0, x1);
          float p0 = key_ok ? ex2f(x0) : 0.f, p1 = key_ok ? ex2f(x1) : 0.f;
          if (edge) {
            const int qg = q0 + c * 32 + k;
            if (qg < kv_glob) p0 = 0.f;
            if (qg + 1 < kv_glob) p1 = 0.f;
          }
          pt[c * 16 + (k >> 1)] = pack_bf16(p0, p1);
          ds[c * 16 + (k >> 1)] = pack_bf16(p0 * (__uint_as_float(dp[k]) - e0) * scale, p1 * (__uint_as_float(dp[k + 1]) - e1) * scale);
        }
      }
      tc_fence_before();
      named_bar_sync(GQ_BAR_EXCH, 256);                    // every thread has read S^T / dP^T: P^T / dS^T may overwrite them
      tc_fence_after();
      {
        uint32_t(&a0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&pt[0]);
        uint32_t(&a1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&pt[16]);
        uint32_t(&b0)[16] = *reinterpret_cast<uint32_t(*)[16]>(&ds[0]);
        uint32_t(&b1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&ds[16]);
        tmem_st_32x16(tlane + half * 32, a0);
        tmem_st_32x16(tlane + half * 32 + 16, a1);
        tmem_st_32x16(tlane + 128 + half * 32, b0);
        tmem_st_32x16(tlane + 128 + half * 32 + 16, b1);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(sb + GK_PT_FULL);
    }
    mbar_wait_a(sb + GK_ACC_DONE, 0);
    tc_fence_after();
    const bool live = kv_glob < S;
    const size_t o = ((static_cast<size_t>(b) * S + kv_glob) * Hkv + hkv) * GQ_T + half * 64;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint4* dst = reinterpret_cast<uint4*>((which == 0 ? dv : dk) + o);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tlane + 256 + which * 128 + half * 64 + c * 32, r);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(r[8 * u]), __uint_as_float(r[8 * u + 1]));
            v.y = pack_bf16(__uint_as_float(r[8 * u + 2]), __uint_as_float(r[8 * u + 3]));
            v.z = pack_bf16(__uint_as_float(r[8 * u + 4]), __uint_as_float(r[8 * u + 5]));
            v.w = pack_bf16(__uint_as_float(r[8 * u + 6]), __uint_as_float(r[8 * u + 7]));
            dst[c * 4 + u] = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

int launch_gqa_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo, const void* out,
                   const void* d_out, const float* lse, float* dsum_ws, const int* kv_len, void* dq, void* dk, void* dv, int B,
                   int S, int Hq, int Hkv, float scale, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    AL_CHECK_CUDA(cudaFuncSetAttribute(gqa_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GD_SMEM));
    AL_CHECK_CUDA(cudaFuncSetAttribute(gqa_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GK_SMEM));
    attr_set = true;
  }
  const long long rows = static_cast<long long>(B) * S * Hq;
  gqa_rowdot_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(d_out), dsum_ws, B, S, Hq);
  AL_CHECK_CUDA(cudaGetLastError());
  const int nt = (S + GQ_T - 1) / GQ_T;
  gqa_bwd_dq_kernel<<<dim3(nt, Hq, B), GQ_THREADS, GD_SMEM, stream>>>(tq, tk, tv, tdo, lse, dsum_ws, kv_len,
                                                                       reinterpret_cast<__nv_bfloat16*>(dq), S, Hq, Hkv, scale);
  AL_CHECK_CUDA(cudaGetLastError());
  gqa_bwd_dkv_kernel<<<dim3(nt, Hkv, B), GQ_THREADS, GK_SMEM, stream>>>(tq, tk, tv, tdo, lse, dsum_ws, kv_len,
                                                                         reinterpret_cast<__nv_bfloat16*>(dk),
                                                                         reinterpret_cast<__nv_bfloat16*>(dv), S, Hq, Hkv, scale);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace al
