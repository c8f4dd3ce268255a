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
// Three persistent kernels (one CTA per SM, 512 TMEM columns and ~195 KB of shared memory each) that walk a list of work
// items, heaviest first, in a serpentine order over the SMs; plus a row-dot kernel for D = rowsum(dO * O):
//   gqa_fwd_kernel      item = (128-query tile, query head, sample). Q rows in TMEM, S = Q K_j^T as a TS product
//                       double-buffered in TMEM -> online softmax, four threads per query row (exact row maximum exchanged
//                       through shared memory, lazy rescale of O, 1/4 of the exponentials on the FMA pipe) -> P bf16 in
//                       TMEM -> O += P V_j (A from TMEM, V MN-major). Writes O (TMA store) and the row log-sum-exp.
//   gqa_bwd_dq_kernel   same items. Q and dO rows in TMEM; over kv tiles j <= i: S = Q K_j^T, dP = dO V_j^T (TS products),
//                       P = 2^(S c - lse), dS = P (dP - D) -> bf16 over the thread's own dP columns -> dQ += dS K_j
//                       (A from TMEM, K MN-major); dQ * 1/sqrt(d) leaves through a TMA store.
//   gqa_bwd_dkv_kernel  item = (kv tile, kv head, sample); over the group's query heads and query tiles i >= j, everything
//                       TRANSPOSED so that the products that contract over queries take their A operand from TMEM:
//                       S^T = K_j Q_i^T, dP^T = V_j dO_i^T, then dV += P^T dO_i and dK += dS^T Q_i (dO / Q as MN-major B
//                       operands); the S^T and dP^T chains are staggered on the tensor pipe.
// Scores are recomputed in the backward (nothing but O and lse is saved). Design notes and measurements: DESIGN.md 11.6 / 11.7.
#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace al {

constexpr int GQ_T = 128;                      // tile rows (queries or keys) and head_dim
constexpr int GQ_TILE_BYTES = GQ_T * GQ_T * 2; // 32 KB: one 128 x 128 bf16 tile = two [128][64] SW128 boxes
constexpr int GQ_BOX_BYTES = GQ_T * 64 * 2;    // 16 KB
#define GQ_SPLIT 4                             // compute threads per query row of the forward (16 softmax warps: with 8 the
                                               // exponent phase is latency-bound, tools/ncu_hot.py: 'wait' + 'branch' stalls)
constexpr int GQ_CW = GQ_T / GQ_SPLIT;         // columns of a 128-wide tile row that one compute thread owns
constexpr int GQ_COMPUTE = GQ_T * GQ_SPLIT;    // compute threads
constexpr int GQ_THREADS = 128 + GQ_COMPUTE;   // warps 0-3 control (TMA, MMA, TMEM allocator, idle), then the compute warps
constexpr float GQ_TAU = 8.0f;                 // lazy-rescale threshold of the forward, log2 units
#ifndef GQ_POLY_NUM
#define GQ_POLY_NUM 1
#define GQ_POLY_DEN 4
#endif
constexpr int GQ_BAR_EXCH = 1;                 // named barrier of the compute threads
constexpr float LOG2E_GQ = 1.4426950408889634f;

// -DGQ_CYCLES: cycle accounting by one compute warp per CTA (development builds only: tools/build_gqa_cycles.sh;
// tools/bench_gqa_attention.py CYCLES=1 reads it back). Slots per kernel (forward 0.., dQ 16.., dK/dV 32..):
// [0] kv-tile steps, [1] cycles inside the tile loops, [2] cycles between tile loops (epilogue + next item's set-up), [3] items.
#ifdef GQ_CYCLES
__device__ unsigned long long gq_cyc[64];
#define GQC_DECL(...) long long __VA_ARGS__
#define GQC_NOW() clock64()
#define GQC_ADD(slot, v) atomicAdd(&gq_cyc[slot], static_cast<unsigned long long>(v))
#else
#define GQC_DECL(...)
#define GQC_NOW() 0
#define GQC_ADD(slot, v)
#endif

__device__ __forceinline__ void tma_load_4d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d_a(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// Byte offset, inside a [128 rows][128 bf16] tile staged as two SW128 boxes, of the 16-byte chunk `chunk` (0..15) of row
// `row`: box = chunk / 8, the chunk's position inside the 128-byte row is XORed with row % 8.
__device__ __forceinline__ uint32_t gq_tile_chunk(uint32_t row, uint32_t chunk) {
  return (chunk >> 3) * GQ_BOX_BYTES + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
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
// The MMA issuer is ONE thread, and every instruction it spends on building operand descriptors delays the next product
// (measured: ~80 cycles per product with descriptors rebuilt from addresses, against 32 - 64 cycles of tensor time).
// So descriptors are kept as a 32-bit low word (start address >> 4 | LBO >> 4 << 16) computed once per tile, the per
// k-step advance is an add of a compile-time constant, and the high word (SBO 1024, version 1, SWIZZLE_128B) is constant.
constexpr uint32_t GQ_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t gq_desc_kmajor(uint32_t tile) { return ((tile & 0x3FFFFu) >> 4) | ((16u >> 4) << 16); }
__device__ __forceinline__ uint32_t gq_desc_mnmajor(uint32_t tile) {   // the two 64-column atoms are 16 KB apart (LBO)
  return ((tile & 0x3FFFFu) >> 4) | ((static_cast<uint32_t>(GQ_BOX_BYTES) >> 4) << 16);
}
__device__ __forceinline__ void umma_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(GQ_DESC_HI)
      : "memory");
}
__device__ __forceinline__ void umma_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(GQ_DESC_HI)
      : "memory");
}
// D[tmem 128 x N] (+)= A[smem tile, 128 rows, K-major over d] * B[rows 64 H .. of a smem tile, K-major over d]^T, N = 128 or 64
// (contraction over head_dim: 8 k-steps across the two 64-column SW128 boxes)
template <int N, int H>
__device__ __forceinline__ void gq_mma_kk(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, bool accumulate) {
  constexpr uint32_t IDESC = umma_idesc_bf16(GQ_T, N);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint32_t off = ((k >> 2) * GQ_BOX_BYTES + (k & 3) * 32) >> 4;
    umma_ss_lo(d_tmem, a_lo + off, b_lo + off + ((H * 8192) >> 4), IDESC, (k != 0) || accumulate);
  }
}
// D[tmem 128 x 128] (+)= A[tmem: 128 lanes x (16 STEPS) bf16] * B[contraction rows 64 H .. of a smem tile, 128 columns, MN-major]
// (16 contraction rows = 2048 B per step)
template <int STEPS, int H>
__device__ __forceinline__ void gq_mma_tm(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, bool accumulate) {
  constexpr uint32_t IDESC = umma_idesc_bf16(GQ_T, GQ_T, 0, 1);
#pragma unroll
  for (int k = 0; k < STEPS; ++k)
    umma_ts_lo(d_tmem, a_tmem + k * 8, b_lo + ((2048 * (4 * H + k)) >> 4), IDESC, (k != 0) || accumulate);
}

// D[tmem 128 x 128] = A[tmem: 128 lanes x 128 bf16 = 64 columns] * B[smem tile, K-major over d]^T
__device__ __forceinline__ void gq_mma_tk(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo) {
  constexpr uint32_t IDESC = umma_idesc_bf16(GQ_T, GQ_T);
#pragma unroll
  for (int k = 0; k < 8; ++k)
    umma_ts_lo(d_tmem, a_tmem + k * 8, b_lo + (((k >> 2) * GQ_BOX_BYTES + (k & 3) * 32) >> 4), IDESC, k != 0);
}

// ============================================================================ forward
constexpr uint32_t GF_OFF_K = 0;
constexpr uint32_t GF_OFF_V = GF_OFF_K + 2 * GQ_TILE_BYTES;
constexpr uint32_t GF_OFF_STG = GF_OFF_V + 2 * GQ_TILE_BYTES;        // one tile: the next item's Q rows on their way to TMEM (TMA in),
                                                                     // then the finished O tile (TMA out)
constexpr uint32_t GF_OFF_EXCH = GF_OFF_STG + GQ_TILE_BYTES;         // [2 parities][GQ_SPLIT][128] f32 row maxima, then [GQ_SPLIT][128] sums
constexpr uint32_t GF_OFF_BARS = GF_OFF_EXCH + 3 * GQ_SPLIT * GQ_T * 4;
constexpr uint32_t GF_STG_FULL = GF_OFF_BARS + 128, GF_STG_EMPTY = GF_STG_FULL + 8;
constexpr uint32_t GF_Q_FULL = GF_OFF_BARS, GF_K_FULL = GF_Q_FULL + 8, GF_K_EMPTY = GF_K_FULL + 16, GF_V_FULL = GF_K_EMPTY + 16,
                   GF_V_EMPTY = GF_V_FULL + 16, GF_S_FULL = GF_V_EMPTY + 16, GF_S_EMPTY = GF_S_FULL + 16, GF_P_FULL = GF_S_EMPTY + 16,
                   GF_O_FULL = GF_P_FULL + 8, GF_TMEM_PTR = GF_O_FULL + 8;
constexpr int GF_SMEM = GF_OFF_BARS + 256 + 1024;

// Work items = (query tile, query head, batch), heaviest (largest query tile: most kv tiles under the causal mask) first.
// The kernels are PERSISTENT: one CTA per SM walks the item list in a serpentine order (round r gives CTA c item
// r G + c, the next round r G + (G - 1 - c)), which on a descending list is within ~1 % of a dynamic longest-first
// schedule, and the TMEM allocation, barrier set-up, launch latency and pipeline fill are paid once per SM instead of
// once per item (measured before: ~10 000 of ~27 000 cycles per item went to those, with nothing to overlap them).
struct GqItem {
  int qt, hq, b;
};
__device__ __forceinline__ bool gq_item(int round, int nq, int Hq, int B, GqItem& it) {
  const int G = static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
  const int idx = round * G + ((round & 1) ? (G - 1 - c) : c);
  const int per = Hq * B;
  if (idx >= nq * per) return false;
  it.qt = nq - 1 - idx / per;
  const int rem = idx % per;
  it.b = rem / Hq;
  it.hq = rem % Hq;
  return true;
}
__device__ __forceinline__ int gq_rounds(int nq, int Hq, int B) {
  return (nq * Hq * B + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
}

__global__ void __launch_bounds__(GQ_THREADS, 1)
gqa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, float* __restrict__ lse,
               const int* __restrict__ kv_len, int B, int S, int Hq, int Hkv, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = opaque_u32((smem_u32(smem_raw) + 1023u) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (S + GQ_T - 1) / GQ_T;
  const int rounds = gq_rounds(nq, Hq, B);
  const int G_heads = Hq / Hkv;
  auto tiles_of = [&](const GqItem& it) {
    const int kvl = max(1, min(S, kv_len ? kv_len[it.b] : S));
    return min(it.qt + 1, (kvl + GQ_T - 1) / GQ_T);
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmO);
    mbar_init_a(sb + GF_STG_FULL, 1);
    mbar_init_a(sb + GF_STG_EMPTY, 1);
    mbar_init_a(sb + GF_Q_FULL, GQ_COMPUTE);
    for (int s = 0; s < 2; ++s) {
      mbar_init_a(sb + GF_K_FULL + 8 * s, 1);
      mbar_init_a(sb + GF_K_EMPTY + 8 * s, 1);
      mbar_init_a(sb + GF_V_FULL + 8 * s, 1);
      mbar_init_a(sb + GF_V_EMPTY + 8 * s, 1);
      mbar_init_a(sb + GF_S_FULL + 8 * s, 1);
      mbar_init_a(sb + GF_S_EMPTY + 8 * s, GQ_COMPUTE);
    }
    mbar_init_a(sb + GF_P_FULL, GQ_COMPUTE);
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
  // TMEM columns: S0 0..127 | S1 128..255 | O 256..383 | P 384..447 (bf16 pairs) | Q 448..511 (bf16 pairs: the A operand of
  // every S = Q K_j^T of an item, so those products read only K from shared memory)
  // Every ring / hand-off below is indexed by g, the number of (item, kv tile) steps this CTA has taken so far: stages
  // and mbarrier phases run on across item boundaries.

  if (warp == 0) {
    if (elect_one()) {                                     // ---------------- TMA producer (runs ahead into the next item)
      int g = 0, n_item = 0;
      for (int r = 0; r < rounds; ++r) {
        GqItem it;
        if (!gq_item(r, nq, Hq, B, it)) continue;
        if (n_item == 0) gq_load_tile(sb + GF_OFF_STG, &tmQ, sb + GF_STG_FULL, it.hq, it.qt * GQ_T, it.b);
        const int n_tiles = tiles_of(it), hkv = it.hq / G_heads;
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const int s = g & 1;
          if (g >= 2) mbar_wait_a(sb + GF_K_EMPTY + 8 * s, ((g >> 1) - 1) & 1);
          gq_load_tile(sb + GF_OFF_K + s * GQ_TILE_BYTES, &tmK, sb + GF_K_FULL + 8 * s, hkv, j * GQ_T, it.b);
          if (g >= 2) mbar_wait_a(sb + GF_V_EMPTY + 8 * s, ((g >> 1) - 1) & 1);
          gq_load_tile(sb + GF_OFF_V + s * GQ_TILE_BYTES, &tmV, sb + GF_V_FULL + 8 * s, hkv, j * GQ_T, it.b);
        }
        // the NEXT item's Q tile, as soon as the staging tile is free (this item's rows are in TMEM and the previous
        // item's O tile has left)
        GqItem nx;
        bool has_next = false;
        for (int r2 = r + 1; r2 < rounds && !has_next; ++r2) has_next = gq_item(r2, nq, Hq, B, nx);
        if (has_next) {
          mbar_wait_a(sb + GF_STG_EMPTY, n_item & 1);
          gq_load_tile(sb + GF_OFF_STG, &tmQ, sb + GF_STG_FULL, nx.hq, nx.qt * GQ_T, nx.b);
        }
        ++n_item;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: S_j one tile ahead of P V_{j-1}
    const uint32_t k_lo = gq_desc_kmajor(sb + GF_OFF_K), v_lo = gq_desc_mnmajor(sb + GF_OFF_V);
    auto issue_qk = [&](int g) {
      const int s = g & 1;
      mbar_wait_a(sb + GF_K_FULL + 8 * s, (g >> 1) & 1);
      if (g >= 2) mbar_wait_a(sb + GF_S_EMPTY + 8 * s, ((g >> 1) - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        gq_mma_tk(tmem_base + s * 128, tmem_base + 448, k_lo + s * (GQ_TILE_BYTES >> 4));
        umma_commit_a(sb + GF_K_EMPTY + 8 * s);
        umma_commit_a(sb + GF_S_FULL + 8 * s);
      }
      __syncwarp();
    };
    int g = 0, n_item = 0;
    for (int r = 0; r < rounds; ++r) {
      GqItem it;
      if (!gq_item(r, nq, Hq, B, it)) continue;
      const int n_tiles = tiles_of(it);
      mbar_wait_a(sb + GF_Q_FULL, n_item & 1);             // this item's Q rows are in TMEM
      ++n_item;
      issue_qk(g);
      for (int j = 0; j < n_tiles; ++j, ++g) {
        if (j + 1 < n_tiles) issue_qk(g + 1);
        const int s = g & 1;
        mbar_wait_a(sb + GF_V_FULL + 8 * s, (g >> 1) & 1);
        mbar_wait_a(sb + GF_P_FULL, g & 1);                // (for j = 0 this also says: the previous item's O has been read)
        tc_fence_after();
        if (elect_one()) {
          gq_mma_tm<8, 0>(tmem_base + 256, tmem_base + 384, v_lo + s * (GQ_TILE_BYTES >> 4), j != 0);
          umma_commit_a(sb + GF_V_EMPTY + 8 * s);
          umma_commit_a(sb + GF_O_FULL);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax: GQ_SPLIT threads per query row
    const int part = (warp - 4) >> 2;
    const uint32_t row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((row & ~31u) << 16);
    constexpr uint32_t EX_STRIDE = GQ_T * 4, EX_SET = GQ_SPLIT * GQ_T * 4;   // one part's slots | one parity's set
    const uint32_t ex_row = sb + GF_OFF_EXCH + row * 4;
    // An item's Q rows: staged by TMA as a swizzled tile; this thread moves its GQ_CW columns into TMEM (bf16 pairs are
    // already the A-operand layout of the TS product). Only after the last S of the previous item has been produced.
    int n_rows = 0;                                        // staged tiles consumed so far (STG_FULL phase)
    auto take_q = [&]() {
      mbar_wait_a(sb + GF_STG_FULL, n_rows & 1);
      ++n_rows;
#pragma unroll
      for (int c = 0; c < GQ_CW / 32; ++c) {
        uint32_t rr[16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t a = sb + GF_OFF_STG + gq_tile_chunk(row, part * (GQ_CW / 8) + 4 * c + u);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(rr[4 * u]), "=r"(rr[4 * u + 1]), "=r"(rr[4 * u + 2]), "=r"(rr[4 * u + 3]) : "r"(a));
        }
        tmem_st_32x16(tlane + 448 + part * (GQ_CW / 2) + 16 * c, rr);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(sb + GF_Q_FULL);
    };
    int g = 0;
    bool first = true;
#ifdef GQ_CYCLES
    long long tq_prev = 0;
#endif
    for (int r = 0; r < rounds; ++r) {
      GqItem it;
      if (!gq_item(r, nq, Hq, B, it)) continue;
      const int qt = it.qt, hq = it.hq, b = it.b;
      const int kvl = max(1, min(S, kv_len ? kv_len[b] : S));
      const int n_tiles = min(qt + 1, (kvl + GQ_T - 1) / GQ_T);
      const int q_glob = qt * GQ_T + static_cast<int>(row);
      if (first) {                                         // (later items: taken at the end of the previous one)
        take_q();
        named_bar_sync(GQ_BAR_EXCH, GQ_COMPUTE);           // every thread has read the staged tile
        if (warp == 4 && elect_one()) mbar_arrive_a(sb + GF_STG_EMPTY);
        first = false;
      }
      float m_ref = -INFINITY, l = 0.f;
#ifdef GQ_CYCLES
      const long long tq0 = GQC_NOW();
      if (lane == 0 && warp == 4 && tq_prev != 0) GQC_ADD(2, tq0 - tq_prev);
#endif
      for (int j = 0; j < n_tiles; ++j, ++g) {
        const int sbuf = g & 1;
        mbar_wait_a(sb + GF_S_FULL + 8 * sbuf, (g >> 1) & 1);
        tc_fence_after();
        uint32_t s[GQ_CW];
#pragma unroll
        for (int c = 0; c < GQ_CW / 32; ++c) {
          uint32_t(&sc)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32 * c]);
          tmem_ld_32x32(tlane + sbuf * 128 + part * GQ_CW + 32 * c, sc);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_a(sb + GF_S_EMPTY + 8 * sbuf);
        const int col0 = j * GQ_T + part * GQ_CW;
        if (j == qt || col0 + GQ_CW > kvl) {               // diagonal tile or the tile holding the padding boundary
#pragma unroll
          for (int k = 0; k < GQ_CW; ++k)
            if (col0 + k > q_glob || col0 + k >= kvl) s[k] = 0xff800000u;
        }
        float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
#pragma unroll
        for (int k = 2; k < GQ_CW; k += 4) {
          mx0 = fmax3(mx0, __uint_as_float(s[k]), __uint_as_float(s[k + 1]));
          if (k + 2 < GQ_CW) mx1 = fmax3(mx1, __uint_as_float(s[k + 2]), __uint_as_float(s[k + 3]));
        }
        sts_f32(ex_row + sbuf * EX_SET + part * EX_STRIDE, fmaxf(mx0, mx1));
        named_bar_sync(GQ_BAR_EXCH, GQ_COMPUTE);
        float tile_max = lds_f32(ex_row + sbuf * EX_SET);
#pragma unroll
        for (int qq = 1; qq < GQ_SPLIT; ++qq) tile_max = fmaxf(tile_max, lds_f32(ex_row + sbuf * EX_SET + qq * EX_STRIDE));
        tile_max *= scale_log2;
        float alpha = 1.0f;
        const bool rescale = __any_sync(0xffffffffu, tile_max > m_ref + GQ_TAU);
        if (rescale && tile_max > m_ref + GQ_TAU) {        // (per row; the TMEM traffic below stays warp-uniform)
          alpha = ex2f(m_ref - tile_max);                  // 0 on the first tile
          m_ref = tile_max;
          l *= alpha;
        }
        const unsigned long long c2 = pk2(scale_log2, scale_log2), nm2 = pk2(-m_ref, -m_ref);
        unsigned long long l2a = pk2(0.f, 0.f), l2b = pk2(0.f, 0.f);
        uint32_t pk[GQ_CW / 2];
#pragma unroll
        for (int k = 0; k < GQ_CW; k += 2) {
          const unsigned long long x2 = ffma2(pk2(__uint_as_float(s[k]), __uint_as_float(s[k + 1])), c2, nm2);
          float p0, p1;
          if (((k >> 1) % GQ_POLY_DEN) < GQ_POLY_NUM) {     // this share of the exponentials on the FMA pipe (the XU is the
            poly_exp2_pair<false>(x2, p0, p1);              // narrower one: 16 ex2 / cycle / SM against 16384 per tile)
          } else {
            float x0, x1;
            unpk2(x2, x0, x1);
            p0 = ex2f(x0);
            p1 = ex2f(x1);
          }
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
        if (j > 0) {                                       // the previous P V has read P and finished its part of O
          mbar_wait_a(sb + GF_O_FULL, (g - 1) & 1);
          tc_fence_after();
          if (rescale) {
#pragma unroll
            for (int c = 0; c < GQ_CW / 32; ++c) {
              uint32_t rr[32];
              tmem_ld_32x32(tlane + 256 + part * GQ_CW + c * 32, rr);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 32; ++k) rr[k] = __float_as_uint(__uint_as_float(rr[k]) * alpha);
              tmem_st_32x32(tlane + 256 + part * GQ_CW + c * 32, rr);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < GQ_CW / 32; ++c) {
          uint32_t(&pc)[16] = *reinterpret_cast<uint32_t(*)[16]>(&pk[16 * c]);
          tmem_st_32x16(tlane + 384 + part * (GQ_CW / 2) + 16 * c, pc);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_a(sb + GF_P_FULL);
      }
#ifdef GQ_CYCLES
      tq_prev = GQC_NOW();
      if (lane == 0 && warp == 4) {
        GQC_ADD(0, n_tiles);
        GQC_ADD(1, tq_prev - tq0);
        GQC_ADD(3, 1);
      }
#endif
      // row sum of all parts; then the next item's Q rows into TMEM (every S of this item has been produced, so no product
      // reads the Q columns any more): the issuer starts the next item's first S under this item's epilogue
      GqItem nx;
      bool has_next = false;
      for (int r2 = r + 1; r2 < rounds && !has_next; ++r2) has_next = gq_item(r2, nq, Hq, B, nx);
      sts_f32(ex_row + 2 * EX_SET + part * EX_STRIDE, l);
      named_bar_sync(GQ_BAR_EXCH, GQ_COMPUTE);
      float l_tot = lds_f32(ex_row + 2 * EX_SET);
#pragma unroll
      for (int qq = 1; qq < GQ_SPLIT; ++qq) l_tot += lds_f32(ex_row + 2 * EX_SET + qq * EX_STRIDE);
      const float inv_l = 1.0f / l_tot;
      if (has_next) take_q();
      mbar_wait_a(sb + GF_O_FULL, (g - 1) & 1);
      tc_fence_after();
      if (q_glob < S) {
        if (part == 0) lse[(static_cast<size_t>(b) * Hq + hq) * S + q_glob] = m_ref + log2f(l_tot);
      }
      // epilogue: normalised O tile -> staging tile (swizzled) -> one TMA store (rows past S are clipped by the tensor map)
      uint32_t ov[GQ_CW / 2];
#pragma unroll
      for (int c = 0; c < GQ_CW / 32; ++c) {
        uint32_t rr[32];
        tmem_ld_32x32(tlane + 256 + part * GQ_CW + c * 32, rr);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k)
          ov[16 * c + k] = pack_bf16(__uint_as_float(rr[2 * k]) * inv_l, __uint_as_float(rr[2 * k + 1]) * inv_l);
      }
      tc_fence_before();
      named_bar_sync(GQ_BAR_EXCH, GQ_COMPUTE);             // every thread has read the staged Q tile of the next item
      // (and has read its O columns: the next item's first P V is issued only after all of them have delivered its P)
#pragma unroll
      for (int u = 0; u < GQ_CW / 8; ++u)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(sb + GF_OFF_STG + gq_tile_chunk(row, part * (GQ_CW / 8) + u)), "r"(ov[4 * u]), "r"(ov[4 * u + 1]),
                       "r"(ov[4 * u + 2]), "r"(ov[4 * u + 3])
                     : "memory");
      fence_proxy_async_smem();
      named_bar_sync(GQ_BAR_EXCH, GQ_COMPUTE);
      if (warp == 4 && elect_one()) {
        tma_store_4d_a(&tmO, sb + GF_OFF_STG, 0, hq, qt * GQ_T, b);
        tma_store_4d_a(&tmO, sb + GF_OFF_STG + GQ_BOX_BYTES, 64, hq, qt * GQ_T, b);
        tma_commit_group();
        if (has_next) {
          tma_wait_group_read<0>();                        // the tile has left shared memory: the staging tile is free
          mbar_arrive_a(sb + GF_STG_EMPTY);
        } else {
          tma_wait_group<0>();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

static int gq_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

int launch_gqa_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, float* lse,
                   const int* kv_len, int B, int S, int Hq, int Hkv, float scale, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    AL_CHECK_CUDA(cudaFuncSetAttribute(gqa_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GF_SMEM));
    attr_set = true;
  }
  const int items = ((S + GQ_T - 1) / GQ_T) * Hq * B;
  gqa_fwd_kernel<<<std::min(items, gq_num_sms()), GQ_THREADS, GF_SMEM, stream>>>(tq, tk, tv, to, lse, kv_len, B, S, Hq, Hkv,
                                                                                 scale * 1.4426950408889634f);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}


// ============================================================================ backward
__device__ __forceinline__ void gq_issue_tile(uint32_t dst, const CUtensorMap* m, uint32_t bar, int h, int s0, int b) {
  tma_load_4d_a(dst, m, bar, 0, h, s0, b);
  tma_load_4d_a(dst + GQ_BOX_BYTES, m, bar, 64, h, s0, b);
}

// D[b][h][s] = sum_d dO[b][s][h][d] * O[b][s][h][d]: 16 lanes per 256-byte row (one 16-byte load of each tensor per lane),
// four rows per thread in flight, grid-stride over groups of 64 rows.
__global__ void __launch_bounds__(256)
gqa_rowdot_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, float* __restrict__ dsum,
                  int B, int S, int H) {
  const long long rows = static_cast<long long>(B) * S * H;
  const int sub = threadIdx.x & 15, rin = threadIdx.x >> 4;   // lane within the row, row within a 16-row slab
  for (long long base = static_cast<long long>(blockIdx.x) * 64; base < rows; base += static_cast<long long>(gridDim.x) * 64) {
    uint4 a[4], g[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long row = base + u * 16 + rin;
      const bool ok = row < rows;
      a[u] = ok ? __ldg(reinterpret_cast<const uint4*>(o + row * GQ_T) + sub) : make_uint4(0, 0, 0, 0);
      g[u] = ok ? __ldg(reinterpret_cast<const uint4*>(d_o + row * GQ_T) + sub) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t aw[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, gw[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(gw[e] << 16), acc);
        acc = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(gw[e] & 0xffff0000u), acc);
      }
#pragma unroll
      for (int off = 8; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      const long long row = base + u * 16 + rin;
      if (sub == 0 && row < rows) {
        const int h = static_cast<int>(row % H);
        const long long bs = row / H;
        const int sidx = static_cast<int>(bs % S);
        const int b = static_cast<int>(bs / S);
        dsum[(static_cast<long long>(b) * H + h) * S + sidx] = acc;
      }
    }
  }
}

// ---------------------------------------------------------------------------- dQ
// Q_i and dO_i are the A operands of every product of this kernel's loop (S = Q K_j^T, dP = dO V_j^T), so they live
// in TMEM for the whole CTA (written once from global memory by the compute threads) and the products read only their
// B operand from shared memory: shared-memory bandwidth, not the tensor pipe, is what bounds the SS form here
// (tools/att_lab/mma_rate.cu: 8 KB of operands per 64-cycle product = the full 128 B / cycle of the SM).
constexpr int GD_SPLIT = 4;                    // compute threads per query row (16 warps: the exponent phase is latency-bound
                                               // with fewer - tools/ncu_hot.py on the 8-warp form: 'wait' + 'branch' stalls)
constexpr int GD_CW = GQ_T / GD_SPLIT;         // 32 columns per thread
constexpr int GD_COMPUTE = GQ_T * GD_SPLIT;
constexpr int GD_THREADS = 128 + GD_COMPUTE;   // warps 0-3 control, then the compute warps
constexpr uint32_t GD_OFF_K = 0, GD_OFF_V = 2 * GQ_TILE_BYTES;
constexpr uint32_t GD_OFF_STG = 4 * GQ_TILE_BYTES;                   // two tiles: the next item's Q | dO rows on their way to TMEM
                                                                     // (TMA in), then the finished dQ tile in the first (TMA out)
constexpr uint32_t GD_OFF_BARS = 6 * GQ_TILE_BYTES;
constexpr uint32_t GD_STG_FULL = GD_OFF_BARS + 128, GD_STG_EMPTY = GD_STG_FULL + 8;
constexpr uint32_t GD_QDO_FULL = GD_OFF_BARS, GD_KV_FULL = GD_QDO_FULL + 8, GD_KV_EMPTY = GD_KV_FULL + 16, GD_S_FULL = GD_KV_EMPTY + 16,
                   GD_S_EMPTY = GD_S_FULL + 8, GD_DP_FULL = GD_S_EMPTY + 8, GD_DS_FULL = GD_DP_FULL + 8, GD_DQ_DONE = GD_DS_FULL + 8,
                   GD_TMEM_PTR = GD_DQ_DONE + 8;
constexpr int GD_SMEM = GD_OFF_BARS + 256 + 1024;
// TMEM columns
static_assert(GD_CW == 32, "the dQ kernel's compute loop handles one 32-column chunk per thread");
constexpr uint32_t GD_TM_S = 0, GD_TM_DP = 128, GD_TM_DQ = 256, GD_TM_Q = 384, GD_TM_DO = 448;

__global__ void __launch_bounds__(GD_THREADS, 1)
gqa_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                  const __grid_constant__ CUtensorMap tmDQ, const float* __restrict__ lse, const float* __restrict__ dsum,
                  const int* __restrict__ kv_len, int B, int S, int Hq, int Hkv, float scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = opaque_u32((smem_u32(smem_raw) + 1023u) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (S + GQ_T - 1) / GQ_T;
  const int rounds = gq_rounds(nq, Hq, B);               // persistent CTAs over the heavy-first item list (see the forward)
  const int G_heads = Hq / Hkv;
  const float scale_log2 = scale * LOG2E_GQ;
  auto tiles_of = [&](const GqItem& it) {
    const int kvl = max(1, min(S, kv_len ? kv_len[it.b] : S));
    return min(it.qt + 1, (kvl + GQ_T - 1) / GQ_T);
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    mbar_init_a(sb + GD_STG_FULL, 1);
    mbar_init_a(sb + GD_STG_EMPTY, 1);
    mbar_init_a(sb + GD_QDO_FULL, GD_COMPUTE);
    for (int s = 0; s < 2; ++s) {
      mbar_init_a(sb + GD_KV_FULL + 8 * s, 1);
      mbar_init_a(sb + GD_KV_EMPTY + 8 * s, 1);
    }
    mbar_init_a(sb + GD_S_FULL, 1);
    mbar_init_a(sb + GD_S_EMPTY, GD_COMPUTE);
    mbar_init_a(sb + GD_DP_FULL, 1);
    mbar_init_a(sb + GD_DS_FULL, GD_COMPUTE);
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
  // TMEM columns: S 0..127 | dP 128..255 (each thread's dS, bf16 pairs, over the first half of its own dP columns)
  //               | dQ 256..383 | Q 384..447 | dO 448..511 (bf16 pairs, A operands)
  // g = (item, kv tile) steps taken by this CTA so far: ring stages and mbarrier phases run on across items.

  if (warp == 0) {
    if (elect_one()) {                                     // ---------------- TMA producer
      auto load_rows = [&](const GqItem& it) {             // an item's Q and dO tiles into the staging pair
        mbar_arrive_expect_tx_a(sb + GD_STG_FULL, 2 * GQ_TILE_BYTES);
        gq_issue_tile(sb + GD_OFF_STG, &tmQ, sb + GD_STG_FULL, it.hq, it.qt * GQ_T, it.b);
        gq_issue_tile(sb + GD_OFF_STG + GQ_TILE_BYTES, &tmDO, sb + GD_STG_FULL, it.hq, it.qt * GQ_T, it.b);
      };
      int g = 0, n_item = 0;
      for (int r = 0; r < rounds; ++r) {
        GqItem it;
        if (!gq_item(r, nq, Hq, B, it)) continue;
        if (n_item == 0) load_rows(it);
        const int n_tiles = tiles_of(it), hkv = it.hq / G_heads;
        for (int j = 0; j < n_tiles; ++j, ++g) {
          const int s = g & 1;
          if (g >= 2) mbar_wait_a(sb + GD_KV_EMPTY + 8 * s, ((g >> 1) - 1) & 1);
          mbar_arrive_expect_tx_a(sb + GD_KV_FULL + 8 * s, 2 * GQ_TILE_BYTES);
          gq_issue_tile(sb + GD_OFF_K + s * GQ_TILE_BYTES, &tmK, sb + GD_KV_FULL + 8 * s, hkv, j * GQ_T, it.b);
          gq_issue_tile(sb + GD_OFF_V + s * GQ_TILE_BYTES, &tmV, sb + GD_KV_FULL + 8 * s, hkv, j * GQ_T, it.b);
        }
        // the NEXT item's rows, as soon as the staging pair is free (this item's rows are in TMEM and the previous
        // item's dQ tile has left): about two kv tiles ahead of the moment the compute threads ask for them
        GqItem nx;
        bool has_next = false;
        for (int r2 = r + 1; r2 < rounds && !has_next; ++r2) has_next = gq_item(r2, nq, Hq, B, nx);
        if (has_next) {
          mbar_wait_a(sb + GD_STG_EMPTY, n_item & 1);
          load_rows(nx);
        }
        ++n_item;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // Order: S_{g+1} as soon as S_g is in registers; dQ_g when dS_g is written; dP_{g+1} behind dQ_g (dS_g lives in
    // the dP columns; tensor-core operations of one thread execute in issue order).
    const uint32_t k_lo = gq_desc_kmajor(sb + GD_OFF_K), v_lo = gq_desc_kmajor(sb + GD_OFF_V), kmn_lo = gq_desc_mnmajor(sb + GD_OFF_K);
    constexpr uint32_t STAGE = GQ_TILE_BYTES >> 4;
    auto issue_s = [&](int g) {                            // needs K_g in shared memory and S_{g-1} in registers
      const int s = g & 1;
      mbar_wait_a(sb + GD_KV_FULL + 8 * s, (g >> 1) & 1);
      if (g >= 1) mbar_wait_a(sb + GD_S_EMPTY, (g - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        gq_mma_tk(tmem_base + GD_TM_S, tmem_base + GD_TM_Q, k_lo + s * STAGE);
        umma_commit_a(sb + GD_S_FULL);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int g) {                           // (behind dQ_{g-1} in issue order)
      if (elect_one()) {
        gq_mma_tk(tmem_base + GD_TM_DP, tmem_base + GD_TM_DO, v_lo + (g & 1) * STAGE);
        umma_commit_a(sb + GD_DP_FULL);
      }
      __syncwarp();
    };
    int g = 0, n_item = 0;
    for (int r = 0; r < rounds; ++r) {
      GqItem it;
      if (!gq_item(r, nq, Hq, B, it)) continue;
      const int n_tiles = tiles_of(it);
      mbar_wait_a(sb + GD_QDO_FULL, n_item & 1);           // this item's Q and dO rows are in TMEM
      ++n_item;
      issue_s(g);
      issue_dp(g);
      for (int j = 0; j < n_tiles; ++j, ++g) {
        if (j + 1 < n_tiles) issue_s(g + 1);
        mbar_wait_a(sb + GD_DS_FULL, g & 1);
        tc_fence_after();
        if (elect_one()) {                                 // dQ += dS K_g: k-step k = keys 16 k .., from the part that owns them
          constexpr uint32_t IDESC = umma_idesc_bf16(GQ_T, GQ_T, 0, 1);
          const uint32_t b_lo = kmn_lo + (g & 1) * STAGE;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_ts_lo(tmem_base + GD_TM_DQ, tmem_base + GD_TM_DP + (16 * k / GD_CW) * GD_CW + ((16 * k % GD_CW) >> 1),
                       b_lo + ((2048 * k) >> 4), IDESC, (k != 0) || (j != 0));
          umma_commit_a(sb + GD_KV_EMPTY + 8 * (g & 1));
          if (j + 1 == n_tiles) umma_commit_a(sb + GD_DQ_DONE);
        }
        __syncwarp();
        if (j + 1 < n_tiles) issue_dp(g + 1);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ dS: GD_SPLIT threads per query row
    const int part = (warp - 4) >> 2;
    const uint32_t row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((row & ~31u) << 16);
    // An item's Q and dO rows: staged by TMA as two swizzled tiles; this thread moves its GD_CW columns of both rows into
    // TMEM (bf16 pairs are already the A-operand layout) and fetches the row's -lse and -D.
    int n_rows = 0;                                        // staged row sets consumed so far (STG_FULL phase)
    auto take_rows = [&](const GqItem& it, float& neg_lse, float& neg_d) {
      const int qg = it.qt * GQ_T + static_cast<int>(row);
      const bool live = qg < S;
      const size_t stat = (static_cast<size_t>(it.b) * Hq + it.hq) * S + (live ? qg : 0);
      neg_lse = live ? -lse[stat] : -INFINITY;             // rows past S: p = 0 (their Q / dO rows arrive as zeros)
      neg_d = live ? -dsum[stat] : 0.f;
      mbar_wait_a(sb + GD_STG_FULL, n_rows & 1);
      ++n_rows;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        uint32_t rr[GD_CW / 2];
#pragma unroll
        for (int u = 0; u < GD_CW / 8; ++u) {
          const uint32_t a = sb + GD_OFF_STG + which * GQ_TILE_BYTES + gq_tile_chunk(row, part * (GD_CW / 8) + u);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(rr[4 * u]), "=r"(rr[4 * u + 1]), "=r"(rr[4 * u + 2]), "=r"(rr[4 * u + 3]) : "r"(a));
        }
        tmem_st_32x16(tlane + (which == 0 ? GD_TM_Q : GD_TM_DO) + part * (GD_CW / 2), rr);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive_a(sb + GD_QDO_FULL);
    };
    int g = 0, n_item = 0;
    float neg_lse = 0.f, neg_d = 0.f;
    bool first = true;
#ifdef GQ_CYCLES
    long long tq_prev = 0;
#endif
    for (int r = 0; r < rounds; ++r) {
      GqItem it;
      if (!gq_item(r, nq, Hq, B, it)) continue;
      const int qt = it.qt, hq = it.hq, b = it.b;
      const int kvl = max(1, min(S, kv_len ? kv_len[b] : S));
      const int n_tiles = min(qt + 1, (kvl + GQ_T - 1) / GQ_T);
      const int q_glob = qt * GQ_T + static_cast<int>(row);
      if (first) {                                         // (later items: taken at the end of the previous one)
        take_rows(it, neg_lse, neg_d);
        named_bar_sync(GQ_BAR_EXCH, GD_COMPUTE);           // every thread has read the staged rows
        if (warp == 4 && elect_one()) mbar_arrive_a(sb + GD_STG_EMPTY);
        first = false;
      }
      const unsigned long long c2 = pk2(scale_log2, scale_log2), nl2 = pk2(neg_lse, neg_lse), nd2 = pk2(neg_d, neg_d);
#ifdef GQ_CYCLES
      const long long tq0 = GQC_NOW();
      if (lane == 0 && warp == 4 && tq_prev != 0) GQC_ADD(18, tq0 - tq_prev);
#endif
      for (int j = 0; j < n_tiles; ++j, ++g) {
        mbar_wait_a(sb + GD_S_FULL, g & 1);
        tc_fence_after();
        const int col0 = j * GQ_T + part * GD_CW;
        float pf[GD_CW];                                   // this thread's probabilities
        {
          uint32_t sv[32];
          tmem_ld_32x32(tlane + GD_TM_S + part * GD_CW, sv);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive_a(sb + GD_S_EMPTY);                  // S of this tile is in registers
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const unsigned long long x2 = ffma2(pk2(__uint_as_float(sv[k]), __uint_as_float(sv[k + 1])), c2, nl2);
            if (((k >> 1) % GQ_POLY_DEN) < GQ_POLY_NUM) {
              poly_exp2_pair<true>(x2, pf[k], pf[k + 1]);  // (x <= 0 up to rounding for live entries; masked ones can be large)
            } else {
              float x0, x1;
              unpk2(x2, x0, x1);
              pf[k] = ex2f(x0);
              pf[k + 1] = ex2f(x1);
            }
          }
          if ((j == qt) || (col0 + GD_CW > kvl)) {         // diagonal tile or the tile holding the padding boundary
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (col0 + k > q_glob || col0 + k >= kvl) pf[k] = 0.f;
          }
        }
        mbar_wait_a(sb + GD_DP_FULL, g & 1);
        tc_fence_after();
        {
          uint32_t dp[32], ds[16];
          tmem_ld_32x32(tlane + GD_TM_DP + part * GD_CW, dp);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const unsigned long long t2 = fadd2(pk2(__uint_as_float(dp[k]), __uint_as_float(dp[k + 1])), nd2);
            const unsigned long long d2 = fmul2(pk2(pf[k], pf[k + 1]), t2);
            float d0, d1;
            unpk2(d2, d0, d1);
            ds[k >> 1] = pack_bf16(d0, d1);                // (the 1/sqrt(d) factor is applied to dQ at the end)
          }
          // dS over the first half of this thread's own dP columns (columns it has read; nobody else touches them)
          tmem_st_32x16(tlane + GD_TM_DP + part * GD_CW, ds);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_a(sb + GD_DS_FULL);
      }
#ifdef GQ_CYCLES
      tq_prev = GQC_NOW();
      if (lane == 0 && warp == 4) {
        GQC_ADD(16, n_tiles);
        GQC_ADD(17, tq_prev - tq0);
        GQC_ADD(19, 1);
      }
#endif
      // Next item's rows: every S and dP of this item has been produced (this thread has seen the last DP_FULL), so the Q
      // and dO columns of TMEM are free; the issuer starts the next item's first products under this item's epilogue.
      GqItem nx;
      bool has_next = false;
      for (int r2 = r + 1; r2 < rounds && !has_next; ++r2) has_next = gq_item(r2, nq, Hq, B, nx);
      if (has_next) take_rows(nx, neg_lse, neg_d);
      mbar_wait_a(sb + GD_DQ_DONE, n_item & 1);
      ++n_item;
      tc_fence_after();
      // epilogue: dQ tile -> first staging tile (swizzled) -> one TMA store (rows past S are clipped by the tensor map)
      {
        uint32_t rr[32];
        tmem_ld_32x32(tlane + GD_TM_DQ + part * GD_CW, rr);
        tmem_ld_wait();
        named_bar_sync(GQ_BAR_EXCH, GD_COMPUTE);           // every thread has read the staged rows of the next item
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t v0 = pack_bf16(__uint_as_float(rr[8 * u]) * scale, __uint_as_float(rr[8 * u + 1]) * scale);
          const uint32_t v1 = pack_bf16(__uint_as_float(rr[8 * u + 2]) * scale, __uint_as_float(rr[8 * u + 3]) * scale);
          const uint32_t v2 = pack_bf16(__uint_as_float(rr[8 * u + 4]) * scale, __uint_as_float(rr[8 * u + 5]) * scale);
          const uint32_t v3 = pack_bf16(__uint_as_float(rr[8 * u + 6]) * scale, __uint_as_float(rr[8 * u + 7]) * scale);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                       ::"r"(sb + GD_OFF_STG + gq_tile_chunk(row, part * (GD_CW / 8) + u)), "r"(v0), "r"(v1), "r"(v2), "r"(v3)
                       : "memory");
        }
        fence_proxy_async_smem();
        named_bar_sync(GQ_BAR_EXCH, GD_COMPUTE);
        if (warp == 4 && elect_one()) {
          tma_store_4d_a(&tmDQ, sb + GD_OFF_STG, 0, hq, qt * GQ_T, b);
          tma_store_4d_a(&tmDQ, sb + GD_OFF_STG + GQ_BOX_BYTES, 64, hq, qt * GQ_T, b);
          tma_commit_group();
          if (has_next) {
            tma_wait_group_read<0>();                      // the tile has left shared memory: the staging pair is free
            mbar_arrive_a(sb + GD_STG_EMPTY);
          } else {
            tma_wait_group<0>();
          }
        }
      }
      tc_fence_before();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------- dK, dV
// Work item = (kv tile, kv head, sample): K_j and V_j stay in shared memory while the group's query heads and the query
// tiles i >= j stream through a two-stage Q / dO ring. Everything is TRANSPOSED (lane = key, column = query) so that the
// products that contract over queries take their A operand from TMEM. All 512 TMEM columns are in use (S^T, dP^T, dV, dK),
// so nothing is double-buffered; instead the two dependency chains of an iteration are staggered on the tensor pipe:
//     S^T_g -> (P^T) -> dV_g -> S^T_{g+1}          dP^T_g -> (dS^T) -> dK_g -> dP^T_{g+1}
// dK_{g-1} and dP^T_g run under the exponent phase of iteration g, dV_g and S^T_{g+1} under its dS phase (tensor-core
// operations of one thread execute in issue order, so a product that overwrites a region is simply issued behind the one
// that reads it). P^T / dS^T are written by each thread over the first half of its OWN S^T / dP^T columns. Persistent CTAs
// over the item list, heaviest kv tile (tile 0 meets every query tile) first.
constexpr int GK_SPLIT = 4;                    // compute threads per key row
constexpr int GK_CW = GQ_T / GK_SPLIT;         // 32 query columns per thread
constexpr int GK_COMPUTE = GQ_T * GK_SPLIT;
constexpr int GK_THREADS = 128 + GK_COMPUTE;
static_assert(GK_CW == 32, "the dK / dV kernel's compute loop handles one 32-column chunk per thread");
constexpr uint32_t GK_OFF_K = 0, GK_OFF_V = GQ_TILE_BYTES, GK_OFF_Q = 2 * GQ_TILE_BYTES, GK_OFF_DO = 4 * GQ_TILE_BYTES;
constexpr uint32_t GK_OFF_STAT = 6 * GQ_TILE_BYTES;                 // [2 stages][-lse 128 | -D 128] f32
constexpr uint32_t GK_OFF_BARS = GK_OFF_STAT + 2 * 2 * GQ_T * 4;
constexpr uint32_t GK_KV_FULL = GK_OFF_BARS, GK_KV_EMPTY = GK_KV_FULL + 8, GK_QDO_FULL = GK_KV_EMPTY + 8, GK_QDO_EMPTY = GK_QDO_FULL + 16,
                   GK_ST_FULL = GK_QDO_EMPTY + 16, GK_DPT_FULL = GK_ST_FULL + 8, GK_PT_FULL = GK_DPT_FULL + 8, GK_DST_FULL = GK_PT_FULL + 8,
                   GK_ACC_DONE = GK_DST_FULL + 8, GK_TMEM_PTR = GK_ACC_DONE + 8;
constexpr int GK_SMEM = GK_OFF_BARS + 256 + 1024;

struct GkItem {
  int kt, hkv, b;
};
__device__ __forceinline__ bool gk_item(int round, int nq, int Hkv, int B, GkItem& it) {
  const int G = static_cast<int>(gridDim.x), c = static_cast<int>(blockIdx.x);
  const int idx = round * G + ((round & 1) ? (G - 1 - c) : c);
  const int per = Hkv * B;
  if (idx >= nq * per) return false;
  it.kt = idx / per;                                       // kv tile 0 first: it meets every query tile
  const int rem = idx % per;
  it.b = rem / Hkv;
  it.hkv = rem % Hkv;
  return true;
}

__global__ void __launch_bounds__(GK_THREADS, 1)
gqa_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const float* __restrict__ lse, const float* __restrict__ dsum, const int* __restrict__ kv_len,
                   __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int B, int S, int Hq, int Hkv, float scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = opaque_u32((smem_u32(smem_raw) + 1023u) & ~1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = Hq / Hkv;
  const int nq = (S + GQ_T - 1) / GQ_T;
  const int rounds = (nq * Hkv * B + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const float scale_log2 = scale * LOG2E_GQ;
  // iterations of an item: the group's G query heads x the query tiles i >= kt; 0 when every key of the tile is padding
  auto iters_of = [&](const GkItem& it) {
    const int kvl = max(1, min(S, kv_len ? kv_len[it.b] : S));
    return (it.kt * GQ_T >= kvl) ? 0 : G * (nq - it.kt);
  };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmDO);
    mbar_init_a(sb + GK_KV_FULL, 1);
    mbar_init_a(sb + GK_KV_EMPTY, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init_a(sb + GK_QDO_FULL + 8 * s, 1);
      mbar_init_a(sb + GK_QDO_EMPTY + 8 * s, 1);
    }
    mbar_init_a(sb + GK_ST_FULL, 1);
    mbar_init_a(sb + GK_DPT_FULL, 1);
    mbar_init_a(sb + GK_PT_FULL, GK_COMPUTE);
    mbar_init_a(sb + GK_DST_FULL, GK_COMPUTE);
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
  // TMEM columns: S^T 0..127 | dP^T 128..255 | dV 256..383 | dK 384..511; thread (key row, part p) owns columns 32 p .. 32 p + 31
  // of S^T and dP^T and writes its 32 P^T / dS^T values (bf16 pairs) over the first 16 of them.
  // g = iterations taken by this CTA so far, nl = items with work so far: stages and mbarrier phases run on across items.

  if (warp == 0) {
    if (elect_one()) {                                     // ---------------- TMA producer
      int g = 0, nl = 0;
      for (int r = 0; r < rounds; ++r) {
        GkItem it;
        if (!gk_item(r, nq, Hkv, B, it)) continue;
        const int n_it = iters_of(it);
        if (n_it == 0) continue;
        const int n_q_local = nq - it.kt;
        if (nl >= 1) mbar_wait_a(sb + GK_KV_EMPTY, (nl - 1) & 1);   // the previous item's last S^T / dP^T have read K / V
        mbar_arrive_expect_tx_a(sb + GK_KV_FULL, 2 * GQ_TILE_BYTES);
        gq_issue_tile(sb + GK_OFF_K, &tmK, sb + GK_KV_FULL, it.hkv, it.kt * GQ_T, it.b);
        gq_issue_tile(sb + GK_OFF_V, &tmV, sb + GK_KV_FULL, it.hkv, it.kt * GQ_T, it.b);
        for (int t = 0; t < n_it; ++t, ++g) {
          const int s = g & 1;
          const int hq = it.hkv * G + t / n_q_local, i = it.kt + t % n_q_local;
          if (g >= 2) mbar_wait_a(sb + GK_QDO_EMPTY + 8 * s, ((g >> 1) - 1) & 1);
          mbar_arrive_expect_tx_a(sb + GK_QDO_FULL + 8 * s, 2 * GQ_TILE_BYTES);
          gq_issue_tile(sb + GK_OFF_Q + s * GQ_TILE_BYTES, &tmQ, sb + GK_QDO_FULL + 8 * s, hq, i * GQ_T, it.b);
          gq_issue_tile(sb + GK_OFF_DO + s * GQ_TILE_BYTES, &tmDO, sb + GK_QDO_FULL + 8 * s, hq, i * GQ_T, it.b);
        }
        ++nl;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t k_lo = gq_desc_kmajor(sb + GK_OFF_K), v_lo = gq_desc_kmajor(sb + GK_OFF_V), q_lo = gq_desc_kmajor(sb + GK_OFF_Q),
                   do_lo = gq_desc_kmajor(sb + GK_OFF_DO), qmn_lo = gq_desc_mnmajor(sb + GK_OFF_Q),
                   domn_lo = gq_desc_mnmajor(sb + GK_OFF_DO);
    constexpr uint32_t STAGE = GQ_TILE_BYTES >> 4;
    auto issue_s = [&](int g) {                            // S^T = K Q^T (behind dV_{g-1}, which reads P^T_{g-1} from the same columns)
      const int s = g & 1;
      mbar_wait_a(sb + GK_QDO_FULL + 8 * s, (g >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        gq_mma_kk<128, 0>(tmem_base, k_lo, q_lo + s * STAGE, false);
        umma_commit_a(sb + GK_ST_FULL);
      }
      __syncwarp();
    };
    auto issue_dp = [&](int g, bool last_of_item) {        // dP^T = V dO^T (behind dK_{g-1})
      const int s = g & 1;
      if (elect_one()) {
        gq_mma_kk<128, 0>(tmem_base + 128, v_lo, do_lo + s * STAGE, false);
        umma_commit_a(sb + GK_DPT_FULL);
        if (last_of_item) umma_commit_a(sb + GK_KV_EMPTY); // no later product of this item reads K / V from shared memory
      }
      __syncwarp();
    };
    // D[tmem 128 x 128] (+)= A[tmem: this iteration's P^T or dS^T, 16 queries = 8 columns per k-step, in the owner's columns]
    //                        * B[smem tile, 128 query rows = contraction, MN-major]
    auto issue_acc = [&](uint32_t d_tmem, uint32_t a_base, uint32_t b_lo, bool accumulate) {
      constexpr uint32_t IDESC = umma_idesc_bf16(GQ_T, GQ_T, 0, 1);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_ts_lo(d_tmem, a_base + (16 * k / GK_CW) * GK_CW + ((16 * k % GK_CW) >> 1), b_lo + ((2048 * k) >> 4), IDESC,
                   (k != 0) || accumulate);
    };
    int g = 0, nl = 0;
    for (int r = 0; r < rounds; ++r) {
      GkItem it;
      if (!gk_item(r, nq, Hkv, B, it)) continue;
      const int n_it = iters_of(it);
      if (n_it == 0) continue;
      mbar_wait_a(sb + GK_KV_FULL, nl & 1);
      issue_s(g);
      issue_dp(g, n_it == 1);
      for (int t = 0; t < n_it; ++t, ++g) {
        const int s = g & 1;
        const bool more = t + 1 < n_it;
        mbar_wait_a(sb + GK_PT_FULL, g & 1);               // (t = 0: also says the previous item's accumulators have been read)
        tc_fence_after();
        if (elect_one()) issue_acc(tmem_base + 256, tmem_base, domn_lo + s * STAGE, t != 0);          // dV += P^T dO
        __syncwarp();
        if (more) issue_s(g + 1);
        mbar_wait_a(sb + GK_DST_FULL, g & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_acc(tmem_base + 384, tmem_base + 128, qmn_lo + s * STAGE, t != 0);                     // dK += dS^T Q
          umma_commit_a(sb + GK_QDO_EMPTY + 8 * s);
          if (!more) umma_commit_a(sb + GK_ACC_DONE);
        }
        __syncwarp();
        if (more) issue_dp(g + 1, t + 2 == n_it);
      }
      ++nl;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ P^T, dS^T: lane = key row, columns = queries
    const int part = (warp - 4) >> 2;
    const uint32_t row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tmem_base + ((row & ~31u) << 16);
    const int tid_c = threadIdx.x - 128;
    const unsigned long long c2 = pk2(scale_log2, scale_log2);
    int g = 0, nl = 0;
#ifdef GQ_CYCLES
    long long tq_prev = 0;
#endif
    for (int r = 0; r < rounds; ++r) {
      GkItem item;
      if (!gk_item(r, nq, Hkv, B, item)) continue;
      const int kt = item.kt, hkv = item.hkv, b = item.b;
      const int kvl = max(1, min(S, kv_len ? kv_len[b] : S));
      const int kv_glob = kt * GQ_T + static_cast<int>(row);
      const size_t o = ((static_cast<size_t>(b) * S + kv_glob) * Hkv + hkv) * GQ_T + part * GK_CW;
      const int n_it = iters_of(item);
      if (n_it == 0) {                                     // every key of this tile is padding: zero gradients
        if (kv_glob < S) {
          const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int u = 0; u < GK_CW / 8; ++u) {
            reinterpret_cast<uint4*>(dk + o)[u] = z;
            reinterpret_cast<uint4*>(dv + o)[u] = z;
          }
        }
        continue;
      }
      const int n_q_local = nq - kt;
      const bool key_ok = kv_glob < kvl;
      auto load_stat = [&](int t) -> float {               // -lse (threads 0..127) / -D (128..255) of iteration t's queries
        const int hq = hkv * G + t / n_q_local, i = kt + t % n_q_local;
        const int qi = i * GQ_T + (tid_c & 127);
        const size_t stat = (static_cast<size_t>(b) * Hq + hq) * S + min(qi, S - 1);
        float v = (tid_c < 128) ? -lse[stat] : -dsum[stat];
        if (qi >= S) v = (tid_c < 128) ? -INFINITY : 0.f;  // rows past S: p = 2^(-inf) = 0
        return v;
      };
      float stat_next = (tid_c < 256) ? load_stat(0) : 0.f;
#ifdef GQ_CYCLES
      const long long tq0 = GQC_NOW();
      if (lane == 0 && warp == 4 && tq_prev != 0) GQC_ADD(34, tq0 - tq_prev);
#endif
      for (int t = 0; t < n_it; ++t, ++g) {
        const int s = g & 1;
        const int i = kt + t % n_q_local;
        if (tid_c < 256) {
          sts_f32(sb + GK_OFF_STAT + s * (2 * GQ_T * 4) + tid_c * 4, stat_next);
          if (t + 1 < n_it) stat_next = load_stat(t + 1);
        }
        named_bar_sync(GQ_BAR_EXCH, GK_COMPUTE);
        const int q0 = i * GQ_T + part * GK_CW;            // first query column of this thread
        const uint32_t st_nl = sb + GK_OFF_STAT + s * (2 * GQ_T * 4) + part * (GK_CW * 4), st_nd = st_nl + GQ_T * 4;
        float pf[GK_CW];                                   // this thread's probabilities
        mbar_wait_a(sb + GK_ST_FULL, g & 1);
        tc_fence_after();
        {
          uint32_t sv[32], pt[16];
          tmem_ld_32x32(tlane + part * GK_CW, sv);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            unsigned long long nl_a, nl_b;
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(nl_a), "=l"(nl_b) : "r"(st_nl + k * 4));
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int kk = k + 2 * u;
              const unsigned long long x2 = ffma2(pk2(__uint_as_float(sv[kk]), __uint_as_float(sv[kk + 1])), c2, u ? nl_b : nl_a);
              if (((kk >> 1) % GQ_POLY_DEN) < GQ_POLY_NUM) {
                poly_exp2_pair<true>(x2, pf[kk], pf[kk + 1]);   // (masked entries can be large: clamped, then zeroed below)
              } else {
                float x0, x1;
                unpk2(x2, x0, x1);
                pf[kk] = ex2f(x0);
                pf[kk + 1] = ex2f(x1);
              }
            }
          }
          if (i == kt || !key_ok) {                        // diagonal tile: query < key is masked; padded keys: everything
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (q0 + k < kv_glob || !key_ok) pf[k] = 0.f;
          }
#pragma unroll
          for (int k = 0; k < 32; k += 2) pt[k >> 1] = pack_bf16(pf[k], pf[k + 1]);
          tmem_st_32x16(tlane + part * GK_CW, pt);         // over columns this thread has read; nobody else touches them
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_a(sb + GK_PT_FULL);
        mbar_wait_a(sb + GK_DPT_FULL, g & 1);
        tc_fence_after();
        {
          uint32_t dp[32], ds[16];
          tmem_ld_32x32(tlane + 128 + part * GK_CW, dp);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            unsigned long long nd_a, nd_b;
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(nd_a), "=l"(nd_b) : "r"(st_nd + k * 4));
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int kk = k + 2 * u;
              const unsigned long long t2 = fadd2(pk2(__uint_as_float(dp[kk]), __uint_as_float(dp[kk + 1])), u ? nd_b : nd_a);
              const unsigned long long d2 = fmul2(pk2(pf[kk], pf[kk + 1]), t2);
              float d0, d1;
              unpk2(d2, d0, d1);
              ds[kk >> 1] = pack_bf16(d0, d1);             // (the 1/sqrt(d) factor is applied to dK at the end)
            }
          }
          tmem_st_32x16(tlane + 128 + part * GK_CW, ds);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive_a(sb + GK_DST_FULL);
      }
#ifdef GQ_CYCLES
      tq_prev = GQC_NOW();
      if (lane == 0 && warp == 4) {
        GQC_ADD(32, n_it);
        GQC_ADD(33, tq_prev - tq0);
        GQC_ADD(35, 1);
      }
#endif
      mbar_wait_a(sb + GK_ACC_DONE, nl & 1);
      ++nl;
      tc_fence_after();
      const bool live = kv_glob < S;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        uint4* dst = reinterpret_cast<uint4*>((which == 0 ? dv : dk) + o);
        const float f = which == 0 ? 1.0f : scale;
        uint32_t rr[32];
        tmem_ld_32x32(tlane + 256 + which * 128 + part * GK_CW, rr);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(rr[8 * u]) * f, __uint_as_float(rr[8 * u + 1]) * f);
            v.y = pack_bf16(__uint_as_float(rr[8 * u + 2]) * f, __uint_as_float(rr[8 * u + 3]) * f);
            v.z = pack_bf16(__uint_as_float(rr[8 * u + 4]) * f, __uint_as_float(rr[8 * u + 5]) * f);
            v.w = pack_bf16(__uint_as_float(rr[8 * u + 6]) * f, __uint_as_float(rr[8 * u + 7]) * f);
            dst[u] = v;
          }
        }
      }
      tc_fence_before();                                   // (the next item's first dV is issued only after every thread's P^T)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

int launch_gqa_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo, const CUtensorMap& tdq,
                   const void* out,
                   const void* d_out, const float* lse, float* dsum_ws, const int* kv_len, void* dq, void* dk, void* dv, int B,
                   int S, int Hq, int Hkv, float scale, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    AL_CHECK_CUDA(cudaFuncSetAttribute(gqa_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GD_SMEM));
    AL_CHECK_CUDA(cudaFuncSetAttribute(gqa_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GK_SMEM));
    attr_set = true;
  }
  const long long rows = static_cast<long long>(B) * S * Hq;
  gqa_rowdot_kernel<<<static_cast<unsigned>(std::min<long long>((rows + 63) / 64, 8LL * gq_num_sms())), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(d_out), dsum_ws, B, S, Hq);
  AL_CHECK_CUDA(cudaGetLastError());
  const int nt = (S + GQ_T - 1) / GQ_T;
  gqa_bwd_dq_kernel<<<std::min(nt * Hq * B, gq_num_sms()), GD_THREADS, GD_SMEM, stream>>>(tq, tk, tv, tdo, tdq, lse, dsum_ws, kv_len,
                                                                                          B, S, Hq, Hkv, scale);
  AL_CHECK_CUDA(cudaGetLastError());
  gqa_bwd_dkv_kernel<<<std::min(nt * Hkv * B, gq_num_sms()), GK_THREADS, GK_SMEM, stream>>>(
      tq, tk, tv, tdo, lse, dsum_ws, kv_len, reinterpret_cast<__nv_bfloat16*>(dk), reinterpret_cast<__nv_bfloat16*>(dv), B, S, Hq, Hkv,
      scale);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace al

#ifdef GQ_CYCLES
extern "C" int al_debug_gqa_cycles(unsigned long long* out64, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out64, al::gq_cyc, sizeof(al::gq_cyc));
  if (reset) {
    unsigned long long z[64] = {0};
    cudaMemcpyToSymbol(al::gq_cyc, z, sizeof(z));
  }
  return 0;
}
#endif
