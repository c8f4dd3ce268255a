// Non-causal, unmasked multi-head attention forward for head_dim 64 on sm_100a (tcgen05 / TMEM / TMA).
//
// Replaces HF WhisperAttention's softmax(Q K^T) V (modeling_whisper.py:215-238 via :339-349; scaling = 1.0
// because q_proj's output is pre-scaled, :310) on the fused qkv buffer the QKV GEMM writes:
//   qkv [B][T][3*H*64] bf16  (q | k | v, head h at columns h*64..h*64+63 of each third)
//   out [B][T][H*64]   bf16
//
// One CTA = one query tile of 128 rows of one (batch, head); TWO CTAs are resident per SM (256 TMEM columns and
// ~82 KB of shared memory each), so one CTA's start-up, barrier waits and MMAs run under the other's exps. 8 warps:
//   warp 0   TMA producer: Q tile once, then K tiles through a 3-stage ring and V tiles through a 2-stage ring
//            (a K stage is free as soon as Q K_j^T has run, long before V_j is consumed)
//   warp 1   MMA issuer:   S = Q K_j^T (SS, fp32 in TMEM), O += P V_j (A = P from TMEM, B = V MN-major smem)
//   warp 2   TMEM allocator (256 columns: S | O | P)
//   warps 4-7  softmax warpgroup, one thread per query row.
// Softmax design (the kernel is exp-bound at head_dim 64: 128x128 exps per 2x 256-cycle MMAs):
//   * a tile's 128 scores are pulled from TMEM into registers in one go and the S buffer is released at once, so
//     Q K_{j+1}^T runs under the softmax of tile j;
//   * O accumulates in TMEM across kv tiles. The running reference max is only advanced when a row's tile max
//     exceeds it by more than 2^8 (P stays well inside bf16 / fp32 range), and only then is O rescaled in TMEM;
//     the final O / l is exact whatever reference was used;
//   * scale-and-subtract, row sums and the polynomial run as packed fp32x2 (FFMA2 / FADD2), the max as 3-input
//     FMNMX3; POLY_PAIRS of every 4 element pairs take exp2 on the FMA pipe (Cody-Waite split + degree-3
//     minimax, rel. error 7.7e-5, far below P's bf16 rounding) to relieve the MUFU.
#include "common.cuh"
#include "kernels.h"

namespace al {

constexpr int ATT_BQ = 128;       // query rows per tile
constexpr int ATT_BKV = 128;      // kv rows per tile
constexpr int ATT_HD = 64;
constexpr int ATT_K_STAGES = 3;
constexpr int ATT_V_STAGES = 2;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;   // 16 KB: any of Q / K / V tile
constexpr int ATT_THREADS = 384;                 // warps 0-3 control (TMA, MMA, TMEM allocator, idle), warps 4-11 softmax
constexpr int ATT_CTRL_REGS = 32;
constexpr int ATT_SOFTMAX_REGS = 104;
constexpr int ATT_SMEM = (1 + ATT_K_STAGES + ATT_V_STAGES) * ATT_TILE_BYTES + 5 * 2 * 128 * 4 + 256 + 1024;
constexpr float LOG2E = 1.4426950408889634f;
constexpr int ATT_TAU_LOG2 = 40;          // the reference follows a tile whose probabilities summed to more than 2^40
constexpr float ATT_RAW_LIMIT = 40.0f;    // QLOG2 fast form only while every first score of the warp's rows is within 2^+-40
constexpr float ATT_RISK_SUM = 1.2676506e30f;   // 2^100: a tile sum this large sends the CTA's rows to the exact path
#ifndef ATT_POLY_NUM
#define ATT_POLY_NUM 3
#define ATT_POLY_DEN 8
#endif
constexpr int POLY_NUM = ATT_POLY_NUM, POLY_DEN = ATT_POLY_DEN;   // of every DEN element pairs, NUM take exp2 on the FMA pipe
// softmax warps -> MMA warp hand-offs are named barriers (256 arrive + 32 sync); id 0 is __syncthreads
constexpr int ATT_BAR_P_FULL = 1;
constexpr int ATT_BAR_S_EMPTY = 2;
constexpr int ATT_BAR_SOFTMAX = 3;        // the 256 softmax threads among themselves
constexpr int ATT_BAR_COUNT = 256 + 32;

#ifdef ATT_TRACE
// Development-only event timeline (tools/att_lab): lane 0 of each role of a few CTAs records (tag, kv tile, %clock).
constexpr int TR_SLOTS = 32, TR_EVENTS = 256;
__device__ uint32_t g_att_trace[TR_SLOTS][4][TR_EVENTS][2];
__device__ uint32_t g_att_trace_n[TR_SLOTS][4];
__device__ uint32_t g_att_trace_sm[TR_SLOTS];
struct Tracer {
  int slot, role, n;
  __device__ Tracer(int role_, int lane) : role(role_), n(0) {
    const int lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    slot = lin < 16 ? lin : (lin >= 148 && lin < 164 ? lin - 132 : -1);
    if (lane != 0) slot = -1;
  }
  __device__ __forceinline__ void ev(int tag, int j) {
    if (slot >= 0 && n < TR_EVENTS) {
      uint32_t c;
      asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
      g_att_trace[slot][role][n][0] = (tag << 16) | j;
      g_att_trace[slot][role][n][1] = c;
      ++n;
    }
  }
  __device__ void done() {
    if (slot >= 0) {
      g_att_trace_n[slot][role] = n;
      uint32_t sm;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
      g_att_trace_sm[slot] = sm;
    }
  }
};
#define TR_DECL(role) Tracer tr(role, lane)
#define TR(tag, j) tr.ev(tag, j)
#define TR_DONE() tr.done()
#else
#define TR_DECL(role)
#define TR(tag, j)
#define TR_DONE()
#endif

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Exact attention for ONE query row by one warp (plain loads, online softmax in fp32). Only used for the rows of a CTA in
// which the fast path flagged a possible overflow of its lagged softmax reference (see the softmax warps below): a
// score that outgrows the reference by more than 2^100 within two kv tiles. Slow, exact, practically never taken.
__device__ __noinline__ void attention_row_exact(const __nv_bfloat16* __restrict__ qkv_b, __nv_bfloat16* __restrict__ out_row,
                                                 int q, int T, int d, int h, int lane, float to_log2) {
  const size_t ld = static_cast<size_t>(3) * d;
  const uint4* qp = reinterpret_cast<const uint4*>(qkv_b + static_cast<size_t>(q) * ld + h * ATT_HD);
  uint4 qv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = __ldg(qp + i);
  float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int base = 0; base < T; base += 32) {
    const int kv = base + lane;
    float sc = -INFINITY;
    if (kv < T) {
      const uint4* kp = reinterpret_cast<const uint4*>(qkv_b + static_cast<size_t>(kv) * ld + d + h * ATT_HD);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 kk = __ldg(kp + i);
        const uint32_t qa[4] = {qv[i].x, qv[i].y, qv[i].z, qv[i].w};
        const uint32_t ka[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc = fmaf(__uint_as_float(qa[e] << 16), __uint_as_float(ka[e] << 16), acc);
          acc = fmaf(__uint_as_float(qa[e] & 0xffff0000u), __uint_as_float(ka[e] & 0xffff0000u), acc);
        }
      }
      sc = acc * to_log2;
    }
    const float m_new = fmaxf(m, warp_max(sc));
    const float scale = exp2f(m - m_new);          // 0 on the first chunk
    m = m_new;
    const float p = exp2f(sc - m);                 // 0 for kv >= T
    l = l * scale + p;
    o0 *= scale;
    o1 *= scale;
    const int n = min(32, T - base);
    for (int i = 0; i < n; ++i) {
      const float pi = __shfl_sync(0xffffffffu, p, i);
      const uint32_t vv = __ldg(reinterpret_cast<const uint32_t*>(qkv_b + static_cast<size_t>(base + i) * ld + 2 * d + h * ATT_HD) + lane);
      o0 = fmaf(pi, __uint_as_float(vv << 16), o0);
      o1 = fmaf(pi, __uint_as_float(vv & 0xffff0000u), o1);
    }
  }
  l = warp_sum(l);
  const float inv = 1.0f / l;
  reinterpret_cast<uint32_t*>(out_row)[lane] = pack_bf16(o0 * inv, o1 * inv);
}

#ifdef ATT_CYCLES
__device__ unsigned long long g_att_cycles[2];   // [0] sum over CTAs of (last - first clock), [1] CTAs
#endif

// Shared memory, as byte offsets from the 1024-aligned base (every address in the kernel is base + constant)
constexpr uint32_t ATT_OFF_Q = 0;
constexpr uint32_t ATT_OFF_K = ATT_TILE_BYTES;
constexpr uint32_t ATT_OFF_V = ATT_OFF_K + ATT_K_STAGES * ATT_TILE_BYTES;
constexpr uint32_t ATT_OFF_PART = ATT_OFF_V + ATT_V_STAGES * ATT_TILE_BYTES;   // [4 tiles + final][2 halves][128 rows] f32 sums
constexpr uint32_t ATT_OFF_BARS = ATT_OFF_PART + 5 * 2 * ATT_BQ * 4;
constexpr uint32_t ATT_BAR_Q_FULL = ATT_OFF_BARS;
constexpr uint32_t ATT_BAR_K_FULL = ATT_BAR_Q_FULL + 8;
constexpr uint32_t ATT_BAR_K_EMPTY = ATT_BAR_K_FULL + 8 * ATT_K_STAGES;
constexpr uint32_t ATT_BAR_V_FULL = ATT_BAR_K_EMPTY + 8 * ATT_K_STAGES;
constexpr uint32_t ATT_BAR_V_EMPTY = ATT_BAR_V_FULL + 8 * ATT_V_STAGES;
constexpr uint32_t ATT_BAR_S_FULL = ATT_BAR_V_EMPTY + 8 * ATT_V_STAGES;         // "S drained" / "P stored" are named barriers
constexpr uint32_t ATT_BAR_O_FULL = ATT_BAR_S_FULL + 8;
constexpr uint32_t ATT_OFF_TMEM_PTR = ATT_BAR_O_FULL + 8;
constexpr uint32_t ATT_OFF_EXACT = ATT_OFF_TMEM_PTR + 4;                        // "a row of this tile needs the exact path"
static_assert(ATT_OFF_EXACT + 4 + 1024 <= ATT_SMEM, "shared-memory carve-up exceeds ATT_SMEM");

// QLOG2: the caller folded log2(e) into q as well (the scores arrive in log2 units). Then, while a warp's rows need no
// reference (first scores within 2^+-ATT_RAW_LIMIT and no reference move yet -- the normal case), exp2 is taken
// straight from the accumulator values: the scale-and-subtract FFMA2 of every element pair disappears.
template <bool QLOG2>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOut,
                     const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int H) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = opaque_u32((smem_u32(smem_raw) + 1023u) & ~1023u);   // the one base register

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef ATT_CYCLES
  const long long cyc0 = clock64();
#endif
  const int nkv = (T + ATT_BKV - 1) / ATT_BKV;

  // elect.sync, not threadIdx.x == 0: ptxas then knows ONE thread runs the branch and emits no per-lane "waterfall"
  // loop (ELECT / R2UR.BROADCAST / BRA.U.ANY) around every uniform-datapath instruction (UTMALDG, UTCHMMA, UTCBAR) --
  // with lane == 0 each MMA cost ~110 cycles to issue and the issue itself was the critical path of the kernel.
  if (warp == 0 && elect_one()) {
    const int d = H * ATT_HD, h = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * ATT_BQ;
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmOut);
    for (uint32_t a = ATT_BAR_Q_FULL; a <= ATT_BAR_O_FULL; a += 8) mbar_init_a(sb + a, 1);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sb + ATT_OFF_EXACT), "r"(0u) : "memory");
    fence_barrier_init();
    // the first loads need nothing but their barriers: they start before the TMEM allocation and the CTA-wide sync
    mbar_arrive_expect_tx_a(sb + ATT_BAR_Q_FULL, ATT_TILE_BYTES);
    tma_load_3d_a(sb + ATT_OFF_Q, &tmQKV, sb + ATT_BAR_Q_FULL, h * ATT_HD, q0, b);
    mbar_arrive_expect_tx_a(sb + ATT_BAR_K_FULL, ATT_TILE_BYTES);
    tma_load_3d_a(sb + ATT_OFF_K, &tmQKV, sb + ATT_BAR_K_FULL, d + h * ATT_HD, 0, b);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb + ATT_OFF_TMEM_PTR), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sb + ATT_OFF_TMEM_PTR) : "memory");
  // TMEM columns: S 0..127 | O 128..191 | P 192..255 (bf16 pairs: 128 kv -> 64 columns)

  // Register budget: launched with 80 regs x 384 threads (2 CTAs / SM). The control warpgroup drops to 32, which frees
  // 48 x 128 = 6144 registers; the 8 softmax warps grow to 104, which takes 24 x 256 = 6144. (Asking for more than was
  // freed makes setmaxnreg.inc wait forever.)
  if (warp < 4) {
    setmaxnreg_dec<ATT_CTRL_REGS>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      TR_DECL(0);
      if (elect_one()) {
        const int d = H * ATT_HD, h = blockIdx.y, b = blockIdx.z;
        // (Q and K_0 were requested by this same thread before the CTA-wide sync)
        // K runs ahead of V: K_{j+1} is requested before V_j so that Q K_{j+1}^T is never starved
        int ks = 1, vs = 0;
        uint32_t kph = 0, vph = 0;
#pragma unroll 1
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) {
            mbar_wait_a(sb + ATT_BAR_K_EMPTY + 8 * ks, kph ^ 1);
            TR(1, j + 1);
            mbar_arrive_expect_tx_a(sb + ATT_BAR_K_FULL + 8 * ks, ATT_TILE_BYTES);
            tma_load_3d_a(sb + ATT_OFF_K + ks * ATT_TILE_BYTES, &tmQKV, sb + ATT_BAR_K_FULL + 8 * ks, d + h * ATT_HD,
                          (j + 1) * ATT_BKV, b);
            if (++ks == ATT_K_STAGES) { ks = 0; kph ^= 1; }
          }
          mbar_wait_a(sb + ATT_BAR_V_EMPTY + 8 * vs, vph ^ 1);
          TR(2, j);
          mbar_arrive_expect_tx_a(sb + ATT_BAR_V_FULL + 8 * vs, ATT_TILE_BYTES);
          tma_load_3d_a(sb + ATT_OFF_V + vs * ATT_TILE_BYTES, &tmQKV, sb + ATT_BAR_V_FULL + 8 * vs, 2 * d + h * ATT_HD,
                        j * ATT_BKV, b);
          if (++vs == ATT_V_STAGES) { vs = 0; vph ^= 1; }
        }
      }
      TR_DONE();
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer 1: S = Q K_j^T
      // Two issuing warps (this one and warp 3 for O += P V_j): their loops only meet in the tensor pipe's queue, so a
      // P V product is issued the moment its P is stored, not after the issue (and queueing) of the next Q K^T.
      constexpr uint32_t IDESC_S = umma_idesc_bf16(ATT_BQ, ATT_BKV);             // Q K^T: both K-major
      TR_DECL(1);
      TR(10, 0);
      mbar_wait_a(sb + ATT_BAR_Q_FULL, 0);
      int ks = 0;
      uint32_t kph = 0;
#pragma unroll 1
      for (int j = 0; j < nkv; ++j) {
        mbar_wait_a(sb + ATT_BAR_K_FULL + 8 * ks, kph);
        TR(13, j);
        if (j > 0) named_bar_sync(ATT_BAR_S_EMPTY, ATT_BAR_COUNT);   // S drained; blocks in hardware, no polling
        tc_fence_after();
        TR(11, j);
        if (elect_one()) {
          const uint32_t tS = opaque_u32(tmem_base);                 // (opaque: nothing hoisted out of the loop and spilled)
          const uint64_t qd = umma_desc_sw128(sb + ATT_OFF_Q, 16, 1024);
          const uint64_t kd = umma_desc_sw128(sb + ATT_OFF_K + ks * ATT_TILE_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS, qd + 2 * k, kd + 2 * k, IDESC_S, k != 0);
          umma_commit_a(sb + ATT_BAR_K_EMPTY + 8 * ks);   // the K stage is free once these MMAs have run
          umma_commit_a(sb + ATT_BAR_S_FULL);
        }
        __syncwarp();
        TR(12, j);
        if (++ks == ATT_K_STAGES) { ks = 0; kph ^= 1; }
      }
      TR_DONE();
    } else if (warp == 3) {
      // ---------------------------------------------------------------- MMA issuer 2: O += P V_j
      constexpr uint32_t IDESC_O = umma_idesc_bf16(ATT_BQ, ATT_HD, 0, 1);        // P V: V is MN-major
      TR_DECL(3);
      int vs = 0;
      uint32_t vph = 0;
#pragma unroll 1
      for (int j = 0; j < nkv; ++j) {
        mbar_wait_a(sb + ATT_BAR_V_FULL + 8 * vs, vph);
        TR(14, j);
        named_bar_sync(ATT_BAR_P_FULL, ATT_BAR_COUNT);
        tc_fence_after();
        TR(15, j);
        if (elect_one()) {
          // V tile: [kv 128 rows][64 d] bf16, 128 B rows, SW128 -> MN-major B operand. One UMMA_K = 16 kv rows = 2048 B
          // = +128 in the descriptor's (>>4) start-address field.
          const uint32_t tb = opaque_u32(tmem_base);
          const uint64_t vd = umma_desc_sw128(sb + ATT_OFF_V + vs * ATT_TILE_BYTES, 1024, 1024);
#pragma unroll
          for (int k = 0; k < ATT_BKV / 16; ++k)
            umma_ts(tb + 128, tb + 192 + k * 8, vd + 128 * k, IDESC_O, (k != 0) || (j != 0));
          umma_commit_a(sb + ATT_BAR_V_EMPTY + 8 * vs);
          umma_commit_a(sb + ATT_BAR_O_FULL);
        }
        __syncwarp();
        TR(16, j);
        if (++vs == ATT_V_STAGES) { vs = 0; vph ^= 1; }
      }
      TR_DONE();
    }
  } else {
    // ------------------------------------------------------------------ softmax: 8 warps, TWO threads per query row
    // Warps 4-7 take kv columns 0..63 of every 128-wide score tile, warps 8-11 columns 64..127 (a warp may only touch the
    // TMEM lane quarter warp % 4). Four softmax warps per scheduler (two CTAs per SM) instead of two hide the
    // fixed-latency stalls of the exp chains; 64 scores per thread fit 104 registers.
    //
    // No row maximum. The two threads of a row would have to exchange it before the first exp; instead the exponent
    // reference r = r0 + K (log2 units) is LAGGED: r0 is the row's first score, and K is raised at the start of tile j
    // from the sum of tile j-2's probabilities (both halves, through shared memory -- visible by then without an extra
    // barrier: each half publishes its sum before it arrives on "P stored", and the other half has since waited for the
    // commit of the P V product that arrival released). sum <= 128 max, so floor(log2 sum) stands for the maximum to
    // within 7 binary orders, which is all a reference needs: it only keeps the numbers inside fp32 / bf16 range; the
    // result O / l is exact for any reference. K moves by integers (exact power-of-two rescale of O and l) and only
    // when the tile outgrew the reference by 2^40.
    // What a lagged reference cannot rule out is a score more than 2^100 above it (a jump of ~69 in the natural-log
    // domain inside two kv tiles, or over the row's first score): such a row's sums reach 2^100 / inf, the thread raises
    // a flag, and the CTA recomputes its rows with attention_row_exact after the fast path. Never seen on real
    // attention scores; tests/test_gpu_attention.py forces it.
    setmaxnreg_inc<ATT_SOFTMAX_REGS>();
    const int half = (warp - 4) >> 2;
    const uint32_t row = (warp & 3) * 32 + lane;
    // per-thread TMEM base (lane quarter in bits 16+): S at tSi (this half's 64 columns), O at tOP + 128, P at tOP + 192
    const uint32_t tOP = opaque_u32(tmem_base + ((row & ~31u) << 16) + half * 32);
    const uint32_t tSi = opaque_u32(tOP + half * 32);
    // tile sums: slot * 1024 + half * 512 + row * 4
    const uint32_t my_part = opaque_u32(sb + ATT_OFF_PART + half * (ATT_BQ * 4) + row * 4);
    const uint32_t other_part = my_part ^ (ATT_BQ * 4);      // (ATT_OFF_PART is a multiple of 1024)
    TR_DECL(2);
#ifdef ATT_TRACE
    if (warp != 4) tr.slot = -1;
#endif
    // reference base r0: the row's first score (column 0 of tile 0), the same for both halves, log2 units
    constexpr float SCALE = QLOG2 ? 1.0f : LOG2E;
    float neg_r0;
    bool raw = false;                              // warp-uniform: exponent = accumulator value (no reference at all)
    {
      mbar_wait_a(sb + ATT_BAR_S_FULL, 0);
      tc_fence_after();
      uint32_t s_first[1];
      tmem_ld_32x1(tSi - half * 64, s_first);
      tmem_ld_wait();
      neg_r0 = -__uint_as_float(s_first[0]) * SCALE;
      if (QLOG2) {
        // both halves of a row sit in warps that hold the same 32 rows: they take the same decision
        raw = __all_sync(0xffffffffu, fabsf(neg_r0) < ATT_RAW_LIMIT);
        if (raw) neg_r0 = 0.f;
      }
    }
    int k_m1 = 0, k_m2 = 0;                        // integer reference offsets of tiles j-1 and j-2
    unsigned long long l2 = pk2(0.f, 0.f);         // this half's row sum (two partial sums), relative to r0 + k_m1
    float biggest = 0.f;                           // largest tile sum seen (inf sticks): the overflow sentinel
    const unsigned long long SCALE2 = pk2(SCALE, SCALE);

    // 16 scores -> 8 packed bf16 pairs of probabilities; KB = first column (decides which pairs take the FMA-pipe exp2)
#define ATT_EXPS16(sv, pv, KB, RAW)                                                                                   \
  _Pragma("unroll") for (int k = (KB); k < (KB) + 16; k += 2) {                                                        \
    const unsigned long long s2 = pk2(__uint_as_float(sv[k]), __uint_as_float(sv[k + 1]));                             \
    const unsigned long long x2 = (RAW) ? s2 : ffma2(s2, SCALE2, nm2);                                                 \
    float p0, p1;                                                                                                      \
    if (((k >> 1) % POLY_DEN) < POLY_NUM) {                                                                            \
      poly_exp2_pair(x2, p0, p1);                                                                                      \
    } else {                                                                                                           \
      float x0, x1;                                                                                                    \
      unpk2(x2, x0, x1);                                                                                               \
      p0 = fast_exp2(x0);                                                                                              \
      p1 = fast_exp2(x1);                                                                                              \
    }                                                                                                                  \
    if ((k >> 1) & 1) t2b = fadd2(t2b, pk2(p0, p1));                                                                   \
    else t2a = fadd2(t2a, pk2(p0, p1));                                                                                \
    pv[k >> 1] = pack_bf16(p0, p1);                                                                                    \
  }

#pragma unroll 1
    for (int j = 0; j < nkv; ++j) {
      TR(20, j);
      mbar_wait_a(sb + ATT_BAR_S_FULL, j & 1);
      tc_fence_after();
      TR(21, j);
      uint32_t s0[32], s1[32];
      tmem_ld_32x32(tSi, s0);
      // reference offset of this tile, from the sums of tile j-2 (under the TMEM load)
      int k_cur = k_m1;
      if (j >= 2) {
        const uint32_t slot = ((j + 2) & 3) * 1024;
        const float tsum = lds_f32(my_part + slot) + lds_f32(other_part + slot);                 // both halves, relative to k_m2
        const int mag = k_m2 + static_cast<int>((__float_as_uint(tsum) >> 23) & 0xff) - 127;     // ~ log2 of the largest p
        if (mag > k_m1 + ATT_TAU_LOG2) k_cur = mag - 4;
      }
      const int dk = k_cur - k_m1;                                           // 0, or >= ATT_TAU_LOG2 - 3
      const bool rescale = __any_sync(0xffffffffu, dk != 0);
      if (rescale) raw = false;                                              // (rare) from now on the general form
      k_m2 = k_m1;
      k_m1 = k_cur;
      const float nm = neg_r0 - static_cast<float>(k_cur);
      const unsigned long long nm2 = pk2(nm, nm);
      const int kv_valid = T - j * ATT_BKV - half * 64;                      // < 64 only in the last tile: zero-filled K rows
      tmem_ld_wait_regs(s0);
      if (kv_valid < 32) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k >= kv_valid) s0[k] = 0xff800000u;  // -inf
      }
      // p = 2^(s*log2e - r); P -> TMEM as bf16 (columns kk/2), tile sums in fp32
      unsigned long long t2a = pk2(0.f, 0.f), t2b = pk2(0.f, 0.f);
      uint32_t pa[16], pb[16];
      if (QLOG2 && raw) {
        ATT_EXPS16(s0, pa, 0, true);
      } else {
        ATT_EXPS16(s0, pa, 0, false);
      }
      tmem_ld_32x32(tSi + 32, s1);                 // the other 32 scores land under the second quarter's exps
      if (QLOG2 && raw) {
        ATT_EXPS16(s0, pa, 16, true);
      } else {
        ATT_EXPS16(s0, pa, 16, false);
      }
      tmem_ld_wait_regs(s1);
      tc_fence_before();
      TR(22, j);
      if (j + 1 < nkv) named_bar_arrive(ATT_BAR_S_EMPTY, ATT_BAR_COUNT);   // S is in registers: Q K_{j+1}^T may overwrite it
      if (kv_valid < 64) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k + 32 >= kv_valid) s1[k] = 0xff800000u;
      }
      if (QLOG2 && raw) {
        ATT_EXPS16(s1, pb, 0, true);
        ATT_EXPS16(s1, pb, 16, true);
      } else {
        ATT_EXPS16(s1, pb, 0, false);
        ATT_EXPS16(s1, pb, 16, false);
      }
      float down = 1.0f;
      if (rescale) {                                                         // rare
        if (dk != 0) down = dk < 127 ? __int_as_float((127 - dk) << 23) : 0.f;   // 2^-dk, exact
        l2 = fmul2(l2, pk2(down, down));
      }
      {
        const unsigned long long t2 = fadd2(t2a, t2b);
        l2 = fadd2(l2, t2);
        float ta, tb;
        unpk2(t2, ta, tb);
        const float tsum = ta + tb;
        sts_f32(my_part + (j & 3) * 1024, tsum);
        biggest = fmaxf(biggest, tsum);
      }
      if (j > 0) {                                 // P is still being read by the previous tile's P V until then
        TR(23, j);
        mbar_wait_a(sb + ATT_BAR_O_FULL, (j - 1) & 1);
        tc_fence_after();
        TR(24, j);
        if (rescale) {                             // rare: O (this half's 32 columns) follows the reference
          uint32_t r[32];
          tmem_ld_32x32(tOP + 128, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * down);
          tmem_st_32x32(tOP + 128, r);
        }
      }
      tmem_st_32x16(tOP + 192, pa);
      tmem_st_32x16(tOP + 192 + 16, pb);
      TR(25, j);
      tmem_st_wait();
      tc_fence_before();
      TR(26, j);
      named_bar_arrive(ATT_BAR_P_FULL, ATT_BAR_COUNT);
    }
    // the two halves of a row: total l, and whether any row of the tile needs the exact path
    {
      float la, lb;
      unpk2(l2, la, lb);
      sts_f32(my_part + 4 * 1024, la + lb);      // slot 4: the tile-sum slots may still be read by the other half
      if (!(biggest < ATT_RISK_SUM)) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sb + ATT_OFF_EXACT), "r"(1u) : "memory");
    }
    named_bar_sync(ATT_BAR_SOFTMAX, 256);
    const float inv_l = 1.0f / (lds_f32(my_part + 4 * 1024) + lds_f32(other_part + 4 * 1024));
    uint32_t exact;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(exact) : "r"(sb + ATT_OFF_EXACT) : "memory");
    // normalise and store this half's 32 columns
    mbar_wait_a(sb + ATT_BAR_O_FULL, (nkv - 1) & 1);
    tc_fence_after();
    TR(27, nkv);
    const int d = H * ATT_HD, h = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * ATT_BQ;
    {
      // The normalised tile leaves through ONE TMA store: each thread writes its 64 bytes into the Q buffer (dead since
      // the last Q K^T ran) in the 128-byte-swizzled layout of the output map; rows >= T are clipped by the map. (Stored
      // straight from the threads, each lane's 64 bytes go to a different 2.5 KB-strided row: 32 cache lines per
      // instruction, ~1000 L1 tag cycles per tile.)
      uint32_t r[32];
      tmem_ld_32x32(tOP + 128, r);
      tmem_ld_wait();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t v0 = pack_bf16(__uint_as_float(r[8 * u]) * inv_l, __uint_as_float(r[8 * u + 1]) * inv_l);
        const uint32_t v1 = pack_bf16(__uint_as_float(r[8 * u + 2]) * inv_l, __uint_as_float(r[8 * u + 3]) * inv_l);
        const uint32_t v2 = pack_bf16(__uint_as_float(r[8 * u + 4]) * inv_l, __uint_as_float(r[8 * u + 5]) * inv_l);
        const uint32_t v3 = pack_bf16(__uint_as_float(r[8 * u + 6]) * inv_l, __uint_as_float(r[8 * u + 7]) * inv_l);
        const uint32_t chunk = static_cast<uint32_t>(half * 4 + u);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                     ::"r"(sb + ATT_OFF_Q + row * 128 + ((chunk ^ (row & 7)) << 4)), "r"(v0), "r"(v1), "r"(v2), "r"(v3)
                     : "memory");
      }
      fence_proxy_async_smem();
      named_bar_sync(ATT_BAR_SOFTMAX, 256);
      if (warp == 4 && elect_one()) {
        tma_store_3d_a(&tmOut, sb + ATT_OFF_Q, h * ATT_HD, q0, b);
        tma_commit_group();
        if (exact) tma_wait_group<0>();            // the exact path below overwrites rows of this tile with plain stores
        else tma_wait_group_read<0>();             // shared memory must outlive the read
      }
    }
    TR(28, nkv);
    TR_DONE();
    if (exact) {
      named_bar_sync(ATT_BAR_SOFTMAX, 256);        // every fast-path store of the tile is out before it is overwritten
      const __nv_bfloat16* qkv_b = qkv + static_cast<size_t>(b) * T * 3 * d;
      for (int r = warp - 4; r < ATT_BQ && q0 + r < T; r += 8)
        attention_row_exact(qkv_b, out + (static_cast<size_t>(b) * T + q0 + r) * d + h * ATT_HD, q0 + r, T, d, h, lane, SCALE);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<256>(tmem_base);
#ifdef ATT_CYCLES
  if (threadIdx.x == 0) {
    atomicAdd(&g_att_cycles[0], static_cast<unsigned long long>(clock64() - cyc0));
    atomicAdd(&g_att_cycles[1], 1ull);
  }
#endif
}

#ifdef ATT_CYCLES
void att_cycles_read(unsigned long long* out2, bool reset) {
  cudaMemcpyFromSymbol(out2, g_att_cycles, 16);
  if (reset) {
    unsigned long long z[2] = {0, 0};
    cudaMemcpyToSymbol(g_att_cycles, z, 16);
  }
}
#endif
#ifdef ATT_TRACE
void att_trace_reset() {
  static uint32_t zeros[TR_SLOTS][4];
  memset(zeros, 0, sizeof(zeros));
  cudaMemcpyToSymbol(g_att_trace_n, zeros, sizeof(zeros));
}
void att_trace_dump() {
  static uint32_t ev[TR_SLOTS][4][TR_EVENTS][2];
  static uint32_t n[TR_SLOTS][4], sm[TR_SLOTS];
  cudaMemcpyFromSymbol(ev, g_att_trace, sizeof(ev));
  cudaMemcpyFromSymbol(n, g_att_trace_n, sizeof(n));
  cudaMemcpyFromSymbol(sm, g_att_trace_sm, sizeof(sm));
  for (int s = 0; s < TR_SLOTS; ++s)
    for (int r = 0; r < 4; ++r)
      for (uint32_t i = 0; i < n[s][r]; ++i)
        printf("TRACE slot %d sm %u role %d tag %u j %u clk %u\n", s, sm[s], r, ev[s][r][i][0] >> 16, ev[s][r][i][0] & 0xffff,
               ev[s][r][i][1]);
}
#endif

int launch_attention(const CUtensorMap& tm_qkv, const void* qkv, void* out, int B, int T, int H, int q_log2, cudaStream_t stream) {
  // output map [B][T][H * 64] bf16, box = one (128-row tile, head); rebuilt only when the buffer or the shape changes (the
  // encoder plan calls this 32 times per step on the same buffer)
  static thread_local struct { const void* out; int B, T, H; CUtensorMap tm; } oc = {nullptr, 0, 0, 0, {}};
  if (oc.out != out || oc.B != B || oc.T != T || oc.H != H) {
    const uint64_t dd = static_cast<uint64_t>(H) * ATT_HD;
    const uint64_t dims[3] = {dd, static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    const uint64_t str[3] = {2, dd * 2, dd * 2 * static_cast<uint64_t>(T)};
    const uint32_t box[3] = {ATT_HD, ATT_BQ, 1};
    if (make_tmap(&oc.tm, out, 2, 3, dims, str, box, true) != 0) return -1;
    oc.out = out; oc.B = B; oc.T = T; oc.H = H;
  }
  static bool attr_set = false;
  if (!attr_set) {
    AL_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    AL_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  dim3 grid((T + ATT_BQ - 1) / ATT_BQ, H, B);
  if (q_log2)
    attention_fwd_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, oc.tm, reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                                        reinterpret_cast<__nv_bfloat16*>(out), T, H);
  else
    attention_fwd_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tm_qkv, oc.tm, reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                                         reinterpret_cast<__nv_bfloat16*>(out), T, H);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace al
