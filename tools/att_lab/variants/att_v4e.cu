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
#include "../../../audio_llama_b200/csrc/common.cuh"
#include "../../../audio_llama_b200/csrc/kernels.h"

namespace al {

constexpr int ATT_BQ = 128;       // query rows per tile
constexpr int ATT_BKV = 128;      // kv rows per tile
constexpr int ATT_HD = 64;
constexpr int ATT_K_STAGES = 3;
constexpr int ATT_V_STAGES = 2;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;   // 16 KB: any of Q / K / V tile
constexpr int ATT_SMEM = (1 + ATT_K_STAGES + ATT_V_STAGES) * ATT_TILE_BYTES + 256 + 1024;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float ATT_TAU = 8.0f;   // lazy-rescale threshold, log2 units
constexpr int POLY_PAIRS = 1;     // of every 4 pairs, how many use the FMA-pipe exp2
// softmax warps -> MMA warp hand-offs are named barriers (128 arrive + 32 sync); id 0 is __syncthreads
constexpr int ATT_BAR_P_FULL = 1;
constexpr int ATT_BAR_S_EMPTY = 2;

#ifdef ATT_TRACE
// Development-only event timeline (tools/att_lab): lane 0 of each role of a few CTAs records (tag, kv tile, %clock).
constexpr int TR_SLOTS = 32, TR_EVENTS = 256;
__device__ uint32_t g_att_trace[TR_SLOTS][3][TR_EVENTS][2];
__device__ uint32_t g_att_trace_n[TR_SLOTS][3];
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
// 2^x for a pair, x <= ~9, on the FMA / ALU pipes.
__device__ __forceinline__ void poly_exp2_pair(unsigned long long x2, float& o0, float& o1) {
  float x0, x1;
  unpk2(x2, x0, x1);
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const unsigned long long xc = pk2(x0, x1);
  const unsigned long long MAGIC = pk2(12582912.0f, 12582912.0f);          // 1.5 * 2^23
  const unsigned long long NMAGIC = pk2(-12582912.0f, -12582912.0f);
  const unsigned long long NEG1 = pk2(-1.0f, -1.0f);
  const unsigned long long t = fadd2(xc, MAGIC);                           // low mantissa bits = round(x)
  const unsigned long long xr = fadd2(t, NMAGIC);
  const unsigned long long f = ffma2(xr, NEG1, xc);                        // x - round(x) in [-0.5, 0.5]
  unsigned long long p = ffma2(pk2(0.05508868396282196f, 0.05508868396282196f), f,
                               pk2(0.24260404706001282f, 0.24260404706001282f));
  p = ffma2(p, f, pk2(0.6932762265205383f, 0.6932762265205383f));
  p = ffma2(p, f, pk2(0.9999289512634277f, 0.9999289512634277f));
  float t0, t1, p0, p1;
  unpk2(t, t0, t1);
  unpk2(p, p0, p1);
  o0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  o1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

__global__ void __launch_bounds__(256, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, int T, int H) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 1 tile
  uint8_t* sK = smem + ATT_TILE_BYTES;                  // K ring
  uint8_t* sV = sK + ATT_K_STAGES * ATT_TILE_BYTES;     // V ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_V_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                              // 1
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + ATT_K_STAGES;
  uint64_t* v_full = k_empty + ATT_K_STAGES;
  uint64_t* v_empty = v_full + ATT_V_STAGES;
  uint64_t* s_full = v_empty + ATT_V_STAGES;            // "S drained" and "P stored" are named barriers (see below)
  uint64_t* o_full = s_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int d = H * ATT_HD;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int q0 = blockIdx.x * ATT_BQ;
  const int nkv = (T + ATT_BKV - 1) / ATT_BKV;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_K_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < ATT_V_STAGES; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
    // the first loads need nothing but their barriers: they start before the TMEM allocation and the CTA-wide sync
    mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
    tma_load_3d(sQ, &tmQKV, q_full, h * ATT_HD, q0, b);
    mbar_arrive_expect_tx(&k_full[0], ATT_TILE_BYTES);
    tma_load_3d(sK, &tmQKV, &k_full[0], d + h * ATT_HD, 0, b);
  }
  if (warp == 2) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tS = tmem_base;            // 128 columns
  const uint32_t tO = tmem_base + 128;      // 64 columns
  const uint32_t tP = tmem_base + 192;      // 64 columns (bf16 pairs: 128 kv -> 64 columns)

  // Register budget: launched with 128 regs x 256 threads (2 CTAs / SM). The 4 control warps drop to 48, which
  // frees 80 x 128 = 10240 registers; the 4 softmax warps grow to 208, which takes 80 x 128 = 10240. (Asking for
  // more than was freed makes setmaxnreg.inc wait forever.)
  if (warp < 4) {
    setmaxnreg_dec<48>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      TR_DECL(0);
      if (elect_one()) {   // elect.sync, not lane == 0: ptxas then knows ONE thread issues and emits no per-lane waterfall loop around the uniform-datapath instructions
        // (Q and K_0 were requested by this same thread before the CTA-wide sync)
        // K runs ahead of V: K_{j+1} is requested before V_j so that Q K_{j+1}^T is never starved
        int ks = 1, vs = 0;
        uint32_t kph = 0, vph = 0;
        auto load_k = [&](int j) {
          mbar_wait(&k_empty[ks], kph ^ 1);
          TR(1, j);
          mbar_arrive_expect_tx(&k_full[ks], ATT_TILE_BYTES);
          tma_load_3d(sK + ks * ATT_TILE_BYTES, &tmQKV, &k_full[ks], d + h * ATT_HD, j * ATT_BKV, b);
          if (++ks == ATT_K_STAGES) { ks = 0; kph ^= 1; }
        };
        for (int j = 0; j < nkv; ++j) {
          if (j + 1 < nkv) load_k(j + 1);
          mbar_wait(&v_empty[vs], vph ^ 1);
          TR(2, j);
          mbar_arrive_expect_tx(&v_full[vs], ATT_TILE_BYTES);
          tma_load_3d(sV + vs * ATT_TILE_BYTES, &tmQKV, &v_full[vs], 2 * d + h * ATT_HD, j * ATT_BKV, b);
          if (++vs == ATT_V_STAGES) { vs = 0; vph ^= 1; }
        }
      }
      TR_DONE();
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t IDESC_S = umma_idesc_bf16(ATT_BQ, ATT_BKV);             // Q K^T: both K-major
      constexpr uint32_t IDESC_O = umma_idesc_bf16(ATT_BQ, ATT_HD, 0, 1);        // P V: V is MN-major
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t k_addr = smem_u32(sK);
      const uint32_t v_addr = smem_u32(sV);
      auto issue_s = [&](int stage) {
        const uint64_t qd = umma_desc_sw128(q_addr, 16, 1024);
        const uint64_t kd = umma_desc_sw128(k_addr + stage * ATT_TILE_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS, qd + 2 * k, kd + 2 * k, IDESC_S, k != 0);
        umma_commit(&k_empty[stage]);          // the K stage is free once these MMAs have run
        umma_commit(s_full);
      };
      auto issue_o = [&](int stage, bool first_tile) {
        // V tile: [kv 128 rows][64 d] bf16, 128 B rows, SW128 -> MN-major B operand. One UMMA_K = 16 kv rows = 2048 B
        // = +128 in the descriptor's (>>4) start-address field.
        const uint64_t vd = umma_desc_sw128(v_addr + stage * ATT_TILE_BYTES, 1024, 1024);
#pragma unroll
        for (int k = 0; k < ATT_BKV / 16; ++k)
          umma_ts(tO, tP + k * 8, vd + 128 * k, IDESC_O, (k != 0) || !first_tile);
        umma_commit(&v_empty[stage]);
        umma_commit(o_full);
      };
      TR_DECL(1);
      TR(10, 0);
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      TR(11, 0);
      if (elect_one()) issue_s(0);
      __syncwarp();
      TR(12, 0);
      int ks = 1, vs = 0;                            // next K stage to consume, current V stage
      uint32_t kph = 0, vph = 0;
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) {                           // scores of the next kv tile as soon as S is drained
          mbar_wait(&k_full[ks], kph);
          TR(13, j + 1);
          named_bar_sync(ATT_BAR_S_EMPTY, 160);        // blocks in hardware: no polling next to the softmax warps
          tc_fence_after();
          TR(11, j + 1);
          if (elect_one()) issue_s(ks);
          __syncwarp();
          TR(12, j + 1);
          if (++ks == ATT_K_STAGES) { ks = 0; kph ^= 1; }
        }
        mbar_wait(&v_full[vs], vph);
        TR(14, j);
        named_bar_sync(ATT_BAR_P_FULL, 160);
        tc_fence_after();
        TR(15, j);
        if (elect_one()) issue_o(vs, j == 0);
        __syncwarp();
        TR(16, j);
        if (++vs == ATT_V_STAGES) { vs = 0; vph ^= 1; }
      }
      TR_DONE();
    }
  } else {
    // ------------------------------------------------------------------ softmax warpgroup
    setmaxnreg_inc<208>();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tSi = tS + lane_off;
    const uint32_t tOi = tO + lane_off;
    const uint32_t tPi = tP + lane_off;
    float m_ref = -INFINITY;                       // reference max, log2 units (score * log2 e)
    unsigned long long l2a = pk2(0.f, 0.f), l2b = pk2(0.f, 0.f);   // row-sum accumulators (4 partial sums)
    const unsigned long long LOG2E2 = pk2(LOG2E, LOG2E);
    TR_DECL(2);
#ifdef ATT_TRACE
    if (warp != 4) tr.slot = -1;
#endif

    for (int j = 0; j < nkv; ++j) {
      TR(20, j);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      TR(21, j);
      uint32_t s[128];
      {
        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
        uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
        uint32_t(&s3)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
        tmem_ld_32x32(tSi, s0);
        tmem_ld_32x32(tSi + 32, s1);
        tmem_ld_32x32(tSi + 64, s2);
        tmem_ld_32x32(tSi + 96, s3);
        tmem_ld_wait();
      }
      tc_fence_before();
      TR(22, j);
      if (j + 1 < nkv) named_bar_arrive(ATT_BAR_S_EMPTY, 160);   // S is in registers: Q K_{j+1}^T may overwrite it now
      if (j == nkv - 1) {                          // kv tail: columns >= kv_valid are zero-filled K rows
        const int kv_valid = T - j * ATT_BKV;
        if (kv_valid < ATT_BKV) {
#pragma unroll
          for (int k = 0; k < 128; ++k)
            if (k >= kv_valid) s[k] = 0xff800000u;  // -inf
        }
      }
      // row max: 4 independent FMNMX3 chains
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]), mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int k = 4; k < 128; k += 8) {
        mx0 = fmax3(mx0, __uint_as_float(s[k]), __uint_as_float(s[k + 1]));
        mx1 = fmax3(mx1, __uint_as_float(s[k + 2]), __uint_as_float(s[k + 3]));
        if (k + 4 < 128) {
          mx2 = fmax3(mx2, __uint_as_float(s[k + 4]), __uint_as_float(s[k + 5]));
          mx3 = fmax3(mx3, __uint_as_float(s[k + 6]), __uint_as_float(s[k + 7]));
        }
      }
      const float mx_s = fmaxf(fmax3(mx0, mx1, mx2), mx3) * LOG2E;
      bool o_ready = (j == 0);
      if (__any_sync(0xffffffffu, mx_s > m_ref + ATT_TAU)) {
        // advance the reference (whole warp, so the TMEM ld/st below stay warp-uniform) and rescale l and O
        const float new_ref = fmaxf(m_ref, mx_s);
        const float scale = fast_exp2(m_ref - new_ref);        // 0 on the first tile (m_ref = -inf)
        m_ref = new_ref;
        const unsigned long long sc2 = pk2(scale, scale);
        l2a = fmul2(l2a, sc2);
        l2b = fmul2(l2b, sc2);
        if (j > 0) {
          mbar_wait(o_full, (j - 1) & 1);                      // P V of the previous tile has landed in O
          tc_fence_after();
          o_ready = true;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tOi + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(__uint_as_float(r[k]) * scale);
            tmem_st_32x32(tOi + c * 32, r);
          }
          tmem_st_wait();
        }
      }
      const float nm = -m_ref;
      const unsigned long long nm2 = pk2(nm, nm);
      // p = 2^(s*log2e - m_ref); P -> TMEM as bf16 (columns kk/2), row sums in fp32
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          const int idx = c * 32 + k;
          const unsigned long long x2 = ffma2(pk2(__uint_as_float(s[idx]), __uint_as_float(s[idx + 1])), LOG2E2, nm2);
          float p0, p1;
          if (((k >> 1) & 3) < POLY_PAIRS) {
            poly_exp2_pair(x2, p0, p1);
          } else {
            float x0, x1;
            unpk2(x2, x0, x1);
            p0 = fast_exp2(x0);
            p1 = fast_exp2(x1);
          }
          if ((k >> 1) & 1) l2b = fadd2(l2b, pk2(p0, p1));
          else l2a = fadd2(l2a, pk2(p0, p1));
          pk[k >> 1] = pack_bf16(p0, p1);
        }
        if (c == 0 && !o_ready) {                  // P is still being read by the previous tile's P V until then
          TR(23, j);
          mbar_wait(o_full, (j - 1) & 1);
          tc_fence_after();
          TR(24, j);
        }
        tmem_st_32x16(tPi + c * 16, pk);
      }
      TR(25, j);
      tmem_st_wait();
      tc_fence_before();
      TR(26, j);
      named_bar_arrive(ATT_BAR_P_FULL, 160);
    }
    // normalise and store
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    TR(27, nkv);
    float la, lb, lc, ld;
    unpk2(l2a, la, lb);
    unpk2(l2b, lc, ld);
    const float inv_l = 1.0f / ((la + lb) + (lc + ld));
    const int q = q0 + row;
    uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * T + q) * d + h * ATT_HD);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tOi + c * 32, r);
      tmem_ld_wait();
      if (q < T) {
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
    TR(28, nkv);
    TR_DONE();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<256>(tmem_base);
}

#ifdef ATT_TRACE
void att_trace_reset() {
  static uint32_t zeros[TR_SLOTS][3];
  memset(zeros, 0, sizeof(zeros));
  cudaMemcpyToSymbol(g_att_trace_n, zeros, sizeof(zeros));
}
void att_trace_dump() {
  static uint32_t ev[TR_SLOTS][3][TR_EVENTS][2];
  static uint32_t n[TR_SLOTS][3], sm[TR_SLOTS];
  cudaMemcpyFromSymbol(ev, g_att_trace, sizeof(ev));
  cudaMemcpyFromSymbol(n, g_att_trace_n, sizeof(n));
  cudaMemcpyFromSymbol(sm, g_att_trace_sm, sizeof(sm));
  for (int s = 0; s < TR_SLOTS; ++s)
    for (int r = 0; r < 3; ++r)
      for (uint32_t i = 0; i < n[s][r]; ++i)
        printf("TRACE slot %d sm %u role %d tag %u j %u clk %u\n", s, sm[s], r, ev[s][r][i][0] >> 16, ev[s][r][i][0] & 0xffff,
               ev[s][r][i][1]);
}
#endif

int launch_attention(const CUtensorMap& tm_qkv, void* out, int B, int T, int H, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    AL_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  dim3 grid((T + ATT_BQ - 1) / ATT_BQ, H, B);
  attention_fwd_kernel<<<grid, 256, ATT_SMEM, stream>>>(tm_qkv, reinterpret_cast<__nv_bfloat16*>(out), T, H);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace al
