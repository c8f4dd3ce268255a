// Tensor-core log-mel front end (fp32-accurate): the 400-point real DFT of every frame as four ~100 x 100
// products on tcgen05, operands split into fp16 hi + lo (three products: hi*hi, lo*hi, hi*lo, fp32 accumulation
// in TMEM), then |.|^2, the banded mel projection and the log on the CUDA cores.
//
// Replaces the same reference code as mel.cu (HF feature_extraction_whisper.py:135-164 as called by
// /root/reference/src/inference.py:100-105; mode 1 = /root/reference/src/dataset.py:125-133). mel.cu stays as the
// CUDA-core FFT form (any filter bank; AUDIOLLM_B200_MEL=fft or al_mel_set_mode(0)).
//
// Why a product and not an FFT: an FFT-400 costs ~9 k fp32 operations per frame on the CUDA cores, which puts the
// instruction-issue floor of the whole front end ABOVE its HBM time (3.46 MB per 30 s clip = 0.53 us); the dense
// DFT is 3 x 84 tensor-pipe dispatches per 256 frames (0.41 us per clip at the tensor rate) and leaves the CUDA
// cores ~4 k operations per frame (fold, split, power, mel, log).
//
// Math. y[n] = w[n] x[160 t + n], w = periodic Hann (w[0] = 0, w[400-n] = w[n], w[200-n] = 1 - w[n]).
//   fold 1 (n <-> 400-n): e[n] = y[n] + y[400-n], o[n] = y[n] - y[400-n]          (n = 1..199), y[200] alone
//   fold 2 (n <-> 200-n), by the parity of the output bin k:
//     Re X[2j]   =  sum_i ce[i] cos(2 pi i 2j / 400)        ce[i] = e[i] + e[200-i]
//     Im X[2j]   = -sum_i se[i] sin(2 pi i 2j / 400)        se[i] = o[i] - o[200-i]
//     Re X[2j+1] =  sum_i co[i] cos(2 pi i (2j+1) / 400)    co[i] = e[i] - e[200-i]
//     Im X[2j+1] = -sum_i so[i] sin(2 pi i (2j+1) / 400)    so[i] = o[i] + o[200-i]
//   with i = 0..100; the generic expressions evaluated at i = 0 and i = 100 give 2 y[200], 2 (y[100] +- y[300]),
//   so rows 0 and 100 of the twiddle matrices carry a factor 0.5 (exact). K is padded to 112 (window 0 beyond 100).
//
// One CTA PAIR (cluster of 2, tcgen05.mma.cta_group::2, M = 256) per 256 consecutive frames of a clip; each CTA
// owns 128 frames (TMEM lane = frame) and keeps HALF of the twiddle rows (56 of 112 output bins per matrix, hi and
// lo: 98 KB) resident in shared memory for the whole persistent kernel. Per CTA:
//   warp 0      loader: 32-frame sample slots (ring of 5) laid out as 34 rows of 164 floats (one 160-sample hop + 4
//               per row; the 16-byte destination alignment of TMA allows no odd pitch, so the thread-per-frame scalar
//               reads are 4-way bank conflicted). An interior slot is ONE 3-D tensor copy of the overlapping-row view
//               [clip][hop][164] of the wave buffer; slots at the clip edges (reflect padding, zero padding past
//               n_samples) fall back to one bulk copy per in-range hop plus element-wise fills by the warp itself.
//   warp 1      MMA issuer (leader CTA only): A from TMEM, B from shared memory (no-swizzle K-major). Even bins: 6 MMAs
//               per K block behind a two-stage operand hand-off; odd bins: 42 MMAs back to back once the builders have
//               stored all seven blocks and the mel warps have drained the even-bin accumulators.
//   warp 2      TMEM allocator (all 512 columns: 2 accumulators x 112 used for the even and then the odd bins, the odd-bin
//               operands of all 7 K blocks (7 x 32), and a ring of two even-bin operand stages (2 x 32)).
//   warps 4-11  operand builders, one thread per (frame, half of a 16-wide K chunk): fold, window, power-of-two
//               scale (per 32-frame slot, so any input amplitude fits fp16), hi/lo split, tcgen05.st.
//   warps 12-15 one thread per frame: TMEM -> |X|^2 -> mel sums in registers (the bank is banded: every bin feeds
//               at most two mel bins, and WHICH two is a compile-time table, mel_bank_struct.h) -> log -> coalesced
//               128 B stores; per-clip max. The accumulators are released to the next tile's MMAs as soon as the
//               last bin has been read, before the logs and stores.
//
// Order of the contraction. The tensor core adds each 16-wide K block into the fp32 accumulator with truncation, so
// the error grows with the size of the PARTIAL sums. In natural order (block b = samples 16b..16b+15) the partial
// sums of a weak bin next to a strong tone are far larger than the final value and the result misses the 3e-5
// element-wise bound (measured 4.5e-5 .. 6.3e-5 on the synthetic clips). K block b therefore holds the decimated
// samples i = b + 7e (e = 0..15): every block sum is ~1/7 of the final sum and the partial sums stay at the size of
// the result (simulated and measured < 1e-5).
#include <cuda_fp16.h>
#include <string.h>

#include <type_traits>
#include <utility>

#include "common.cuh"
#include "kernels.h"
#include "mel_bank_struct.h"

namespace al {

constexpr int TC_N = 112;                               // output bins per parity, padded (101 / 100 used)
constexpr int TC_SEG = 160, TC_SEG_STRIDE = 164, TC_NSEG = 34;
constexpr int TC_SLOT_FLOATS = TC_NSEG * TC_SEG_STRIDE;          // 5576
constexpr int TC_SLOT_PITCH = 5600;                              // floats between slots: tensor copies need 128 B aligned destinations
constexpr int TC_SLOT_SAMPLES = 31 * 160 + 400;                  // 5360 samples feed 32 frames
constexpr int TC_NSLOT = 5;
constexpr int TC_TILES_PER_CLIP = 12;                            // ceil(3000 / 256)
constexpr int TC_THREADS = 512;
constexpr int TC_CLIP = 480000, TC_FRAMES = 3000;
// TMEM columns
// D (cos part | sin part) is used twice per tile: even bins first, odd bins after the mel warps have drained it.
// A_O holds the odd-bin operands of ALL seven K blocks (32 columns each), A_E is a ring of two 32-column stages.
constexpr uint32_t TC_D_C = 0, TC_D_S = 112, TC_A_O = 224, TC_A_E = 448;

constexpr int TC_OFF_B = 0;
constexpr int TC_OFF_SLOTS = TC_OFF_B + MEL_TC_B_BYTES;
constexpr int TC_OFF_BAR = TC_OFF_SLOTS + TC_NSLOT * TC_SLOT_PITCH * 4;
static_assert(TC_OFF_SLOTS % 128 == 0 && (TC_SLOT_PITCH * 4) % 128 == 0 && TC_SLOT_PITCH >= TC_SLOT_FLOATS, "slot alignment");
constexpr int TC_NBAR = 2 * TC_NSLOT + 2 + 2 + 2 + 2 + 2 + 1;
constexpr int TC_OFF_MISC = TC_OFF_BAR + TC_NBAR * 8;
constexpr int TC_SMEM = TC_OFF_MISC + 128 /*tmem ptr, slot max, scales*/ + 128 /*alignment slack*/;
static_assert(TC_SMEM <= 232448, "mel_tc shared memory over the 227 KB limit");

__constant__ float c_tc_wink[112][2];                // {w[i], w[200 - i]} at K position 16 b + e, i = b + 7 e (0 for i > 100)
__constant__ float c_tc_w[MEL_N_BANKS][201][2];      // filter weights of bin k for mel MEL_BANK_LO / _HI [bank][k]

int mel_tc_set_window(const float* host_win_2x112 /* [112 K positions][2] */) {
  AL_CHECK_CUDA(cudaMemcpyToSymbol(c_tc_wink, host_win_2x112, sizeof(float) * 2 * 112));
  return 0;
}
int mel_tc_set_weights(int bank, const float* host_w_201x2) {
  AL_CHECK_CUDA(cudaMemcpyToSymbol(c_tc_w, host_w_201x2, sizeof(float) * 201 * 2, sizeof(float) * 201 * 2 * bank));
  return 0;
}

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st_32x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// D[tmem, both CTAs] (+)= A[tmem, both CTAs] * B[smem, both CTAs]^T, fp16 operands, M = 256 across the pair.
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// No-swizzle K-major operand: 8-row x 16-byte core matrices; LBO = byte step between the two K halves of one MMA,
// SBO = byte step between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}
// kind::f16 with fp16 A/B (format fields 0), fp32 D.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ unsigned int tc_float_to_ordered(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// Arrive on the leader CTA's barrier: a plain local arrive from the leader itself, a remote arrive without the
// cluster-scope memory fence from the peer (what is handed over lives in TMEM and was ordered by tcgen05.wait /
// tcgen05.fence; mbarrier.arrive.release.cluster costs a MEMBAR.ALL.GPU per call).
__device__ __forceinline__ void arrive_on_leader(uint64_t* bar, uint32_t my_rank) {
  if (my_rank == 0) {
    mbar_arrive(bar);
  } else {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
  }
}
// Explicit shared-space loads: the slot pointers come out of an integer alignment cast, after which the compiler
// would fall back to generic LD.
template <int OFF_BYTES>
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF_BYTES));
  return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// shared-memory float offset of sample n of a frame whose sample 0 sits at a hop boundary of the slot
__host__ __device__ constexpr int tc_off(int n) { return n + 4 * ((n >= 160) + (n >= 320)); }

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(static_cast<F&&>(f));
  }
}

struct MelTcArgs {
  const float* wave;
  const int* n_samples;
  long long wave_stride;
  const uint8_t* b_image;     // [2][MEL_TC_B_BYTES]
  float* out;
  unsigned int* clip_max_bits;
  int n_clips;
  int use_tmap;               // 1: interior slots arrive by one 3-D tensor copy (box 164 x 34 rows of the overlapping-
                              // row view of the wave buffer), 0: one bulk copy per hop
};

// ------------------------------------------------------------------ operand builder: one K block of one frame half
// K block j holds the decimated samples i = j + 7 e (e = 0..15); a thread of half HH builds e = 8 HH .. 8 HH + 7,
// i.e. i = base + 7 t with base = j + 56 HH. The four samples of element t sit at compile-time distances from four
// running pointers (x[i], x[200 + i] move up with j; x[400 - i], x[200 - i] move down); the 4-float pad between hops
// shifts an address only when 400 - i < 320 or 200 - i < 160, which for a given HH depends on j for exactly one t.
template <int HH>
__device__ __forceinline__ void build_tile(uint32_t xb, float S, uint32_t t_lane, uint64_t* e_empty, uint64_t* e_full,
                                           uint64_t* o_empty, uint64_t* o_full, uint64_t* slot_done, uint32_t tile_it,
                                           uint32_t rank, int lane) {
  const uint32_t n_use0 = tile_it * 7u;
#pragma unroll 1
  for (int j = 0; j < 7; ++j) {
    const int base = j + 56 * HH;
    const uint32_t p0 = xb + 4 * base;                 // x[i]        at p0 + 28 t
    const uint32_t p2 = xb + 4 * (204 + base);         // x[200 + i]  at p2 + 28 t           (200 + i in [200, 311])
    const uint32_t p4 = xb + 4 * (408 - base);         // x[400 - i]  at p4 - 28 t (- 16 once 400 - i < 320)
    const uint32_t p1 = xb + 4 * (204 - base);         // x[200 - i]  at p1 - 28 t (- 16 once 200 - i < 160)
    const float4* wk = reinterpret_cast<const float4*>(&c_tc_wink[16 * j + 8 * HH][0]);
    float ce[8], se[8], co[8], so[8];
    static_for<0, 8>([&](auto tt) {
      constexpr int t = decltype(tt)::value;
      // pad corrections: i > 80 (x[400 - i]) and i > 40 (x[200 - i]) with i = base + 7 t, j = 0..6
      constexpr int i_lo = 56 * HH + 7 * t, i_hi = i_lo + 6;
      uint32_t a4 = p4, a1 = p1;
      if constexpr (i_lo > 80) a4 -= 16; else if constexpr (i_hi > 80) a4 -= (base + 7 * t > 80) ? 16u : 0u;
      if constexpr (i_lo > 40) a1 -= 16; else if constexpr (i_hi > 40) a1 -= (base + 7 * t > 40) ? 16u : 0u;
      const float xa = lds_f32<28 * t>(p0), xr = lds_f32<-28 * t>(a4);          // x[i], x[400 - i]
      const float ya = lds_f32<-28 * t>(a1), yr = lds_f32<28 * t>(p2);          // x[200 - i], x[200 + i]
      const float a = xa + xr, c = xa - xr, bb = ya + yr, d = ya - yr;
      const float4 w4 = wk[t >> 1];                     // {w[i], w[200 - i]} of elements t, t + 1
      const float ws = ((t & 1) ? w4.z : w4.x) * S, ws2 = ((t & 1) ? w4.w : w4.y) * S;
      const float r = ws2 * bb, s = ws2 * d;
      ce[t] = fmaf(ws, a, r);
      co[t] = fmaf(ws, a, -r);
      se[t] = fmaf(ws, c, -s);
      so[t] = fmaf(ws, c, s);
    });
    if (j == 6) {                  // last read of this slot's samples: hand it back to the loader before the stores
      __syncwarp();
      if (lane == 0) mbar_arrive(slot_done);
    }
    // hi = fp16(value), lo = fp16(value - hi): 22 significant bits between them
    auto split_store = [&](const float (&v)[8], uint32_t t_hi) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __half2 hp = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        const float2 hf = __half22float2(hp);
        h[e] = *reinterpret_cast<const uint32_t*>(&hp);
        l[e] = pack_h2(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
      }
      tmem_st_32x4(t_hi + 4 * HH, h[0], h[1], h[2], h[3]);
      tmem_st_32x4(t_hi + 8 + 4 * HH, l[0], l[1], l[2], l[3]);
    };
    // even-bin operands: ring of two stages, one hand-off per K block
    const uint32_t n_use = n_use0 + j;               // running K-block count of this CTA
    const uint32_t es = n_use & 1;
    mbar_wait(&e_empty[es], ((n_use >> 1) & 1) ^ 1);
    tc_fence_after();
    split_store(ce, t_lane + TC_A_E + 32 * es);
    split_store(se, t_lane + TC_A_E + 32 * es + 16);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) arrive_on_leader(&e_full[es], rank);
    // odd-bin operands: block j of the tile-wide buffer, handed over once per tile
    if (j == 0) {
      mbar_wait(o_empty, (tile_it & 1) ^ 1);         // the previous tile's odd pass has read the buffer
      tc_fence_after();
    }
    split_store(co, t_lane + TC_A_O + 32 * j);
    split_store(so, t_lane + TC_A_O + 32 * j + 16);
    if (j == 6) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_on_leader(o_full, rank);
    }
  }
}

// ------------------------------------------------------------------ mel side: bins of one 16-column TMEM chunk
template <int BANK, int K>
__device__ __forceinline__ void mel_bin(float (&mel)[128], float re, float im) {
  constexpr int lo = MEL_BANK_LO[BANK][K], hi = MEL_BANK_HI[BANK][K];
  if constexpr (lo >= 0 || hi >= 0) {
    const float pw = fmaf(im, im, re * re);
    if constexpr (lo >= 0) mel[lo] = fmaf(c_tc_w[BANK][K][0], pw, mel[lo]);
    if constexpr (hi >= 0) mel[hi] = fmaf(c_tc_w[BANK][K][1], pw, mel[hi]);
  }
}

template <int BANK>
__global__ void __launch_bounds__(TC_THREADS, 1) mel_tc_kernel(const __grid_constant__ CUtensorMap tm_wave, const MelTcArgs p) {
  constexpr int N_MELS = MEL_BANK_NMELS[BANK];
  constexpr int MODE = MEL_BANK_MODE[BANK];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* sB = smem + TC_OFF_B;
  float* sSlots = reinterpret_cast<float*>(smem + TC_OFF_SLOTS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_OFF_BAR);
  uint64_t* slot_full = bars;                     // [NSLOT] loader -> builders (tx bytes + 1 arrival)
  uint64_t* slot_empty = bars + TC_NSLOT;         // [NSLOT] 2 builder warps -> loader
  uint64_t* e_full = bars + 2 * TC_NSLOT;         // [2] (leader's are used) 16 builder warps of the pair -> MMA, per K block
  uint64_t* e_empty = e_full + 2;                 // [2] MMA commit (multicast) -> builders
  uint64_t* o_full = e_empty + 2;                 // (leader's) 16 builder warps -> MMA, once per tile: all 7 odd-bin blocks stored
  uint64_t* o_empty = o_full + 1;                 // MMA commit (multicast) -> builders: the odd pass has read them
  uint64_t* d_full = o_empty + 1;                 // [2] MMA commit (multicast) -> mel warps: even / odd accumulators complete
  uint64_t* d_drained = d_full + 2;               // [2] (leader's) 8 mel warps of the pair -> MMA: D may be overwritten
  uint64_t* b_full = d_drained + 2;               // twiddle image landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + TC_OFF_MISC);
  float* s_red = reinterpret_cast<float*>(smem + TC_OFF_MISC + 16);      // [2 parity][4 quarter][2 half]
  float* s_inv2 = s_red + 16;                                            // [2 parity][4 quarter]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int total_tiles = p.n_clips * TC_TILES_PER_CLIP;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_wave);
    for (int s = 0; s < TC_NSLOT; ++s) {
      mbar_init(&slot_full[s], 1);
      mbar_init(&slot_empty[s], 2);
    }
    mbar_init(&e_full[0], 16);
    mbar_init(&e_full[1], 16);
    mbar_init(&e_empty[0], 1);
    mbar_init(&e_empty[1], 1);
    mbar_init(o_full, 16);
    mbar_init(o_empty, 1);
    mbar_init(&d_full[0], 1);
    mbar_init(&d_full[1], 1);
    mbar_init(&d_drained[0], 8);
    mbar_init(&d_drained[1], 8);
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  // zero every slot once: the 4-float pads and the unused tail of the last hop are read (scale scan, the w[0] = 0
  // term) and must stay finite
  {
    float4* z = reinterpret_cast<float4*>(sSlots);
    for (int i = threadIdx.x; i < TC_NSLOT * TC_SLOT_PITCH / 4; i += TC_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async_smem();
  }
  if (warp == 2) tmem_alloc_pair<512>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Register budget: 512 threads x 128. The control warpgroup drops to 48 (frees 80 x 128), the mel warpgroup
  // (128 mel sums + a 32-register TMEM chunk per thread) grows by the same amount to 208.
  if (warp < 4) {
    setmaxnreg_dec<48>();
    if (warp == 0) {
      // ---------------------------------------------------------------- loader
      if (elect_one()) {
        mbar_arrive_expect_tx(b_full, MEL_TC_B_BYTES);
        bulk_g2s(sB, p.b_image + static_cast<size_t>(rank) * MEL_TC_B_BYTES, MEL_TC_B_BYTES, b_full);
      }
      int u = 0;
      for (int t = cluster_id; t < total_tiles; t += n_clusters) {
        const int b = t / TC_TILES_PER_CLIP;
        const int f_cta = (t % TC_TILES_PER_CLIP) * 256 + static_cast<int>(rank) * 128;
        // (never past the clip's own row, whatever n_samples says)
        const int nv = max(0, static_cast<int>(min(static_cast<long long>(p.n_samples ? min(__ldg(p.n_samples + b), TC_CLIP) : TC_CLIP),
                                                   p.wave_stride)));
        const float* w = p.wave + static_cast<long long>(b) * p.wave_stride;
        const bool base_aligned = (reinterpret_cast<uintptr_t>(w) & 15) == 0;
        for (int q = 0; q < 4; ++q, ++u) {
          const int slot = u % TC_NSLOT;
          const uint32_t use = static_cast<uint32_t>(u / TC_NSLOT);
          float* dst = sSlots + slot * TC_SLOT_PITCH;
          mbar_wait(&slot_empty[slot], (use & 1) ^ 1);
          const int f0 = f_cta + 32 * q;
          if (f0 >= TC_FRAMES) {                 // frames past the clip: never stored, any finite content will do
            if (lane == 0) mbar_arrive(&slot_full[slot]);
            continue;
          }
          const int s_first = f0 * 160 - 200;    // clip index of the slot's first sample (multiple of 8 -> 16 B steps)
          // Interior slot: rows f0 - 2 .. f0 + 31 of the view [clip][row r = samples 120 + 160 r .. + 163] in ONE tensor
          // copy; its 34 x 164 box is exactly the slot image (the 4 "pad" floats of a row are the next row's first
          // samples). Everything it reads must be real audio: the scale scan looks at the whole slot.
          if (p.use_tmap && s_first >= 0 && s_first + TC_NSEG * TC_SEG + 4 <= nv) {
            if (elect_one()) {
              mbar_arrive_expect_tx(&slot_full[slot], TC_SLOT_FLOATS * 4);
              tma_load_3d(dst, &tm_wave, &slot_full[slot], 0, f0 - 2, b);
            }
            continue;
          }
          // hop h of the slot is copied in bulk when it lies wholly inside [0, nv)
          auto hop_len = [](int h) { return h < TC_NSEG - 1 ? TC_SEG : TC_SLOT_SAMPLES - (TC_NSEG - 1) * TC_SEG; };
          auto hop_bulk = [&](int h) -> bool {
            const int s0 = s_first + h * TC_SEG;
            return base_aligned && s0 >= 0 && s0 + hop_len(h) <= nv;
          };
          const bool mine0 = hop_bulk(lane);
          const bool mine1 = lane + 32 < TC_NSEG && hop_bulk(lane + 32);
          const unsigned m0 = __ballot_sync(0xffffffffu, mine0), m1 = __ballot_sync(0xffffffffu, mine1);
          // the other hops: reflect at both clip ends (torch.stft center=True, pad_mode="reflect"), zero past nv
          if (m0 != 0xffffffffu || m1 != 3u) {
            for (int h = 0; h < TC_NSEG; ++h) {
              const bool bulk = h < 32 ? ((m0 >> h) & 1) : ((m1 >> (h - 32)) & 1);
              if (bulk) continue;
              const int len = hop_len(h);
              float v[5];
#pragma unroll
              for (int e = 0; e < 5; ++e) {
                const int o = lane + 32 * e;
                int s = s_first + h * TC_SEG + o;
                if (s < 0) s = -s;
                if (s >= TC_CLIP) s = 2 * (TC_CLIP - 1) - s;
                v[e] = (o < len && s < nv) ? __ldg(w + s) : 0.f;
              }
#pragma unroll
              for (int e = 0; e < 5; ++e) {
                const int o = lane + 32 * e;
                if (o < len) dst[h * TC_SEG_STRIDE + o] = v[e];
              }
            }
          }
          __syncwarp();
          uint32_t bytes = (__popc(m0) + __popc(m1)) * TC_SEG * 4;
          if (m1 & 2u) bytes -= (TC_NSEG * TC_SEG - TC_SLOT_SAMPLES) * 4;      // hop 33 is half a hop
          if (lane == 0) {
            if (bytes) mbar_arrive_expect_tx(&slot_full[slot], bytes);
            else mbar_arrive(&slot_full[slot]);
          }
          __syncwarp();
          if (mine0) bulk_g2s(dst + lane * TC_SEG_STRIDE, w + s_first + lane * TC_SEG, TC_SEG * 4, &slot_full[slot]);
          if (mine1)
            bulk_g2s(dst + (lane + 32) * TC_SEG_STRIDE, w + s_first + (lane + 32) * TC_SEG, hop_len(lane + 32) * 4,
                     &slot_full[slot]);
        }
      }
    } else if (warp == 1 && rank == 0) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA of the pair)
      constexpr uint32_t IDESC = umma_idesc_f16(256, TC_N);
      const uint32_t b_addr = smem_u32(sB);
      int it = 0;
      uint32_t n_use = 0;
      // the three products of one (matrix, K block): the two small ones first, the large one last
      auto products = [&](uint32_t d, uint32_t a_hi, int mat, int j) {
        const uint64_t b_hi = umma_desc_noswz(b_addr + (2 * mat) * MEL_TC_MAT_BYTES + 2 * j * MEL_TC_KCHUNK_BYTES,
                                              MEL_TC_KCHUNK_BYTES, 128);
        const uint64_t b_lo = umma_desc_noswz(b_addr + (2 * mat + 1) * MEL_TC_MAT_BYTES + 2 * j * MEL_TC_KCHUNK_BYTES,
                                              MEL_TC_KCHUNK_BYTES, 128);
        umma_ts_pair(d, a_hi + 8, b_hi, IDESC, j != 0);
        umma_ts_pair(d, a_hi, b_lo, IDESC, 1);
        umma_ts_pair(d, a_hi, b_hi, IDESC, 1);
      };
      for (int t = cluster_id; t < total_tiles; t += n_clusters, ++it) {
        // even bins: one hand-off per K block (ring of two operand stages)
        mbar_wait(&d_drained[1], (it & 1) ^ 1);     // the previous tile's odd-bin accumulators have been read
        tc_fence_after();
        for (int j = 0; j < 7; ++j, ++n_use) {
          const uint32_t es = n_use & 1;
          mbar_wait(&e_full[es], (n_use >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {   // elect.sync: a single-thread branch ptxas can see (no waterfall loop around each UTCHMMA)
            const uint32_t a0 = tmem_base + TC_A_E + 32 * es;
            products(tmem_base + TC_D_C, a0, 0, j);
            products(tmem_base + TC_D_S, a0 + 16, 1, j);
            umma_commit_pair(&e_empty[es]);
            if (j == 6) umma_commit_pair(&d_full[0]);
          }
          __syncwarp();
        }
        // odd bins: all seven K blocks are already in TMEM, 42 MMAs back to back
        mbar_wait(o_full, it & 1);
        mbar_wait(&d_drained[0], it & 1);           // the even-bin accumulators have been read
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1
          for (int j = 0; j < 7; ++j) {
            const uint32_t a0 = tmem_base + TC_A_O + 32 * j;
            products(tmem_base + TC_D_C, a0, 2, j);
            products(tmem_base + TC_D_S, a0 + 16, 3, j);
          }
          umma_commit_pair(o_empty);
          umma_commit_pair(&d_full[1]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 12) {
    // ------------------------------------------------------------------ operand builders
    const int q = warp & 3;                        // TMEM lane quarter = 32-frame slot of the tile
    const int hh = (warp - 4) >> 2;                // which 8 of each block's 16 K columns
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    mbar_wait(b_full, 0);                          // (the first a_full arrival then implies this CTA's twiddles are in)
    int it = 0;
    for (int t = cluster_id; t < total_tiles; t += n_clusters, ++it) {
      const int u = it * 4 + q;
      const int slot = u % TC_NSLOT;
      const uint32_t use = static_cast<uint32_t>(u / TC_NSLOT);
      const float* sl = sSlots + slot * TC_SLOT_PITCH;
      mbar_wait(&slot_full[slot], use & 1);
      // power-of-two scale of this slot: 4 max|x| 2^sh < 2^15 (a folded value is a sum of four samples)
      float mx = 0.f;
      {
        const uint32_t s4 = smem_u32(sl);
        for (int i = hh * 32 + lane; i < TC_SLOT_FLOATS / 4; i += 64) {
          const float4 v = lds_f32x4(s4 + 16 * i);
          mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        }
        mx = warp_max(mx);
        float* red = s_red + (it & 1) * 8 + q * 2;
        if (lane == 0) red[hh] = mx;
        named_bar_sync(1 + q, 64);
        mx = fmaxf(red[0], red[1]);
      }
      int sh = 0;
      if (mx > 0.f) {
        const int ex = static_cast<int>((__float_as_uint(mx) >> 23) & 0xFF);
        sh = max(-60, min(60, 139 - ex));
      }
      const float S = __uint_as_float(static_cast<uint32_t>(127 + sh) << 23);
      if (hh == 0 && lane == 0) s_inv2[(it & 1) * 4 + q] = __uint_as_float(static_cast<uint32_t>(127 - 2 * sh) << 23);

      const uint32_t xb = smem_u32(sl + lane * TC_SEG_STRIDE);   // this thread's frame: sample n at xb + 4 tc_off(n)
      if (hh == 0) build_tile<0>(xb, S, t_lane, e_empty, e_full, o_empty, o_full, &slot_empty[slot], static_cast<uint32_t>(it), rank, lane);
      else build_tile<1>(xb, S, t_lane, e_empty, e_full, o_empty, o_full, &slot_empty[slot], static_cast<uint32_t>(it), rank, lane);
    }
  } else {
    // ------------------------------------------------------------------ power, mel, log, store
    setmaxnreg_inc<208>();
    const int q = warp & 3;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int it = 0;
    for (int t = cluster_id; t < total_tiles; t += n_clusters, ++it) {
      const int b = t / TC_TILES_PER_CLIP;
      const int frame = (t % TC_TILES_PER_CLIP) * 256 + static_cast<int>(rank) * 128 + q * 32 + lane;
      const bool valid = frame < TC_FRAMES;
      float* ocol = p.out + static_cast<long long>(b) * N_MELS * TC_FRAMES + frame;
      float mel[128];
#pragma unroll
      for (int m = 0; m < 128; ++m) mel[m] = 0.f;
      // two passes over the same accumulator columns: even bins, then (after the MMAs of the odd pass) odd bins. Each
      // pass releases D as soon as its last column has been read.
      static_for<0, 2>([&](auto pp) {
        constexpr int PAR = decltype(pp)::value;
        mbar_wait(&d_full[PAR], it & 1);
        tc_fence_after();
        static_for<0, 7>([&](auto cc) {
          constexpr int c = decltype(cc)::value;
          uint32_t rr[16], ri[16];
          tmem_ld_32x16(t_lane + TC_D_C + 16 * c, rr);
          tmem_ld_32x16(t_lane + TC_D_S + 16 * c, ri);
          tmem_ld_wait();
          if constexpr (c == 6) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_on_leader(&d_drained[PAR], rank);
          }
          static_for<0, 16>([&](auto jj) {
            constexpr int k = 2 * (16 * c + decltype(jj)::value) + PAR;
            if constexpr (k <= (PAR ? 199 : 200))
              mel_bin<BANK, k>(mel, __uint_as_float(rr[decltype(jj)::value]), __uint_as_float(ri[decltype(jj)::value]));
          });
        });
      });
      const float inv_s2 = s_inv2[(it & 1) * 4 + q];     // (written by this tile's builders before their first hand-off)
      float lmax = -INFINITY;
#pragma unroll
      for (int m = 0; m < N_MELS; ++m) {
        float v = mel[m] * inv_s2;
        if constexpr (MODE == 0) {
          // log10 via MUFU lg2 (same expression as mel.cu: zeros give exactly -10)
          v = __log2f(fmaxf(v, 1e-10f)) * 0.30102999566398120f;
          lmax = fmaxf(lmax, v);
        } else {
          v = logf(v + 1e-9f);
        }
        if (valid) ocol[static_cast<long long>(m) * TC_FRAMES] = v;
      }
      if constexpr (MODE == 0) {
        lmax = warp_max(valid ? lmax : -INFINITY);
        if (lane == 0 && lmax > -INFINITY) atomicMax(p.clip_max_bits + b, tc_float_to_ordered(lmax));
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<512>(tmem_base);
}

template <int BANK>
static int launch_bank(const CUtensorMap& tm, const MelTcArgs& a, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    AL_CHECK_CUDA(cudaFuncSetAttribute(mel_tc_kernel<BANK>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    attr_set = true;
  }
  const int tiles = a.n_clips * TC_TILES_PER_CLIP;
  const int pairs = num_sms / 2;
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TC_SMEM;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mel_tc_kernel<BANK>, tm, a));
  return 0;
}

int launch_mel_tc(const float* wave, const int* n_samples, int B, long long wave_stride, const MelTables& tb, int mode,
                  float* out, unsigned int* clip_max_bits, int num_sms, cudaStream_t stream) {
  AL_REQUIRE(tb.tc_bank >= 0 && tb.tc_bank < MEL_N_BANKS && MEL_BANK_MODE[tb.tc_bank] == mode,
             "launch_mel_tc: no compiled bank structure for n_mels=%d mode=%d", tb.n_mels, mode);
  if (mode == 0) AL_CHECK_CUDA(cudaMemsetAsync(clip_max_bits, 0, sizeof(unsigned int) * B, stream));
  MelTcArgs a;
  a.wave = wave;
  a.n_samples = n_samples;
  a.wave_stride = wave_stride;
  a.b_image = tb.tc_b_image;
  a.out = out;
  a.clip_max_bits = clip_max_bits;
  a.n_clips = B;
  // Overlapping-row view of the wave buffer for the slot loader: [clip][row r][164] with row r = samples
  // 120 + 160 r .. 120 + 160 r + 163 of the clip (row stride 640 B). Needs a 16-byte aligned base and clip stride.
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  a.use_tmap = 0;
  if ((reinterpret_cast<uintptr_t>(wave) & 15) == 0 && wave_stride % 4 == 0 && wave_stride >= 120 + 164) {
    const uint64_t rows = static_cast<uint64_t>((wave_stride - 120 - 164) / 160 + 1);
    const uint64_t dims[3] = {164, rows, static_cast<uint64_t>(B)};
    const uint64_t str[3] = {4, 640, static_cast<uint64_t>(wave_stride) * 4};
    const uint32_t box[3] = {164, TC_NSEG, 1};
    if (make_tmap(&tm, wave + 120, 4, 3, dims, str, box, false) != 0) return -1;
    a.use_tmap = 1;
  }
  switch (tb.tc_bank) {
    case 0: return launch_bank<0>(tm, a, num_sms, stream);
    case 1: return launch_bank<1>(tm, a, num_sms, stream);
    default: return launch_bank<2>(tm, a, num_sms, stream);
  }
}

}  // namespace al
