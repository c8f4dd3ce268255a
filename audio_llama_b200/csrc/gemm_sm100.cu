// Persistent, warp-specialised bf16 GEMM for sm_100a: out = epilogue(A @ W^T + bias).
//
//   A   : [batch][m_per_batch][K] bf16, K contiguous, arbitrary (16 B aligned) row / batch strides —
//         the strides are what turn Whisper's conv stem into a plain GEMM (rows overlap: conv1 row stride is
//         Cin, row length 3*Cin; conv2 row stride is 2*d, row length 3*d), see api.cu.
//   W   : [N][K] bf16 (nn.Linear weight layout), K contiguous.
//   out : [batch][m_per_batch][N] bf16 or fp32.
//
// Two forms, same code path (template flag CTA2): one CTA per 128x256 tile, or a CTA PAIR per 256x256 tile with
// tcgen05.mma.cta_group::2 (default; AUDIOLLM_B200_GEMM=single selects the former).
// Roles (384 threads): warp 0 = TMA producer (one lane), warp 1 = tcgen05.mma issuer (one lane),
// warp 2 = TMEM allocator, warps 4-7 and 8-11 = two epilogue warpgroups that take alternate column chunks of
// the tile (TMEM -> registers -> bias / GELU / residual -> swizzled smem -> TMA store). Operands are staged by TMA into 128B-swizzled K-major tiles; the fp32 accumulator lives
// in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// M / N / K tails need no code: TMA zero-fills out-of-bounds loads and clips out-of-bounds stores.
//
// Replaces, on the reference's path: every nn.Linear / Conv1d of HF WhisperEncoder
// (modeling_whisper.py:279-282, 310, 404-406, 567-568, 619-620) and of AudioProjector
// (/root/reference/src/models/projector.py:11-16).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace al {

constexpr int BM = 128;            // rows per CTA
constexpr int BK = 64;
constexpr int STG_BYTES = 16384;   // one epilogue staging buffer: 128 rows x 128 B

// CTA2 = false: one CTA per 128 x 256 tile (A 16 KB + B 32 KB per stage, 4 stages).
// CTA2 = true : a CTA PAIR (cluster of 2, one tcgen05.mma.cta_group::2 per K step) per 256 x 256 tile. Each CTA
//               stages its own 128 rows of A and HALF of B (128 of the 256 weight rows): 32 KB per stage, 6 stages.
//               Per SM that is 2/3 of the shared-memory fill + operand-read traffic of the single-CTA form for
//               the same MMA rate (ncu showed the single-CTA kernel at ~80 % tensor-active with shared memory
//               the busiest unit), and half the L2 -> SM weight traffic.
template <bool CTA2>
__host__ __device__ constexpr int gemm_stages() { return CTA2 ? 6 : 4; }
template <bool CTA2>
__host__ __device__ constexpr int gemm_smem_bytes() {
  return gemm_stages<CTA2>() * (BM * BK * 2 + (CTA2 ? 128 : 256) * BK * 2) + 2 * STG_BYTES + 256 /*barriers*/ +
         1024 /*alignment slack*/;
}

template <int FLAGS, bool CTA2>
__global__ void __launch_bounds__(384, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmAdd, const GemmParams p) {
  constexpr int BN = 256;
  constexpr int STAGES = gemm_stages<CTA2>();
  constexpr bool OUT_F32 = (FLAGS & EPI_OUT_F32) != 0;
  constexpr bool DO_GELU = (FLAGS & EPI_GELU) != 0;
  constexpr bool REDUCE = (FLAGS & EPI_REDUCE_ADD) != 0;
  constexpr bool ROWAUX = (FLAGS & EPI_ROWAUX) != 0;
  constexpr bool RESID = (FLAGS & EPI_RESIDUAL) != 0;
  constexpr bool GELU_GRAD = (FLAGS & EPI_GELU_GRAD) != 0;
  constexpr bool ADD16 = (FLAGS & EPI_ADD_BF16) != 0;
  constexpr bool TN = (FLAGS & EPI_TN) != 0;     // operands row-major over the contraction: staged MN-major
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_ROWS = CTA2 ? 128 : 256;       // weight rows staged by this CTA
  constexpr int B_BYTES = B_ROWS * BK * 2;
  constexpr int CH = OUT_F32 ? 32 : 64;          // output columns per staging chunk (128 B per row)
  constexpr int NCHUNK = BN / CH;
  constexpr uint32_t IDESC = umma_idesc_bf16(CTA2 ? 256 : 128, BN, TN ? 1 : 0, TN ? 1 : 0);
  constexpr int TILE_M = CTA2 ? 256 : 128;       // rows per scheduled tile

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint8_t* sStg = sB + STAGES * B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sStg + 2 * STG_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* addfull = tempty + 2;                // EPI_ADD_BF16: "addend chunk has landed", one per epilogue warpgroup
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(addfull + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0;        // 0 = leader (issues the MMAs)
  const int worker = CTA2 ? (blockIdx.x >> 1) : blockIdx.x;  // tile-scheduler slot
  const int n_workers = CTA2 ? (gridDim.x >> 1) : gridDim.x;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);                    // (pair: only the leader's is used; both CTAs' TMA bytes land on it)
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&addfull[a], 1);
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], CTA2 ? 512 : 256);   // (pair: the leader's collects both CTAs' epilogue threads)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CTA2) tmem_alloc_pair<2 * BN>(tmem_ptr);
    else tmem_alloc<2 * BN>(tmem_ptr);
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_m = (p.m_per_batch + TILE_M - 1) / TILE_M;
  const int out_tiles = p.batch * tiles_m * p.tiles_n;
  // split-K (weight-gradient GEMMs: small output, very long contraction): work item = (output tile, K slice);
  // every slice reduce-adds its partial tile into the fp32 output, slice 0 also adds the bias.
  const int k_splits = p.k_splits > 1 ? p.k_splits : 1;
  const int num_tiles = out_tiles * k_splits;
  // K blocks of the main product, then (LoRA) K2 more blocks of a second operand pair accumulated into the same
  // TMEM accumulator: out = A W^T + A2 W2^T, with A2 = x A_lora^T [M, r] and W2 = scaling * B_lora [N, r].
  const int num_kb1 = (p.K + BK - 1) / BK;
  const int num_kb_all = num_kb1 + (p.K2 + BK - 1) / BK;
  const int kb_per_split = (num_kb_all + k_splits - 1) / k_splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one per CTA)
    // elect.sync (not lane == 0): ptxas then knows a single thread runs this and emits no per-lane "waterfall" loop
    // (ELECT / R2UR.BROADCAST / BRA.U.ANY) around every uniform-datapath instruction (UTMALDG, UTCHMMA, UTCBAR)
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = worker; t < num_tiles; t += n_workers) {
        const int ot = t % out_tiles, split = t / out_tiles;
        const int nt = ot % p.tiles_n;
        const int mt = ot / p.tiles_n;
        const int m0 = (mt % tiles_m) * TILE_M + rank * BM;
        const int b = mt / tiles_m;
        const int n0 = nt * BN + rank * B_ROWS;                  // pair: this CTA's half of the weight rows
        const int kb_begin = split * kb_per_split;
        const int kb_end = min(kb_begin + kb_per_split, num_kb_all);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          const bool second = kb >= num_kb1;
          const int k0 = (second ? kb - num_kb1 : kb) * BK;
          const CUtensorMap* ma = second ? &tmA2 : &tmA;
          const CUtensorMap* mb = second ? &tmB2 : &tmB;
          if constexpr (TN) {
            // MN-major staging: 64-wide column atoms of the row-major sources, [64 k rows][64 cols] = 8 KB each
            if constexpr (CTA2) {
              if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * (A_BYTES + B_BYTES));
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_2d_pair(sA + s * A_BYTES + i * 8192, ma, &full[s], m0 + 64 * i, k0);
#pragma unroll
              for (int i = 0; i < B_ROWS / 64; ++i) tma_load_2d_pair(sB + s * B_BYTES + i * 8192, mb, &full[s], n0 + 64 * i, k0);
            } else {
              mbar_arrive_expect_tx(&full[s], A_BYTES + B_BYTES);
#pragma unroll
              for (int i = 0; i < BM / 64; ++i) tma_load_2d(sA + s * A_BYTES + i * 8192, ma, &full[s], m0 + 64 * i, k0);
#pragma unroll
              for (int i = 0; i < B_ROWS / 64; ++i) tma_load_2d(sB + s * B_BYTES + i * 8192, mb, &full[s], n0 + 64 * i, k0);
            }
          } else if constexpr (CTA2) {
            if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * (A_BYTES + B_BYTES));
            tma_load_3d_pair(sA + s * A_BYTES, ma, &full[s], k0, m0, b);
            tma_load_2d_pair(sB + s * B_BYTES, mb, &full[s], k0, n0);
          } else {
            mbar_arrive_expect_tx(&full[s], A_BYTES + B_BYTES);
            tma_load_3d(sA + s * A_BYTES, ma, &full[s], k0, m0, b);
            tma_load_2d(sB + s * B_BYTES, mb, &full[s], k0, n0);
            tma_load_2d(sB + s * B_BYTES + B_BYTES / 2, mb, &full[s], k0, n0 + 128);
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    int s = 0, as = 0;
    uint32_t ph = 0, aph = 0;
    for (int t = worker; t < num_tiles; t += n_workers) {
      mbar_wait(&tempty[as], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      const int split = t / out_tiles;
      const int num_kb = min(kb_per_split, num_kb_all - split * kb_per_split);   // >= 1 (host guarantees)
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          // K-major: +32 B per UMMA_K inside the 128 B swizzle atom = +2 in the >>4 field.
          // MN-major (TN): atoms of 64 columns 8192 B apart (LBO), 8-row groups 1024 B apart (SBO); one UMMA_K = 16 rows
          // = 2048 B = +128.
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * A_BYTES), TN ? 8192 : 16, 1024);
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * B_BYTES), TN ? 8192 : 16, 1024);
          constexpr int KSTEP = TN ? 128 : 2;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if constexpr (CTA2) umma_ss_pair(d_tmem, adesc + KSTEP * k, bdesc + KSTEP * k, IDESC, (kb | k) != 0);
            else umma_ss(d_tmem, adesc + KSTEP * k, bdesc + KSTEP * k, IDESC, (kb | k) != 0);
          }
          if constexpr (CTA2) {
            umma_commit_pair(&empty[s]);
            if (kb == num_kb - 1) umma_commit_pair(&tfull[as]);
          } else {
            umma_commit(&empty[s]);
            if (kb == num_kb - 1) umma_commit(&tfull[as]);
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (++as == 2) { as = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (2 warpgroups per CTA)
    const int wg = (warp - 4) >> 2;               // chunks c with c % 2 == wg
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;          // row inside this CTA's 128 rows
    const int etid = threadIdx.x - 128 - wg * 128;
    uint8_t* stg = sStg + wg * STG_BYTES;         // one staging buffer per warpgroup
    uint8_t* rowp = stg + row * 128;
    int as = 0;
    uint32_t aph = 0;
    uint32_t addph = 0;                           // phase of this warpgroup's addend barrier
    if constexpr (ADD16) {
      if (etid == 0) tma_prefetch_desc(&tmAdd);
    }
    for (int t = worker; t < num_tiles; t += n_workers) {
      const int ot = t % out_tiles, split = t / out_tiles;
      const int nt = ot % p.tiles_n;
      const int mt = ot / p.tiles_n;
      const int m0 = (mt % tiles_m) * TILE_M + rank * BM;
      const int b = mt / tiles_m;
      const int n0 = nt * BN;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + as * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      const int grow = min(m0 + row, p.m_per_batch - 1);     // clamped row for aux / residual reads
#pragma unroll 1
      for (int c = wg; c < NCHUNK; c += 2) {
        float v[CH];
        {
          uint32_t r0[32];
          tmem_ld_32x32(t_row + c * CH, r0);
          if constexpr (CH == 64) {
            uint32_t r1[32];
            tmem_ld_32x32(t_row + c * CH + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[32 + j] = __uint_as_float(r1[j]);
          } else {
            tmem_ld_wait();
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
        }
        const int col0 = n0 + c * CH;
        const bool fullc = col0 + CH <= p.N;
        if (p.bias != nullptr && split == 0) {
          if (fullc) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] += __ldg(p.bias + min(col0 + j, p.N - 1));
          }
        }
        if constexpr (DO_GELU) {
#pragma unroll
          for (int j = 0; j < CH; j += 2) gelu_erf_fast2(v[j], v[j + 1]);
        }
        if constexpr (GELU_GRAD) {
          // backward of h = gelu(a): v holds the recomputed pre-activation a = x W1^T + b1; out = dh * gelu'(a),
          // gelu'(a) = Phi(a) + a phi(a). dh is a bf16 [M][N] matrix read row-wise (128 B per thread).
          const __nv_bfloat16* gp = p.grad_in + (static_cast<size_t>(b) * p.m_per_batch + grow) * p.grad_ld + col0;
#pragma unroll
          for (int j = 0; j < CH; j += 8) {
            uint4 g = make_uint4(0, 0, 0, 0);
            if (fullc) g = __ldg(reinterpret_cast<const uint4*>(gp + j));
            const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float g0 = __uint_as_float(gw[q] << 16), g1 = __uint_as_float(gw[q] & 0xffff0000u);
              if (!fullc) {
                g0 = (col0 + j + 2 * q < p.N) ? __bfloat162float(gp[j + 2 * q]) : 0.f;
                g1 = (col0 + j + 2 * q + 1 < p.N) ? __bfloat162float(gp[j + 2 * q + 1]) : 0.f;
              }
              const float a0 = v[j + 2 * q], a1 = v[j + 2 * q + 1];
              const float cdf0 = 0.5f * (1.0f + erff(a0 * 0.70710678118654752440f));
              const float cdf1 = 0.5f * (1.0f + erff(a1 * 0.70710678118654752440f));
              const float pdf0 = 0.3989422804014327f * __expf(-0.5f * a0 * a0);
              const float pdf1 = 0.3989422804014327f * __expf(-0.5f * a1 * a1);
              v[j + 2 * q] = g0 * (cdf0 + a0 * pdf0);
              v[j + 2 * q + 1] = g1 * (cdf1 + a1 * pdf1);
            }
          }
        }
        if constexpr (ROWAUX) {
          const float* ap = p.aux + static_cast<size_t>(grow) * p.aux_ld;
          if (fullc) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              const float4 av = __ldg(reinterpret_cast<const float4*>(ap + col0 + j));
              v[j] += av.x; v[j + 1] += av.y; v[j + 2] += av.z; v[j + 3] += av.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] += __ldg(ap + min(col0 + j, p.N - 1));
          }
        }
        if constexpr (RESID) {
          // residual stream, possibly the output buffer itself: this CTA is the only reader and writer of this
          // tile, and the read happens before its own TMA store -> plain (coherent) loads, no atomics
          const float* rp = p.resid + (static_cast<size_t>(b) * p.resid_batch_stride + static_cast<size_t>(grow) * p.resid_ld);
          if (fullc) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              const float4 rv = *reinterpret_cast<const float4*>(rp + col0 + j);
              v[j] += rv.x; v[j + 1] += rv.y; v[j + 2] += rv.z; v[j + 3] += rv.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] += rp[min(col0 + j, p.N - 1)];
          }
        }
        // stage into 128B-swizzled smem (row = 128 B; 16 B unit u of row r lives at u ^ (r & 7))
        // (bulk async-groups belong to the issuing thread: elect.sync with the full mask always elects the same lane)
        if constexpr (ADD16) {
          // + a bf16 addend with the output's shape (possibly the output buffer itself): its [128 rows][64 columns] chunk
          // is brought by TMA into this warpgroup's staging buffer (coalesced, asynchronous -- read straight from the
          // threads, each lane's 128 bytes of a different row cost 5 - 17 % of the GEMM), added in place row by row (a
          // thread reads and writes only its own row), and the same buffer then leaves through the TMA store below.
          const bool rows_live = m0 < p.m_per_batch;                // (pair: the peer's rows may lie wholly past the end)
          if (rows_live) {
            if (etid < 32 && elect_one()) {
              tma_wait_group_read<0>();                             // this warpgroup's previous store has drained the buffer
              mbar_arrive_expect_tx(&addfull[wg], STG_BYTES);
              tma_load_3d(stg, &tmAdd, &addfull[wg], col0, m0, b);
            }
            mbar_wait(&addfull[wg], addph);
            addph ^= 1;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint4 g = *reinterpret_cast<const uint4*>(rowp + ((u ^ (row & 7)) << 4));
              const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[8 * u + 2 * q] += __uint_as_float(gw[q] << 16);
                v[8 * u + 2 * q + 1] += __uint_as_float(gw[q] & 0xffff0000u);
              }
            }
          } else {
            if (etid < 32 && elect_one()) tma_wait_group_read<0>();
            named_bar_sync(1 + 2 * wg, 128);
          }
        } else {
          if (etid < 32 && elect_one()) tma_wait_group_read<0>();   // this warpgroup's previous store has drained the buffer
          named_bar_sync(1 + 2 * wg, 128);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint4 q;
          if constexpr (OUT_F32) {
            q.x = __float_as_uint(v[4 * u]); q.y = __float_as_uint(v[4 * u + 1]);
            q.z = __float_as_uint(v[4 * u + 2]); q.w = __float_as_uint(v[4 * u + 3]);
          } else {
            q.x = pack_bf16(v[8 * u], v[8 * u + 1]); q.y = pack_bf16(v[8 * u + 2], v[8 * u + 3]);
            q.z = pack_bf16(v[8 * u + 4], v[8 * u + 5]); q.w = pack_bf16(v[8 * u + 6], v[8 * u + 7]);
          }
          *reinterpret_cast<uint4*>(rowp + ((u ^ (row & 7)) << 4)) = q;
        }
        fence_proxy_async_smem();
        named_bar_sync(2 + 2 * wg, 128);
        if (etid < 32 && m0 < p.m_per_batch && elect_one()) {     // (pair: the peer's rows may lie wholly past the end)
          if constexpr (REDUCE) tma_reduce_add_3d(&tmO, stg, col0, m0, b);
          else tma_store_3d(&tmO, stg, col0, m0, b);
          tma_commit_group();
        }
      }
      tc_fence_before();
      if constexpr (CTA2) mbar_arrive_cluster(&tempty[as], 0);   // accumulator stage drained: tell the leader's MMA warp
      else mbar_arrive(&tempty[as]);
      if (++as == 2) { as = 0; aph ^= 1; }
    }
    if (etid < 32 && elect_one()) tma_wait_group<0>();
  }

  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();   // pair: the peer's smem / barriers stay alive until both are done
  if (warp == 2) {
    if constexpr (CTA2) tmem_dealloc_pair<2 * BN>(tmem_base);
    else tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ----------------------------------------------------------------------------- host launch
static int g_use_pair = -1;   // -1 = read AUDIOLLM_B200_GEMM (pair | single) once; default pair

static bool use_pair() {
  if (g_use_pair < 0) {
    const char* e = getenv("AUDIOLLM_B200_GEMM");
    g_use_pair = (e && strcmp(e, "single") == 0) ? 0 : 1;
  }
  return g_use_pair == 1;
}
void gemm_set_mode(int pair) { g_use_pair = pair ? 1 : 0; }

template <int FLAGS, bool CTA2>
static int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmA2,
                      const CUtensorMap& tmB2, const CUtensorMap& tmAdd, const GemmParams& p, int num_sms, cudaStream_t stream) {
  constexpr int smem = gemm_smem_bytes<CTA2>();
  static bool attr_set = false;
  auto kern = gemm_bf16_kernel<FLAGS, CTA2>;
  if (!attr_set) {
    AL_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int tile_m = CTA2 ? 256 : 128;
  const int tiles = p.batch * ((p.m_per_batch + tile_m - 1) / tile_m) * p.tiles_n * (p.k_splits > 1 ? p.k_splits : 1);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  if (CTA2) {
    const int pairs = num_sms / 2;
    cfg.gridDim = dim3(2 * (tiles < pairs ? tiles : pairs));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3(tiles < num_sms ? tiles : num_sms);
  }
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  AL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO, tmA2, tmB2, tmAdd, p));
  return 0;
}

int gemm_out_box_cols(int flags) { return (flags & EPI_OUT_F32) ? 32 : 64; }   // (EPI_TN does not change the output)

int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, GemmParams p, int flags,
                int num_sms, cudaStream_t stream) {
  p.K2 = 0;
  return launch_gemm2(tmA, tmB, tmO, tmA, tmB, p, flags, num_sms, stream);
}

template <int FLAGS>
static int launch_mode(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmA2,
                       const CUtensorMap& tmB2, const CUtensorMap& tmAdd, const GemmParams& p, int num_sms, cudaStream_t stream) {
  if (use_pair()) return launch_one<FLAGS, true>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
  return launch_one<FLAGS, false>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
}

int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmA2,
                 const CUtensorMap& tmB2, GemmParams p, int flags, int num_sms, cudaStream_t stream) {
  return launch_gemm2_add(tmA, tmB, tmO, tmA2, tmB2, tmO, p, flags, num_sms, stream);
}

// tmAdd: the EPI_ADD_BF16 addend, a bf16 matrix with the output's shape and the output map's box (it may BE the output)
int launch_gemm2_add(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmA2,
                     const CUtensorMap& tmB2, const CUtensorMap& tmAdd, GemmParams p, int flags, int num_sms, cudaStream_t stream) {
  p.tiles_m_per_batch = 0;   // (derived in the kernel from the tile height of the chosen form)
  p.tiles_n = (p.N + 255) / 256;
  switch (flags) {
    case 0: return launch_mode<0>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
    case EPI_GELU: return launch_mode<EPI_GELU>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
    case EPI_OUT_F32: return launch_mode<EPI_OUT_F32>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
    case EPI_OUT_F32 | EPI_REDUCE_ADD:
      return launch_mode<EPI_OUT_F32 | EPI_REDUCE_ADD>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
    case EPI_OUT_F32 | EPI_RESIDUAL:
      return launch_mode<EPI_OUT_F32 | EPI_RESIDUAL>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
    case EPI_OUT_F32 | EPI_GELU | EPI_ROWAUX:
      return launch_mode<EPI_OUT_F32 | EPI_GELU | EPI_ROWAUX>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
    case EPI_GELU_GRAD:
      return launch_mode<EPI_GELU_GRAD>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
    case EPI_ADD_BF16:
      return launch_mode<EPI_ADD_BF16>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
    case EPI_TN | EPI_OUT_F32 | EPI_REDUCE_ADD:
      return launch_mode<EPI_TN | EPI_OUT_F32 | EPI_REDUCE_ADD>(tmA, tmB, tmO, tmA2, tmB2, tmAdd, p, num_sms, stream);
    default:
      set_error("launch_gemm: unsupported epilogue flags %d", flags);
      return -1;
  }
}

}  // namespace al
