// Non-causal, unmasked multi-head attention forward for head_dim 64 on sm_100a (tcgen05 / TMEM / TMA).
//
// Replaces HF WhisperAttention's softmax(Q K^T) V (modeling_whisper.py:215-238 via :339-349; scaling = 1.0
// because q_proj's output is pre-scaled, :310) on the fused qkv buffer the QKV GEMM writes:
//   qkv [B][T][3*H*64] bf16  (q | k | v, head h at columns h*64..h*64+63 of each third)
//   out [B][T][H*64]   bf16
//
// One CTA = 2 query tiles of 128 rows of one (batch, head). 12 warps:
//   warp 0   TMA producer: Q tiles once, then K/V tiles through a 3-stage ring
//   warp 1   MMA issuer:   S_i = Q_i K_j^T (SS, fp32 in TMEM), O_i = P_i V_j (A = P from TMEM, B = V MN-major smem)
//   warp 2   TMEM allocator (512 columns: S0 S1 | O0 O1 | P0 P1)
//   warps 4-7 / 8-11  softmax warpgroup for query tile 0 / 1: one thread per query row, online softmax in fp32,
//            P written back to TMEM as bf16, per-tile O read back and accumulated in registers with the running
//            rescale (so O in TMEM never needs a correction pass).
// The two warpgroups ping-pong on the MUFU (exp2) while the other tile's MMAs run.
#include "common.cuh"
#include "kernels.h"

namespace al {

constexpr int ATT_BQ = 128;       // query rows per tile
constexpr int ATT_BKV = 128;      // kv rows per tile
constexpr int ATT_HD = 64;
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;   // 16 KB: any of Q / K / V tile
constexpr int ATT_SMEM = 2 * ATT_TILE_BYTES + ATT_KV_STAGES * 2 * ATT_TILE_BYTES + 256 + 1024;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(384, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, int T, int H) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 2 tiles
  uint8_t* sKV = smem + 2 * ATT_TILE_BYTES;             // stages x {K, V}
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + ATT_KV_STAGES * 2 * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                              // 1
  uint64_t* kv_full = bars + 1;                         // KV_STAGES
  uint64_t* kv_empty = kv_full + ATT_KV_STAGES;         // KV_STAGES
  uint64_t* s_full = kv_empty + ATT_KV_STAGES;          // 2
  uint64_t* s_empty = s_full + 2;                       // 2
  uint64_t* p_full = s_empty + 2;                       // 2
  uint64_t* o_full = p_full + 2;                        // 2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int d = H * ATT_HD;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int q0 = blockIdx.x * 2 * ATT_BQ;
  const int nkv = (T + ATT_BKV - 1) / ATT_BKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tS = tmem_base;            // + i*128
  const uint32_t tO = tmem_base + 256;      // + i*64
  const uint32_t tP = tmem_base + 384;      // + i*64 (bf16 pairs: 128 kv -> 64 columns)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      tma_load_3d(sQ, &tmQKV, q_full, h * ATT_HD, q0, b);
      tma_load_3d(sQ + ATT_TILE_BYTES, &tmQKV, q_full, h * ATT_HD, q0 + ATT_BQ, b);
      int s = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * ATT_TILE_BYTES);
        uint8_t* kdst = sKV + s * 2 * ATT_TILE_BYTES;
        tma_load_3d(kdst, &tmQKV, &kv_full[s], d + h * ATT_HD, j * ATT_BKV, b);
        tma_load_3d(kdst + ATT_TILE_BYTES, &tmQKV, &kv_full[s], 2 * d + h * ATT_HD, j * ATT_BKV, b);
        if (++s == ATT_KV_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t IDESC_S = umma_idesc_bf16(ATT_BQ, ATT_BKV);             // Q K^T: both K-major
    constexpr uint32_t IDESC_O = umma_idesc_bf16(ATT_BQ, ATT_HD, 0, 1);        // P V: V is MN-major
    auto issue_s = [&](int i, int stage) {
      const uint64_t qd = umma_desc_sw128(smem_u32(sQ + i * ATT_TILE_BYTES), 16, 1024);
      const uint64_t kd = umma_desc_sw128(smem_u32(sKV + stage * 2 * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
      for (int k = 0; k < ATT_HD / 16; ++k) umma_ss(tS + i * 128, qd + 2 * k, kd + 2 * k, IDESC_S, k != 0);
      umma_commit(&s_full[i]);
    };
    auto issue_o = [&](int i, int stage) {
      // V tile: [kv 128 rows][64 d] bf16, 128 B rows, SW128 -> MN-major B operand. One UMMA_K = 16 kv rows = 2048 B.
      const uint32_t vbase = smem_u32(sKV + stage * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < ATT_BKV / 16; ++k) {
        const uint64_t vd = umma_desc_sw128(vbase + k * 2048, 1024, 1024);
        umma_ts(tO + i * 64, tP + i * 64 + k * 8, vd, IDESC_O, k != 0);
      }
      umma_commit(&o_full[i]);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    if (lane == 0) {
      issue_s(0, 0);
      issue_s(1, 0);
    }
    __syncwarp();
    int s = 0;
    uint32_t ph = 0;
    for (int j = 0; j < nkv; ++j) {
      int sn = s + 1;
      uint32_t phn = ph;
      if (sn == ATT_KV_STAGES) { sn = 0; phn ^= 1; }
      const bool has_next = (j + 1 < nkv);
      if (has_next) mbar_wait(&kv_full[sn], phn);
      for (int i = 0; i < 2; ++i) {
        if (has_next) {
          mbar_wait(&s_empty[i], j & 1);
          tc_fence_after();
          if (lane == 0) issue_s(i, sn);
          __syncwarp();
        }
        mbar_wait(&p_full[i], j & 1);
        tc_fence_after();
        if (lane == 0) issue_o(i, s);
        __syncwarp();
      }
      if (lane == 0) umma_commit(&kv_empty[s]);
      __syncwarp();
      s = sn;
      ph = phn;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax warpgroups
    const int i = (warp - 4) >> 2;                 // query tile of this warpgroup
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t tSi = tS + i * 128 + lane_off;
    const uint32_t tOi = tO + i * 64 + lane_off;
    const uint32_t tPi = tP + i * 64 + lane_off;
    float m = -INFINITY, l = 0.f;
    float acc[ATT_HD];
#pragma unroll
    for (int c = 0; c < ATT_HD; ++c) acc[c] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      const int kv_valid = min(ATT_BKV, T - j * ATT_BKV);   // columns >= kv_valid are padding (zero K rows)
      mbar_wait(&s_full[i], j & 1);
      tc_fence_after();
      // pass 1: row max
      float mx = m;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tSi + c * 32, r);
        tmem_ld_wait();
        if (c * 32 + 32 <= kv_valid) {
#pragma unroll
          for (int k = 0; k < 32; ++k) mx = fmaxf(mx, __uint_as_float(r[k]));
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (c * 32 + k < kv_valid) mx = fmaxf(mx, __uint_as_float(r[k]));
        }
      }
      const float scale = fast_exp2((m - mx) * LOG2E);    // first tile: exp2(-inf) = 0
      // fold in the previous tile's P V (relative to the old max), then rescale to the new max
      if (j > 0) {
        mbar_wait(&o_full[i], (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tOi + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; ++k) acc[c * 32 + k] = (acc[c * 32 + k] + __uint_as_float(r[k])) * scale;
        }
      }
      l *= scale;
      m = mx;
      const float msc = mx * LOG2E;
      // pass 2: p = exp2(s*log2e - m*log2e), row sum, P -> TMEM as bf16
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tSi + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          float p0 = fast_exp2(fmaf(__uint_as_float(r[k]), LOG2E, -msc));
          float p1 = fast_exp2(fmaf(__uint_as_float(r[k + 1]), LOG2E, -msc));
          if (c * 32 + k >= kv_valid) p0 = 0.f;
          if (c * 32 + k + 1 >= kv_valid) p1 = 0.f;
          l += p0 + p1;
          pk[k >> 1] = pack_bf16(p0, p1);
        }
        tmem_st_32x16(tPi + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&s_empty[i]);
      mbar_arrive(&p_full[i]);
    }
    // last tile's P V
    mbar_wait(&o_full[i], (nkv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const int q = q0 + i * ATT_BQ + row;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tOi + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[c * 32 + k] = (acc[c * 32 + k] + __uint_as_float(r[k])) * inv_l;
    }
    if (q < T) {
      uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * T + q) * d + h * ATT_HD);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint4 v;
        v.x = pack_bf16(acc[8 * u], acc[8 * u + 1]);
        v.y = pack_bf16(acc[8 * u + 2], acc[8 * u + 3]);
        v.z = pack_bf16(acc[8 * u + 4], acc[8 * u + 5]);
        v.w = pack_bf16(acc[8 * u + 6], acc[8 * u + 7]);
        dst[u] = v;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

int launch_attention(const CUtensorMap& tm_qkv, void* out, int B, int T, int H, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    AL_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  dim3 grid((T + 2 * ATT_BQ - 1) / (2 * ATT_BQ), H, B);
  attention_fwd_kernel<<<grid, 384, ATT_SMEM, stream>>>(tm_qkv, reinterpret_cast<__nv_bfloat16*>(out), T, H);
  AL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace al
