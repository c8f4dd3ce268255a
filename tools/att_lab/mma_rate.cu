// Issue-rate probe for the tcgen05.mma forms the attention kernels use: one CTA per SM, one thread issues REP products
// of one form back to back into TMEM, commits, waits; prints cycles per product. Shared memory holds garbage (zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../audio_llama_b200/csrc mma_rate.cu -o mma_rate
#include <stdio.h>
#include "common.cuh"
using namespace al;

constexpr int REP = 64;   // x 8 products
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ void ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void ts(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 db, {%2, %5};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d), "r"(a), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI) : "memory");
}
// form 0: SS K-major/K-major N=128   1: SS N=64   2: TS (A in TMEM) B MN-major N=128   3: SS N=128 alternating two D
// 4: TS alternating two accumulators  5: SS N=64 x8 then TS x8   6: TS N=64   7: SS N=256
// 8: SS, B MN-major N=128   9: the dK/dV kernel's sub-step: TS x4 (dV), SS64 x8, TS x4 (dK), SS64 x8
template <int FORM>
__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  if (warp == 1) {
    long long t0 = 0, t1 = 0;
    constexpr uint32_t I128 = umma_idesc_bf16(128, 128), I64 = umma_idesc_bf16(128, 64), ITS = umma_idesc_bf16(128, 128, 0, 1),
                       ITS64 = umma_idesc_bf16(128, 64, 0, 1), I256 = umma_idesc_bf16(128, 256);
    const uint32_t A = ((sb & 0x3FFFF) >> 4) | (1u << 16), B = A + 2048, V = (((sb + 65536) & 0x3FFFF) >> 4) | (1024u << 16);
    if (elect_one()) {
      t0 = clock64();
      for (int r = 0; r < REP; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * 1024 + (k & 3) * 2;
          if (FORM == 0) ss(tm, A + off, B + off, I128, k != 0);
          if (FORM == 1) ss(tm, A + off, B + off, I64, k != 0);
          if (FORM == 2) ts(tm + 256, tm + k * 8, V + 128 * k, ITS, k != 0);
          if (FORM == 3) ss(tm + (r & 1 ? 128 : 0), A + off, B + off, I128, k != 0);
          if (FORM == 4) ts(tm + (r & 1 ? 384 : 256), tm + k * 8, V + 128 * k, ITS, k != 0);
          if (FORM == 5) {
            if (r & 1) ts(tm + 256, tm + k * 8, V + 128 * k, ITS, k != 0);
            else ss(tm, A + off, B + off, I64, k != 0);
          }
          if (FORM == 6) ts(tm + 256, tm + k * 8, V + 128 * k, ITS64, k != 0);
          if (FORM == 7) ss(tm, A + off, B + off, I256, k != 0);
          if (FORM == 8) ss(tm, A + off, V + 128 * k, ITS, k != 0);
        }
        if (FORM == 9) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ts(tm + 256, tm + k * 8, V + 128 * k, ITS, 1);
#pragma unroll
          for (int k = 0; k < 8; ++k) ss(tm, A + (k >> 2) * 1024 + (k & 3) * 2, B + (k >> 2) * 1024 + (k & 3) * 2, I64, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) ts(tm + 384, tm + 128 + k * 8, V + 128 * k, ITS, 1);
#pragma unroll
          for (int k = 0; k < 8; ++k) ss(tm + 128, A + (k >> 2) * 1024 + (k & 3) * 2, B + (k >> 2) * 1024 + (k & 3) * 2, I64, k != 0);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    long long mx = t0;
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));   // the elected lane's start time
    if (blockIdx.x == 0 && threadIdx.x == 32) out[FORM] = t1 - mx;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

template <int F>
void run(long long* d, int grid, const char* name, int per_rep) {
  cudaFuncSetAttribute(probe<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<F><<<grid, 128, 200 * 1024>>>(d);
  probe<F><<<grid, 128, 200 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("grid %3d  %-36s %7.1f cycles / product   (%s)\n", grid, name, double(h[F]) / (REP * per_rep), cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * 8);
  cudaMemset(d, 0, 64 * 8);
  for (int grid : {1, 148}) {
    run<0>(d, grid, "SS N=128", 8);
    run<1>(d, grid, "SS N=64", 8);
    run<2>(d, grid, "TS N=128 (B MN-major)", 8);
    run<3>(d, grid, "SS N=128, two D", 8);
    run<4>(d, grid, "TS, two D", 8);
    run<5>(d, grid, "SS N=64 x8 / TS x8 alternating", 8);
    run<6>(d, grid, "TS N=64", 8);
    run<7>(d, grid, "SS N=256", 8);
    run<8>(d, grid, "SS N=128, B MN-major", 8);
    run<9>(d, grid, "dK/dV sub-step (24 products)", 24);
  }
  return 0;
}
