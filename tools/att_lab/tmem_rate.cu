// TMEM read-port probe: NW warps each loop tcgen05.ld 32x32b.x32 (4 KB per warp-instruction), optionally while one thread
// streams SS / TS products; prints bytes per cycle per SM for the loads and cycles per product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../audio_llama_b200/csrc tmem_rate.cu -o tmem_rate
// WARNING kept from the first version of this probe: it consumed the loaded values through a run-time register index
// (`acc ^= v[r & 31]`), which put the destination array in LOCAL memory — every load was followed by 32 spill stores and
// the 'TMEM port' read 50 B / cycle / SM. Check `-Xptxas -v` for a zero stack frame before believing a number from here.
#include <stdio.h>
#include "common.cuh"
using namespace al;

constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ void ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void ts(uint32_t d, uint32_t a, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 db, {%2, %5};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d), "r"(a), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI) : "memory");
}

// MODE 0: loads only  1: SS products only  2: both  3: TS products + loads  4: stores (32x32) only  5: SS + stores
template <int MODE, int NW, int LOADS, int REP>
__global__ void __launch_bounds__(64 + 32 * NW, 1) probe(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
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
    if (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 5) {
      long long t0 = 0, t1 = 0;
      constexpr uint32_t I128 = umma_idesc_bf16(128, 128), ITS = umma_idesc_bf16(128, 128, 0, 1);
      const uint32_t A = ((sb & 0x3FFFF) >> 4) | (1u << 16), B = A + 2048, V = (((sb + 65536) & 0x3FFFF) >> 4) | (1024u << 16);
      if (elect_one()) {
        t0 = clock64();
        for (int r = 0; r < REP; ++r) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t off = (k >> 2) * 1024 + (k & 3) * 2;
            if (MODE == 3) ts(tm + 256, tm + 384 + k * 8, V + 128 * k, ITS, k != 0);
            else ss(tm + 256, A + off, B + off, I128, k != 0);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, 0);
      t1 = clock64();
      long long mx = t0;
      for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (blockIdx.x == 0 && threadIdx.x == 32) out[0] = t1 - mx;
    }
  } else if (warp >= 2) {
    if (MODE == 0 || MODE == 2 || MODE == 3) {
      const uint32_t tl = tm + (((warp & 3) * 32u) << 16);
      uint32_t acc = 0;
      const long long t0 = clock64();
      for (int r = 0; r < LOADS; ++r) {
        uint32_t v[32];
        tmem_ld_32x32(tl + ((r & 3) * 32), v);
        tmem_ld_wait();
        acc ^= v[0] ^ v[31];
      }
      const long long t1 = clock64();
      if (blockIdx.x == 0 && warp == 2 && (threadIdx.x & 31) == 0) out[1] = t1 - t0;
      if (acc == 0x12345) out[7] = acc;
    }
    if (MODE >= 6) {
      const uint32_t tl = tm + (((warp & 3) * 32u) << 16);
      uint32_t acc = 0;
      const long long t0 = clock64();
      for (int r = 0; r < LOADS; ++r) {
        if (MODE == 6) {          // two x32 loads in flight before one wait (8 KB per iteration)
          uint32_t v[32], w[32];
          tmem_ld_32x32(tl + ((r & 1) * 64), v);
          tmem_ld_32x32(tl + ((r & 1) * 64) + 32, w);
          tmem_ld_wait();
          acc ^= v[0] ^ w[31];
        }
        if (MODE == 7) {          // 32x32b.x64: 8 KB per warp instruction
          uint32_t v[64];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63]) : "r"(tl + ((r & 1) * 64)));
          tmem_ld_wait();
          acc ^= v[0] ^ v[63];
        }
        if (MODE == 8) {          // 16x256b.x8: 32 registers
          uint32_t v[32];
          asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(tl + ((r & 3) * 64)));
          tmem_ld_wait();
          acc ^= v[0] ^ v[31];
        }
        if (MODE == 9) {          // 32x32b.x16 x2 in flight (2 KB each)
          uint32_t v[16], w[16];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(tl + ((r & 3) * 32)));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "r"(tl + ((r & 3) * 32) + 16));
          tmem_ld_wait();
          acc ^= v[0] ^ w[15];
        }
      }
      const long long t1 = clock64();
      if (blockIdx.x == 0 && warp == 2 && (threadIdx.x & 31) == 0) out[1] = (t1 - t0) / ((MODE == 6 || MODE == 7) ? 2 : 1);
      if (acc == 0x12345) out[7] = acc;
    }
    if (MODE == 4 || MODE == 5) {
      const uint32_t tl = tm + (((warp & 3) * 32u) << 16);
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = i + threadIdx.x;
      const long long t0 = clock64();
      for (int r = 0; r < LOADS; ++r) {
        tmem_st_32x32(tl + ((r & 3) * 32), v);
        tmem_st_wait();
      }
      const long long t1 = clock64();
      if (blockIdx.x == 0 && warp == 2 && (threadIdx.x & 31) == 0) out[1] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
}

template <int MODE, int NW>
void run(long long* d, const char* name) {
  constexpr int LOADS = 2048, REP = 64;
  cudaFuncSetAttribute(probe<MODE, NW, LOADS, REP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaMemset(d, 0, 64);
  probe<MODE, NW, LOADS, REP><<<148, 64 + 32 * NW, 200 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s warps %2d : ", name, NW);
  if (h[1]) printf("TMEM port %6.1f B/cycle/SM (%5.1f cycles per 4 KB warp access)  ", double(NW) * LOADS * 4096 / h[1], double(h[1]) / LOADS);
  if (h[0]) printf("%6.1f cycles / product", double(h[0]) / (REP * 8));
  printf("  (%s)\n", cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * 8);
  run<0, 4>(d, "loads only");
  run<0, 8>(d, "loads only");
  run<0, 16>(d, "loads only");
  run<1, 8>(d, "SS N=128 only");
  run<2, 4>(d, "SS N=128 + loads");
  run<2, 8>(d, "SS N=128 + loads");
  run<2, 16>(d, "SS N=128 + loads");
  run<3, 8>(d, "TS N=128 + loads");
  run<6, 4>(d, "2 x (32x32b.x32) per wait");
  run<6, 8>(d, "2 x (32x32b.x32) per wait");
  run<7, 4>(d, "32x32b.x64");
  run<7, 8>(d, "32x32b.x64");
  run<8, 4>(d, "16x256b.x8");
  run<8, 8>(d, "16x256b.x8");
  run<9, 4>(d, "2 x (32x32b.x16) per wait");
  run<9, 8>(d, "2 x (32x32b.x16) per wait");
  run<4, 8>(d, "stores only");
  run<5, 8>(d, "SS N=128 + stores");
  return 0;
}
