// Throughput probe for the softmax inner loop's pipes: ex2.approx.ftz.f32 (XU), packed fp32 FMA (FFMA2), and their mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 xu_rate.cu -o xu_rate
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i);
  unsigned long long w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = 0x3f8000003f800000ull + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = ex2f(v[i]);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = ffma2(w[i], w[(i + 1) & 7], w[(i + 3) & 7]);
    }
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], v[(i + 1) & 15], v[(i + 5) & 15]);
    }
  }
  const long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += __uint_as_float(static_cast<uint32_t>(w[i]));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, float* out, long long* cyc, double ops_per_iter_per_thread) {
  const int iters = 4096;
  probe<MODE><<<148, warps * 32>>>(out, cyc, iters);
  probe<MODE><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SM %2d : %6.2f lane-ops / cycle / SM\n", name, warps, ops_per_iter_per_thread * iters * warps * 32 / double(h));
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  for (int warps : {4, 8, 16, 32}) {
    run<0>("ex2.approx.ftz.f32", warps, out, cyc, 16);
    run<1>("fma.rn.f32x2 (2 flop-lanes)", warps, out, cyc, 16 * 2);
    run<3>("fma.rn.f32", warps, out, cyc, 16);
    run<2>("ex2 x16 + ffma2 x16 mixed", warps, out, cyc, 16);
  }
  return 0;
}
