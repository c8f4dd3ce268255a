// Stand-alone laboratory for the attention kernel: builds ONE translation unit (the kernel source is #included, chosen
// with -DATT_SRC=...), checks sampled rows against a float64 CPU softmax(QK^T)V, times the bench shape with CUDA events
// and, with -DATT_TRACE, dumps the per-role event timeline the kernel recorded. Development tool only (not shipped in
// the library): `nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -DATT_SRC=... att_lab.cu -o x`.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#ifndef ATT_SRC
#define ATT_SRC "../../audio_llama_b200/csrc/attention_sm100.cu"
#endif
#include ATT_SRC
#ifdef ATT_LAB_QLOG2
#define ATT_LAB_QMUL 1.4426950408889634f
#define ATT_LAB_EXPMUL 0.6931471805599453
#else
#define ATT_LAB_QMUL 1.0f
#define ATT_LAB_EXPMUL 1.0
#endif
#ifndef ATT_LAB_CTAS_PER_SM
#define ATT_LAB_CTAS_PER_SM 2
#endif
#ifdef ATT_LAB_OLDSIG
#define ATT_LAUNCH(tm, q, o, B, T, H) al::launch_attention(tm, o, B, T, H, 0)
#else
#ifdef ATT_LAB_QLOG2
#define ATT_LAUNCH(tm, q, o, B, T, H) al::launch_attention(tm, q, o, B, T, H, 1, 0)
#else
#define ATT_LAUNCH(tm, q, o, B, T, H) al::launch_attention(tm, q, o, B, T, H, 0, 0)
#endif
#endif

namespace al {
static char g_err[512];
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    enc = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                   const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled -> %d", (int)r);
    return -1;
  }
  return 0;
}
}  // namespace al

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

static float bf2f(uint16_t v) {
  uint32_t u = (uint32_t)v << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fff + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static uint32_t rng_state = 12345;
static float frand() {   // approx normal: sum of 4 uniforms
  float s = 0;
  for (int i = 0; i < 4; ++i) {
    rng_state = rng_state * 1664525u + 1013904223u;
    s += (rng_state >> 8) * (1.0f / 16777216.0f);
  }
  return (s - 2.0f) * 1.7320508f;
}

static int run_case(int B, int T, int H, float qscale, int iters, bool check) {
  const size_t d = (size_t)H * 64, d3 = 3 * d;
  std::vector<uint16_t> h((size_t)B * T * d3);
  for (size_t i = 0; i < h.size(); ++i) {
    const size_t c = i % d3;
    float v = frand();
    if (c < d) v *= qscale * 0.125f * 3.0f * ATT_LAB_QMUL;
    h[i] = f2bf(v);
  }
  uint16_t *dq, *dout;
  CK(cudaMalloc(&dq, h.size() * 2));
  CK(cudaMalloc(&dout, (size_t)B * T * d * 2));
  CK(cudaMemcpy(dq, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, (size_t)B * T * d * 2));
  CUtensorMap tm;
  const uint64_t dims[3] = {d3, (uint64_t)T, (uint64_t)B};
  const uint64_t str[3] = {2, d3 * 2, d3 * 2 * (uint64_t)T};
  const uint32_t box[3] = {64, ATT_LAB_BOX_ROWS, 1};
  if (al::make_tmap(&tm, dq, 2, 3, dims, str, box, true)) {
    printf("tmap: %s\n", al::g_err);
    return 1;
  }
  int rc = ATT_LAUNCH(tm, dq, dout, B, T, H);
  if (rc) {
    printf("launch failed: %s\n", al::g_err);
    return 1;
  }
  CK(cudaDeviceSynchronize());
  int bad = 0;
  if (check) {
    std::vector<uint16_t> o((size_t)B * T * d);
    CK(cudaMemcpy(o.data(), dout, o.size() * 2, cudaMemcpyDeviceToHost));
    double worst = 0, num = 0, den = 0;
    const int nsamp = 160;
    for (int sidx = 0; sidx < nsamp; ++sidx) {
      rng_state = rng_state * 1664525u + 1013904223u;
      const int b = (rng_state >> 8) % B;
      rng_state = rng_state * 1664525u + 1013904223u;
      const int hh = (rng_state >> 8) % H;
      rng_state = rng_state * 1664525u + 1013904223u;
      int q = (rng_state >> 8) % T;
      if (sidx < 8) q = std::max(0, std::min(T - 1, sidx < 4 ? sidx * 37 : T - 1 - (sidx - 4) * 31));   // edges
      const uint16_t* base = h.data() + (size_t)b * T * d3;
      std::vector<double> sc(T);
      double mx = -1e300;
      for (int j = 0; j < T; ++j) {
        double a = 0;
        for (int c = 0; c < 64; ++c) a += (double)bf2f(base[(size_t)q * d3 + hh * 64 + c]) * bf2f(base[(size_t)j * d3 + d + hh * 64 + c]);
        sc[j] = a;
        mx = std::max(mx, a);
      }
      double l = 0;
      for (int j = 0; j < T; ++j) {
        sc[j] = exp((sc[j] - mx) * ATT_LAB_EXPMUL);
        l += sc[j];
      }
      for (int c = 0; c < 64; ++c) {
        double a = 0;
        for (int j = 0; j < T; ++j) a += sc[j] * bf2f(base[(size_t)j * d3 + 2 * d + hh * 64 + c]);
        a /= l;
        const double got = bf2f(o[((size_t)b * T + q) * d + hh * 64 + c]);
        const double e = fabs(got - a);
        if (!(e <= 3e-2 * std::max(1.0, fabs(a)))) {
          if (bad < 5) printf("  mismatch b=%d h=%d q=%d c=%d got %f want %f\n", b, hh, q, c, got, a);
          ++bad;
        }
        worst = std::max(worst, e);
        num += e * e;
        den += a * a;
      }
    }
    printf("check B=%d T=%d H=%d qs=%.1f: max abs err %.4g rel-L2 %.3g %s\n", B, T, H, qscale, worst, sqrt(num / den),
           bad ? "FAIL" : "ok");
  }
  if (iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) ATT_LAUNCH(tm, dq, dout, B, T, H);
#ifdef ATT_CYCLES
    unsigned long long cy[2];
    CK(cudaDeviceSynchronize());
    al::att_cycles_read(cy, true);
#endif
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) ATT_LAUNCH(tm, dq, dout, B, T, H);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    const double fl = 4.0 * T * T * H * 64 * B;
    printf("time B=%d T=%d H=%d: %.4f ms  %.1f TFLOP/s\n", B, T, H, ms, fl / ms / 1e9);
#ifdef ATT_CYCLES
    {
      al::att_cycles_read(cy, true);
      int nsm = 0, occ = 0;
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
      // CTA-cycles summed over CTAs / (resident CTAs per SM * SMs) ~ kernel cycles when the grid keeps every slot busy
      const double cta_cyc = (double)cy[0] / (double)cy[1];
      const double tiles_per_cta = (T + 127) / 128;
      const double kcyc = (double)cy[0] / iters / (ATT_LAB_CTAS_PER_SM * nsm);
      printf("cycles: %.0f per CTA (%.0f per kv tile), ~%.0f kernel cycles -> %.0f MHz effective, %.0f cycles per 128x128 tile per SM\n",
             cta_cyc, cta_cyc / tiles_per_cta, kcyc, kcyc / (ms * 1e3), cta_cyc / tiles_per_cta / ATT_LAB_CTAS_PER_SM);
      (void)occ;
    }
#endif
  }
  cudaFree(dq);
  cudaFree(dout);
  return bad != 0;
}

int main(int argc, char** argv) {
  const char* name = argc > 1 ? argv[1] : "variant";
  printf("== %s\n", name);
  if (argc > 2 && !strcmp(argv[2], "prof")) {   // one short case for ncu
    run_case(argc > 3 ? atoi(argv[3]) : 8, 1500, 20, 1.0f, 1, false);
    return 0;
  }
  int fails = 0;
  const int cases[][3] = {{1, 128, 1}, {1, 256, 2}, {2, 300, 3}, {1, 1500, 6}, {1, 92, 1}, {1, 1000, 2}, {3, 100, 40}, {2, 1500, 20}};
  const float qs[] = {1.0f, 1.0f, 2.0f, 1.0f, 1.0f, 0.2f, 1.0f, 3.0f};
  for (int i = 0; i < 8; ++i) fails += run_case(cases[i][0], cases[i][1], cases[i][2], qs[i], 0, true);
  // adversarial: strongly growing scores along kv (forces reference updates) and huge scores
  fails += run_case(1, 1500, 2, 12.0f, 0, true);
#ifdef ATT_TRACE
  al::att_trace_reset();
  run_case(4, 1500, 20, 1.0f, 0, false);
  al::att_trace_dump();
#endif
  run_case(32, 1500, 20, 1.0f, argc > 2 ? atoi(argv[2]) : 20, false);
  printf("== %s: %s\n", name, fails ? "FAILED" : "all checks ok");
  return fails ? 1 : 0;
}
