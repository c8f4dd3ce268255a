// C ABI (include/audiollm_b200.h): argument checks, tensor-map construction, the encoder plan, constant tables.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "../../include/audiollm_b200.h"
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "mel_bank_struct.h"

namespace al {

static thread_local char g_err[512] = "";
static long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ----------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  EncodeTiledFn enc = get_encode();
  AL_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  // the driver call needs a current context on THIS thread: autograd runs backward functions on its own threads, where
  // only the runtime has been used so far (CUDA_ERROR_INVALID_CONTEXT otherwise)
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    cudaFree(nullptr);
    ctx_bound = true;
  }
  AL_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base %p is not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      AL_REQUIRE(strides_bytes[i] % 16 == 0, "tensor map stride %llu (dim %d) is not a multiple of 16 bytes",
                 (unsigned long long)strides_bytes[i], i);
      gstr[i - 1] = strides_bytes[i];
    }
  }
  const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AL_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu)", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0));
  return 0;
}

// A operand / output: [batch][rows][cols], cols contiguous; strides in elements.
static int tmap_rows3d(CUtensorMap* m, const void* base, int elem_bytes, uint64_t cols, uint64_t rows, uint64_t batch,
                       uint64_t row_stride, uint64_t batch_stride, uint32_t box_cols, uint32_t box_rows) {
  const uint64_t dims[3] = {cols, rows, batch};
  const uint64_t str[3] = {(uint64_t)elem_bytes, row_stride * elem_bytes, batch_stride * elem_bytes};
  const uint32_t box[3] = {box_cols, box_rows, 1};
  return make_tmap(m, base, elem_bytes, 3, dims, str, box, true);
}
// W operand: [N][K] bf16.
static int tmap_weight(CUtensorMap* m, const void* W, uint64_t N, uint64_t K) {
  const uint64_t dims[2] = {K, N};
  const uint64_t str[2] = {2, K * 2};
  const uint32_t box[2] = {64, 128};   // half of a 256-row weight tile: one CTA of a pair stages one box
  return make_tmap(m, W, 2, 2, dims, str, box, true);
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ----------------------------------------------------------------------------- mel tables
static std::vector<double> linspace(double a, double b, int n) {
  std::vector<double> v(n);
  const double step = (b - a) / (n - 1);
  for (int i = 0; i < n; ++i) v[i] = i * step + a;
  v[n - 1] = b;
  return v;
}
// mode 0: transformers.audio_utils.mel_filter_bank(201, n_mels, 0, 8000, 16000, norm="slaney", mel_scale="slaney")
//         (audio_utils.py:263-333, 356-375, 527-546). mode 1: torchaudio melscale_fbanks(201, 0, 8000, n_mels,
//         16000, norm=None, mel_scale="htk").
// torch.linspace(start, end, steps) in float32 (ATen RangeFactories: symmetric evaluation from both ends).
static std::vector<float> linspace_f32(float a, float b, int n) {
  std::vector<float> v(n);
  const float step = (b - a) / (float)(n - 1);
  const int half = n / 2;
  for (int i = 0; i < n; ++i) v[i] = i < half ? a + step * (float)i : b - step * (float)(n - i - 1);
  return v;
}
static void build_filterbank(int n_mels, int mode, std::vector<double>& fb /*[201][n_mels]*/) {
  const int nf = 201;
  fb.assign((size_t)nf * n_mels, 0.0);
  if (mode == 0) {
    std::vector<double> hz(n_mels + 2);
    const double logstep = 27.0 / log(6.4);
    auto h2m = [&](double f) { return f >= 1000.0 ? 15.0 + log(f / 1000.0) * logstep : 3.0 * f / 200.0; };
    auto m2h = [&](double m) { return m >= 15.0 ? 1000.0 * exp((log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0; };
    std::vector<double> mel = linspace(h2m(0.0), h2m(8000.0), n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) hz[i] = m2h(mel[i]);
    std::vector<double> freqs = linspace(0.0, 8000.0, nf);
    for (int k = 0; k < nf; ++k)
      for (int m = 0; m < n_mels; ++m) {
        const double down = (freqs[k] - hz[m]) / (hz[m + 1] - hz[m]);
        const double up = (hz[m + 2] - freqs[k]) / (hz[m + 2] - hz[m + 1]);
        fb[(size_t)k * n_mels + m] = fmax(0.0, fmin(down, up)) * (2.0 / (hz[m + 2] - hz[m]));
      }
  } else {
    // torchaudio works in float32 tensors with python-float (double) end points; its weights are the parity
    // target for M2, so the same float32 arithmetic is restated here.
    const double mmin = 2595.0 * log10(1.0 + 0.0 / 700.0), mmax = 2595.0 * log10(1.0 + 8000.0 / 700.0);
    std::vector<float> freqs = linspace_f32(0.f, 8000.f, nf);
    std::vector<float> mel = linspace_f32((float)mmin, (float)mmax, n_mels + 2);
    std::vector<float> hz(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) hz[i] = 700.0f * (powf(10.0f, mel[i] / 2595.0f) - 1.0f);
    for (int k = 0; k < nf; ++k)
      for (int m = 0; m < n_mels; ++m) {
        const float down = (-1.0f * (hz[m] - freqs[k])) / (hz[m + 1] - hz[m]);
        const float up = (hz[m + 2] - freqs[k]) / (hz[m + 2] - hz[m + 1]);
        fb[(size_t)k * n_mels + m] = (double)fmaxf(0.f, fminf(down, up));
      }
  }
}

struct MelDeviceTables {
  MelTables t;
};
static std::mutex g_mel_mu;
static std::map<int, MelDeviceTables> g_mel_tables;   // key = mode * 1024 + n_mels (one device per process)
static std::map<int, std::vector<double>> g_mel_override;   // host-supplied banks (al_mel_set_filterbank_host)

// Operands of the tensor-core form (mel_tc.cu): the four folded-DFT twiddle matrices as fp16 hi + lo in the
// shared-memory image of each CTA of the pair, the windows, and the filter weights in the order of the compiled
// bank structure (mel_bank_struct.h). A bank whose non-zero pattern matches none of the compiled structures keeps
// tc_bank = -1 and runs on the FFT kernel.
static const uint8_t* g_mel_tc_image = nullptr;   // same for every bank
static int build_mel_tc_tables(const std::vector<double>& fb, int n_mels, int mode, MelTables* t) {
  const double PI = 3.14159265358979323846;
  t->tc_bank = -1;
  t->tc_b_image = nullptr;
  int bank = -1;
  std::vector<float> wts(201 * 2, 0.f);
  for (int c = 0; c < MEL_N_BANKS && bank < 0; ++c) {
    if (MEL_BANK_NMELS[c] != n_mels || MEL_BANK_MODE[c] != mode) continue;
    bool same = true;
    for (int k = 0; k < 201 && same; ++k)
      for (int m = 0; m < n_mels && same; ++m) {
        const bool nz = (float)fb[(size_t)k * n_mels + m] != 0.f;
        same = nz == (m == MEL_BANK_LO[c][k] || m == MEL_BANK_HI[c][k]);
      }
    if (same) bank = c;
  }
  if (bank < 0) return 0;
  for (int k = 0; k < 201; ++k) {
    if (MEL_BANK_LO[bank][k] >= 0) wts[2 * k] = (float)fb[(size_t)k * n_mels + MEL_BANK_LO[bank][k]];
    if (MEL_BANK_HI[bank][k] >= 0) wts[2 * k + 1] = (float)fb[(size_t)k * n_mels + MEL_BANK_HI[bank][k]];
  }
  int rc = mel_tc_set_weights(bank, wts.data());
  if (rc) return rc;
  if (g_mel_tc_image == nullptr) {
    std::vector<float> win(2 * 112, 0.f);
    for (int kp = 0; kp < 112; ++kp) {           // K position 16 b + e  <->  sample index i = b + 7 e
      const int i = (kp >> 4) + 7 * (kp & 15);
      if (i > 100) continue;
      win[2 * kp] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / 400.0));
      win[2 * kp + 1] = (float)(0.5 - 0.5 * cos(2.0 * PI * (200 - i) / 400.0));
    }
    rc = mel_tc_set_window(win.data());
    if (rc) return rc;
    std::vector<uint8_t> img(2 * (size_t)MEL_TC_B_BYTES, 0);
    for (int rank = 0; rank < 2; ++rank)
      for (int mat = 0; mat < 4; ++mat) {            // Ce, Se, Co, So
        const int parity = mat >> 1, is_sin = mat & 1, n_valid = parity ? 100 : 101;
        for (int nl = 0; nl < 56; ++nl) {
          const int jn = 56 * rank + nl;
          for (int kp = 0; kp < 112; ++kp) {         // K position 16 b + e  <->  sample index i = b + 7 e
            const int i = (kp >> 4) + 7 * (kp & 15);
            double v = 0.0;
            if (jn < n_valid && i <= 100) {
              const double ang = 2.0 * PI * (double)i * (double)(2 * jn + parity) / 400.0;
              v = (is_sin ? sin(ang) : cos(ang)) * ((i == 0 || i == 100) ? 0.5 : 1.0);
            }
            const __half hi = __float2half_rn((float)v);
            const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
            const size_t off = (size_t)(kp / 8) * MEL_TC_KCHUNK_BYTES + (size_t)(nl / 8) * 128 + (size_t)(nl % 8) * 16 + (size_t)(kp % 8) * 2;
            memcpy(&img[(size_t)rank * MEL_TC_B_BYTES + (size_t)(2 * mat) * MEL_TC_MAT_BYTES + off], &hi, 2);
            memcpy(&img[(size_t)rank * MEL_TC_B_BYTES + (size_t)(2 * mat + 1) * MEL_TC_MAT_BYTES + off], &lo, 2);
          }
        }
      }
    void* pi;
    AL_CHECK_CUDA(cudaMalloc(&pi, img.size()));
    AL_CHECK_CUDA(cudaMemcpy(pi, img.data(), img.size(), cudaMemcpyHostToDevice));
    g_mel_tc_image = (const uint8_t*)pi;
  }
  t->tc_b_image = g_mel_tc_image;
  t->tc_bank = bank;
  return 0;
}

static int get_mel_tables(int n_mels, int mode, MelTables* out) {
  std::lock_guard<std::mutex> lk(g_mel_mu);
  const int key = mode * 1024 + n_mels;
  auto it = g_mel_tables.find(key);
  if (it != g_mel_tables.end()) {
    *out = it->second.t;
    return 0;
  }
  const double PI = 3.14159265358979323846;
  std::vector<float> window(400);
  for (int i = 0; i < 400; ++i) window[i] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / 400.0));   // torch.hann_window(400)
  std::vector<float2> tw200(200), tw400(101);
  for (int k = 0; k < 200; ++k) tw200[k] = make_float2((float)cos(2.0 * PI * k / 200.0), (float)-sin(2.0 * PI * k / 200.0));
  for (int k = 0; k < 101; ++k) tw400[k] = make_float2((float)cos(2.0 * PI * k / 400.0), (float)-sin(2.0 * PI * k / 400.0));
  std::vector<double> fb;
  auto ov = g_mel_override.find(key);
  if (ov != g_mel_override.end()) fb = ov->second;
  else build_filterbank(n_mels, mode, fb);
  std::vector<int> col_start(n_mels + 1, 0), nz_freq;
  std::vector<float> nz_w;
  for (int m = 0; m < n_mels; ++m) {
    col_start[m] = (int)nz_w.size();
    for (int k = 0; k < 201; ++k) {
      const float w = (float)fb[(size_t)k * n_mels + m];
      if (w != 0.f) {
        nz_freq.push_back(k);
        nz_w.push_back(w);
      }
    }
  }
  col_start[n_mels] = (int)nz_w.size();
  if (nz_w.empty()) {   // keep cudaMalloc sizes non-zero
    nz_freq.push_back(0);
    nz_w.push_back(0.f);
  }
  MelDeviceTables d;
  int rc_tc = 0;
  void *pw, *p2, *p4, *pc, *pf, *pz;
  AL_CHECK_CUDA(cudaMalloc(&pw, 400 * 4));
  AL_CHECK_CUDA(cudaMalloc(&p2, 200 * 8));
  AL_CHECK_CUDA(cudaMalloc(&p4, 101 * 8));
  AL_CHECK_CUDA(cudaMalloc(&pc, (n_mels + 1) * 4));
  AL_CHECK_CUDA(cudaMalloc(&pf, nz_freq.size() * 4));
  AL_CHECK_CUDA(cudaMalloc(&pz, nz_w.size() * 4));
  AL_CHECK_CUDA(cudaMemcpy(pw, window.data(), 400 * 4, cudaMemcpyHostToDevice));
  AL_CHECK_CUDA(cudaMemcpy(p2, tw200.data(), 200 * 8, cudaMemcpyHostToDevice));
  AL_CHECK_CUDA(cudaMemcpy(p4, tw400.data(), 101 * 8, cudaMemcpyHostToDevice));
  AL_CHECK_CUDA(cudaMemcpy(pc, col_start.data(), (n_mels + 1) * 4, cudaMemcpyHostToDevice));
  AL_CHECK_CUDA(cudaMemcpy(pf, nz_freq.data(), nz_freq.size() * 4, cudaMemcpyHostToDevice));
  AL_CHECK_CUDA(cudaMemcpy(pz, nz_w.data(), nz_w.size() * 4, cudaMemcpyHostToDevice));
  d.t.window = (const float*)pw;
  d.t.tw200 = (const float2*)p2;
  d.t.tw400 = (const float2*)p4;
  d.t.col_start = (const int*)pc;
  d.t.nz_freq = (const int*)pf;
  d.t.nz_w = (const float*)pz;
  d.t.n_mels = n_mels;
  d.t.nnz = (int)nz_w.size();
  rc_tc = build_mel_tc_tables(fb, n_mels, mode, &d.t);
  if (rc_tc) return rc_tc;
  if (d.t.nnz > 1024 || n_mels > 256) {
    set_error("mel filter bank too dense for the kernel's shared-memory tables (nnz=%d, n_mels=%d)", d.t.nnz, n_mels);
    return -1;
  }
  g_mel_tables[key] = d;
  *out = d.t;
  return 0;
}

// ----------------------------------------------------------------------------- resample tables
// torchaudio functional._get_sinc_resample_kernel (TA functional.py:1305-1403), fp64 then cast to fp32.
static std::map<long long, ResampleTable> g_resample_tables;

static int get_resample_table(int orig_freq, int new_freq, ResampleTable* out) {
  std::lock_guard<std::mutex> lk(g_mel_mu);
  const long long key = (static_cast<long long>(orig_freq) << 32) | (unsigned)new_freq;
  auto it = g_resample_tables.find(key);
  if (it != g_resample_tables.end()) {
    *out = it->second;
    return 0;
  }
  int a = orig_freq, b = new_freq;
  while (b) { const int t = a % b; a = b; b = t; }
  const int orig = orig_freq / a, nw = new_freq / a;
  const double PI = 3.14159265358979323846, lpw = 6.0, rolloff = 0.99;
  const double base = (orig < nw ? orig : nw) * rolloff;
  const int width = (int)ceil(lpw * orig / base);
  const int L = 2 * width + orig;
  std::vector<float> dense((size_t)nw * L);
  int max_taps = 1;
  std::vector<int> first(nw, 0), count(nw, 0);
  for (int p = 0; p < nw; ++p) {
    int f = -1, l = -1;
    for (int k = 0; k < L; ++k) {
      double t = ((double)(-p) / nw + (double)(k - width) / orig) * base;
      t = t < -lpw ? -lpw : (t > lpw ? lpw : t);
      const double c = cos(t * PI / lpw / 2.0);
      const double window = c * c;
      t *= PI;
      const double sinc = (t == 0.0) ? 1.0 : sin(t) / t;
      const float w = (float)(sinc * window * (base / orig));
      dense[(size_t)p * L + k] = w;
      if (w != 0.f) {
        if (f < 0) f = k;
        l = k;
      }
    }
    first[p] = f < 0 ? 0 : f;
    count[p] = f < 0 ? 0 : l - f + 1;
    if (count[p] > max_taps) max_taps = count[p];
  }
  std::vector<float> packed((size_t)nw * max_taps, 0.f);
  for (int p = 0; p < nw; ++p)
    for (int i = 0; i < count[p]; ++i) packed[(size_t)p * max_taps + i] = dense[(size_t)p * L + first[p] + i];
  void *pw, *pf;
  AL_CHECK_CUDA(cudaMalloc(&pw, packed.size() * 4));
  AL_CHECK_CUDA(cudaMalloc(&pf, first.size() * 4));
  AL_CHECK_CUDA(cudaMemcpy(pw, packed.data(), packed.size() * 4, cudaMemcpyHostToDevice));
  AL_CHECK_CUDA(cudaMemcpy(pf, first.data(), first.size() * 4, cudaMemcpyHostToDevice));
  ResampleTable t{(const float*)pw, (const int*)pf, orig, nw, width, max_taps};
  g_resample_tables[key] = t;
  *out = t;
  return 0;
}

}  // namespace al

using namespace al;

// ============================================================================= encoder plan
struct LayerW {
  const float *ln1_g, *ln1_b, *bqkv, *bo, *ln2_g, *ln2_b, *b1, *b2;
  const void *wqkv, *wo, *w1, *w2;
  CUtensorMap tm_wqkv, tm_wo, tm_w1, tm_w2;
  bool set = false;
};

struct al_encoder {
  int d, L, H, ffn, n_mels, c_pad, max_batch;
  int att_flags = 0;   // AL_ATT_* for the attention launches (al_encoder_set_options)
  static constexpr int T_MEL = 3000, T = 1500;
  // workspace
  uint8_t* ws;
  size_t ws_bytes;
  void* melT;   // bf16 [Bm][3002][c_pad]
  void* h1;     // bf16 [Bm][3002][d]
  float* x;     // f32  [Bm*1500][d]
  void* xn;     // bf16 [Bm*1500][d]
  void* attn;   // bf16 [Bm*1500][d]
  void* qkv;    // bf16 [Bm*1500][3d]
  void* hff;    // bf16 [Bm*1500][ffn]
  // stem
  const void *conv1_w = nullptr, *conv2_w = nullptr;
  const float *conv1_b = nullptr, *conv2_b = nullptr, *pos = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
  CUtensorMap tm_conv1_w, tm_conv2_w;
  std::vector<LayerW> layers;
  // optional live profiling: CUDA events around every launch of a forward, on the launch stream
  bool profiling = false;
  std::vector<cudaEvent_t> ev;          // pairs, appended per launch
  std::vector<int> ev_kind;
  size_t ev_used = 0;
  // activation maps (depend on B only through the flat row count)
  int maps_B = -1;
  CUtensorMap tm_melT_A, tm_h1_O, tm_h1_A, tm_x_O3, tm_xn_A, tm_qkv_O, tm_qkv_att, tm_attn_A, tm_x_red, tm_hff_O, tm_hff_A;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void encoder_carve(al_encoder* e, bool assign) {
  size_t off = 0;
  const size_t Bm = e->max_batch;
  auto take = [&](size_t bytes) {
    void* p = assign ? e->ws + off : nullptr;
    off = align_up(off + bytes, 1024);
    return p;
  };
  e->melT = take(Bm * 3002 * e->c_pad * 2);
  e->h1 = take(Bm * 3002 * (size_t)e->d * 2);
  e->x = (float*)take(Bm * 1500 * (size_t)e->d * 4);
  e->xn = take(Bm * 1500 * (size_t)e->d * 2);
  e->attn = take(Bm * 1500 * (size_t)e->d * 2);
  e->qkv = take(Bm * 1500 * (size_t)e->d * 3 * 2);
  e->hff = take(Bm * 1500 * (size_t)e->ffn * 2);
  e->ws_bytes = off;
}

extern "C" {

int al_version(void) { return 101; }
int al_gemm_set_mode(int pair) {
  gemm_set_mode(pair);
  return 0;
}
const char* al_last_error(void) { return g_err; }
long long al_launch_count(void) { return g_launches; }

// ----------------------------------------------------------------------------- mel
static int g_mel_tc = -1;   // -1 = read AUDIOLLM_B200_MEL (tc | fft) once; default tc
static bool mel_use_tc() {
  if (g_mel_tc < 0) {
    const char* e = getenv("AUDIOLLM_B200_MEL");
    g_mel_tc = (e && strcmp(e, "fft") == 0) ? 0 : 1;
  }
  return g_mel_tc == 1;
}
int al_mel_set_mode(int tc) {
  g_mel_tc = tc ? 1 : 0;
  return 0;
}

int al_mel_forward(const float* wave, const int* n_samples, int n_clips, long long wave_stride, int n_mels, int mode,
                   float* out, unsigned int* clip_max_ws, al_stream_t stream) {
  return al_mel_forward_ex(wave, n_samples, n_clips, wave_stride, n_mels, mode, 0, out, clip_max_ws, stream);
}

int al_mel_forward_ex(const float* wave, const int* n_samples, int n_clips, long long wave_stride, int n_mels, int mode,
                      int flags, float* out, unsigned int* clip_max_ws, al_stream_t stream) {
  AL_REQUIRE((flags & ~AL_MEL_RAW) == 0, "al_mel_forward_ex: unknown flags 0x%x", flags);
  AL_REQUIRE(n_clips >= 0 && n_mels > 0 && n_mels <= 256, "al_mel_forward: bad n_clips=%d / n_mels=%d", n_clips, n_mels);
  AL_REQUIRE(mode == 0 || mode == 1, "al_mel_forward: mode must be 0 (whisper) or 1 (train), got %d", mode);
  AL_REQUIRE(n_samples != nullptr || wave_stride >= 480000,
             "al_mel_forward: without n_samples every clip must hold 480000 samples (wave_stride=%lld)", wave_stride);
  AL_REQUIRE(mode == 1 || clip_max_ws != nullptr, "al_mel_forward: mode 0 needs clip_max_ws");
  if (n_clips == 0) return 0;
  MelTables tb;
  int rc = get_mel_tables(n_mels, mode, &tb);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (tb.tc_bank >= 0 && mel_use_tc()) rc = launch_mel_tc(wave, n_samples, n_clips, wave_stride, tb, mode, out, clip_max_ws, num_sms(), st);
  else rc = launch_mel(wave, n_samples, n_clips, wave_stride, tb, mode, out, clip_max_ws, st);
  if (rc) return rc;
  g_launches += 1;
  if (mode == 0 && !(flags & AL_MEL_RAW)) {
    rc = launch_mel_finalize(out, clip_max_ws, n_clips, n_mels, st);
    if (rc) return rc;
    g_launches += 1;
  }
  return 0;
}

int al_mel_filterbank_host(int n_mels, int mode, double* out_host) {
  AL_REQUIRE(n_mels > 0 && out_host != nullptr && (mode == 0 || mode == 1), "al_mel_filterbank_host: bad arguments");
  std::vector<double> fb;
  {
    std::lock_guard<std::mutex> lk(g_mel_mu);
    auto ov = g_mel_override.find(mode * 1024 + n_mels);
    if (ov != g_mel_override.end()) fb = ov->second;
  }
  if (fb.empty()) build_filterbank(n_mels, mode, fb);
  memcpy(out_host, fb.data(), fb.size() * sizeof(double));
  return 0;
}

int al_mel_set_filterbank_host(int n_mels, int mode, const double* fb_host) {
  AL_REQUIRE(n_mels > 0 && n_mels <= 256 && fb_host != nullptr && (mode == 0 || mode == 1),
             "al_mel_set_filterbank_host: bad arguments");
  std::lock_guard<std::mutex> lk(g_mel_mu);
  const int key = mode * 1024 + n_mels;
  AL_REQUIRE(g_mel_tables.find(key) == g_mel_tables.end(),
             "al_mel_set_filterbank_host: the (n_mels=%d, mode=%d) tables are already on the device", n_mels, mode);
  g_mel_override[key].assign(fb_host, fb_host + (size_t)201 * n_mels);
  return 0;
}

// ----------------------------------------------------------------------------- ingest
int al_ingest_forward(const float* in, long long clip_stride, long long chan_stride, int n_chan, const int* n_in,
                      int n_in_cap, int orig_freq, int new_freq, float* out, long long out_stride, int out_cap,
                      int* n_out, int n_clips, al_stream_t stream) {
  AL_REQUIRE(in && out, "al_ingest_forward: NULL argument");
  AL_REQUIRE(n_clips >= 0 && n_chan >= 1 && n_in_cap >= 0 && out_cap > 0 && out_stride >= out_cap,
             "al_ingest_forward: bad shape clips=%d chan=%d n_in_cap=%d out_cap=%d", n_clips, n_chan, n_in_cap, out_cap);
  AL_REQUIRE(orig_freq > 0 && new_freq > 0, "al_ingest_forward: bad rates %d -> %d", orig_freq, new_freq);
  int rc;
  if (orig_freq == new_freq) {
    rc = launch_ingest(in, clip_stride, chan_stride, n_chan, n_in, n_in_cap, nullptr, out, out_stride, out_cap, n_out,
                       n_clips, (cudaStream_t)stream);
  } else {
    ResampleTable tb;
    if ((rc = get_resample_table(orig_freq, new_freq, &tb))) return rc;
    rc = launch_ingest(in, clip_stride, chan_stride, n_chan, n_in, n_in_cap, &tb, out, out_stride, out_cap, n_out,
                       n_clips, (cudaStream_t)stream);
  }
  if (rc == 0 && n_clips > 0) g_launches += 1;
  return rc;
}

// ----------------------------------------------------------------------------- building blocks
int al_gemm_bf16(const void* A, long long a_row_stride, long long a_batch_stride, int m_per_batch, int batch,
                 const void* W, int N, int K, const float* bias, void* out, long long o_row_stride,
                 long long o_batch_stride, int flags, const float* aux, int aux_ld, const float* resid,
                 al_stream_t stream) {
  AL_REQUIRE(m_per_batch > 0 && batch > 0 && N > 0 && K > 0, "al_gemm_bf16: bad shape m=%d batch=%d N=%d K=%d",
             m_per_batch, batch, N, K);
  AL_REQUIRE(K % 8 == 0, "al_gemm_bf16: K=%d must be a multiple of 8", K);
  AL_REQUIRE(((flags & AL_EPI_ROWAUX) != 0) == (aux != nullptr), "al_gemm_bf16: AL_EPI_ROWAUX and aux must come together");
  AL_REQUIRE(!(flags & AL_EPI_REDUCE_ADD) || (flags & AL_EPI_OUT_F32), "al_gemm_bf16: REDUCE_ADD needs OUT_F32");
  AL_REQUIRE(((flags & AL_EPI_RESIDUAL) != 0) == (resid != nullptr), "al_gemm_bf16: AL_EPI_RESIDUAL and resid must come together");
  AL_REQUIRE(!(flags & AL_EPI_RESIDUAL) || (flags & AL_EPI_OUT_F32), "al_gemm_bf16: RESIDUAL needs OUT_F32");
  CUtensorMap ta, tb, to;
  int rc = tmap_rows3d(&ta, A, 2, K, m_per_batch, batch, a_row_stride, a_batch_stride, 64, 128);
  if (rc) return rc;
  rc = tmap_weight(&tb, W, N, K);
  if (rc) return rc;
  const int oe = (flags & AL_EPI_OUT_F32) ? 4 : 2;
  rc = tmap_rows3d(&to, out, oe, N, m_per_batch, batch, o_row_stride, o_batch_stride, gemm_out_box_cols(flags), 128);
  if (rc) return rc;
  GemmParams p{};
  p.m_per_batch = m_per_batch;
  p.batch = batch;
  p.N = N;
  p.K = K;
  p.bias = bias;
  p.aux = aux;
  p.aux_ld = aux_ld;
  p.resid = resid;                       // same strides as the output
  p.resid_ld = o_row_stride;
  p.resid_batch_stride = o_batch_stride;
  rc = launch_gemm(ta, tb, to, p, flags, num_sms(), (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

int al_layernorm(const float* x, const float* gamma, const float* beta, void* out, int rows, int d, float eps,
                 int out_dtype, long long out_ld, int rows_per_group, long long out_group_stride,
                 long long out_row_offset, al_stream_t stream) {
  int rc = launch_layernorm(x, gamma, beta, out, rows, d, eps, out_dtype, out_ld, rows_per_group, out_group_stride,
                            out_row_offset, (cudaStream_t)stream);
  if (rc == 0 && rows > 0) g_launches += 1;
  return rc;
}

int al_attention(const void* qkv, void* out, int B, int T, int H, al_stream_t stream) {
  return al_attention_ex(qkv, out, B, T, H, 0, stream);
}

int al_attention_ex(const void* qkv, void* out, int B, int T, int H, int flags, al_stream_t stream) {
  AL_REQUIRE(B > 0 && T > 0 && H > 0, "al_attention: bad shape B=%d T=%d H=%d", B, T, H);
  AL_REQUIRE((flags & ~AL_ATT_Q_LOG2) == 0, "al_attention_ex: unknown flags 0x%x", flags);
  CUtensorMap tm;
  const uint64_t d3 = (uint64_t)3 * H * 64;
  int rc = tmap_rows3d(&tm, qkv, 2, d3, T, B, d3, d3 * T, 64, 128);
  if (rc) return rc;
  rc = launch_attention(tm, qkv, out, B, T, H, (flags & AL_ATT_Q_LOG2) ? 1 : 0, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

// [B][S][H][128] bf16 -> rank-4 map {128, H, S, B}, box {64, 1, 128, 1}: one (row tile, head) half-tile per copy
static int tmap_bshd(CUtensorMap* m, const void* base, uint64_t B, uint64_t S, uint64_t H) {
  const uint64_t dims[4] = {128, H, S, B};
  const uint64_t str[4] = {2, 256, H * 256, S * H * 256};
  const uint32_t box[4] = {64, 1, 128, 1};
  return make_tmap(m, base, 2, 4, dims, str, box, true);
}

int al_gqa_attention_forward(const void* q, const void* k, const void* v, void* out, float* lse, const int* kv_len, int B,
                             int S, int Hq, int Hkv, int head_dim, float scale, al_stream_t stream) {
  AL_REQUIRE(q && k && v && out && lse, "al_gqa_attention_forward: NULL argument");
  AL_REQUIRE(head_dim == 128, "al_gqa_attention_forward: head_dim must be 128, got %d", head_dim);
  AL_REQUIRE(B > 0 && S > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0, "al_gqa_attention_forward: bad shape B=%d S=%d Hq=%d Hkv=%d", B, S, Hq, Hkv);
  CUtensorMap tq, tk, tv, to;
  int rc;
  if ((rc = tmap_bshd(&tq, q, B, S, Hq))) return rc;
  if ((rc = tmap_bshd(&tk, k, B, S, Hkv))) return rc;
  if ((rc = tmap_bshd(&tv, v, B, S, Hkv))) return rc;
  if ((rc = tmap_bshd(&to, out, B, S, Hq))) return rc;
  rc = launch_gqa_fwd(tq, tk, tv, to, lse, kv_len, B, S, Hq, Hkv, scale, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

int al_gqa_attention_backward(const void* q, const void* k, const void* v, const void* out, const float* lse,
                              const void* d_out, const int* kv_len, void* dq, void* dk, void* dv, float* dsum_ws, int B, int S,
                              int Hq, int Hkv, int head_dim, float scale, al_stream_t stream) {
  AL_REQUIRE(q && k && v && out && lse && d_out && dq && dk && dv && dsum_ws, "al_gqa_attention_backward: NULL argument");
  AL_REQUIRE(head_dim == 128, "al_gqa_attention_backward: head_dim must be 128, got %d", head_dim);
  AL_REQUIRE(B > 0 && S > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0, "al_gqa_attention_backward: bad shape B=%d S=%d Hq=%d Hkv=%d", B, S, Hq, Hkv);
  CUtensorMap tq, tk, tv, tdo, tdq;
  int rc;
  if ((rc = tmap_bshd(&tq, q, B, S, Hq))) return rc;
  if ((rc = tmap_bshd(&tk, k, B, S, Hkv))) return rc;
  if ((rc = tmap_bshd(&tv, v, B, S, Hkv))) return rc;
  if ((rc = tmap_bshd(&tdo, d_out, B, S, Hq))) return rc;
  if ((rc = tmap_bshd(&tdq, dq, B, S, Hq))) return rc;
  rc = launch_gqa_bwd(tq, tk, tv, tdo, tdq, out, d_out, lse, dsum_ws, kv_len, dq, dk, dv, B, S, Hq, Hkv, scale, (cudaStream_t)stream);
  if (rc == 0) g_launches += 3;
  return rc;
}

int al_pack_mel(const float* mel, void* out_bf16, int B, int n_mels, int T, int c_pad, al_stream_t stream) {
  return al_pack_mel_ex(mel, nullptr, out_bf16, B, n_mels, T, c_pad, stream);
}

int al_pack_mel_ex(const float* mel, const unsigned int* clip_max_ws, void* out_bf16, int B, int n_mels, int T, int c_pad,
                   al_stream_t stream) {
  AL_REQUIRE(c_pad >= n_mels, "al_pack_mel: c_pad=%d < n_mels=%d", c_pad, n_mels);
  int rc = launch_pack_mel(mel, out_bf16, B, n_mels, T, c_pad, clip_max_ws, (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

int al_f32_to_bf16(const float* x, void* out_bf16, long long n, al_stream_t stream) {
  int rc = launch_f32_to_bf16(x, out_bf16, n, (cudaStream_t)stream);
  if (rc == 0 && n > 0) g_launches += 1;
  return rc;
}

// ----------------------------------------------------------------------------- encoder
size_t al_encoder_workspace_bytes(int d_model, int n_layers, int n_heads, int ffn_dim, int n_mels, int max_batch) {
  (void)n_layers;
  (void)n_heads;
  al_encoder e{};
  e.d = d_model;
  e.ffn = ffn_dim;
  e.n_mels = n_mels;
  e.c_pad = (n_mels + 63) / 64 * 64;
  e.max_batch = max_batch;
  encoder_carve(&e, false);
  return e.ws_bytes;
}

int al_encoder_create(al_encoder** out, int d_model, int n_layers, int n_heads, int ffn_dim, int n_mels,
                      int max_batch, void* workspace, size_t workspace_bytes) {
  AL_REQUIRE(out != nullptr, "al_encoder_create: out is NULL");
  AL_REQUIRE(d_model > 0 && n_heads > 0 && d_model == n_heads * 64,
             "al_encoder_create: d_model=%d must be n_heads=%d x 64 (Whisper head_dim)", d_model, n_heads);
  AL_REQUIRE(d_model % 8 == 0 && ffn_dim % 8 == 0 && n_layers >= 0 && max_batch > 0 && n_mels > 0,
             "al_encoder_create: bad shape");
  AL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "al_encoder_create: workspace must be 1024-byte aligned");
  al_encoder* e = new al_encoder();
  e->d = d_model;
  e->L = n_layers;
  e->H = n_heads;
  e->ffn = ffn_dim;
  e->n_mels = n_mels;
  e->c_pad = (n_mels + 63) / 64 * 64;
  e->max_batch = max_batch;
  e->ws = (uint8_t*)workspace;
  encoder_carve(e, true);
  if (e->ws_bytes > workspace_bytes) {
    set_error("al_encoder_create: workspace of %zu bytes is smaller than the %zu needed", workspace_bytes, e->ws_bytes);
    delete e;
    return -1;
  }
  e->layers.resize(n_layers);
  // conv1 output buffer: rows 0 and 3001 of every clip are conv2's zero padding and are never written again
  cudaError_t ce = cudaMemset(e->h1, 0, (size_t)max_batch * 3002 * d_model * 2);
  if (ce != cudaSuccess) {
    set_error("al_encoder_create: cudaMemset failed: %s", cudaGetErrorString(ce));
    delete e;
    return -2;
  }
  *out = e;
  return 0;
}

int al_encoder_set_stem(al_encoder* e, const void* conv1_w, const float* conv1_b, const void* conv2_w,
                        const float* conv2_b, const float* pos, const float* lnf_g, const float* lnf_b) {
  AL_REQUIRE(e && conv1_w && conv1_b && conv2_w && conv2_b && pos && lnf_g && lnf_b, "al_encoder_set_stem: NULL argument");
  e->conv1_w = conv1_w; e->conv1_b = conv1_b; e->conv2_w = conv2_w; e->conv2_b = conv2_b;
  e->pos = pos; e->lnf_g = lnf_g; e->lnf_b = lnf_b;
  int rc = tmap_weight(&e->tm_conv1_w, conv1_w, e->d, 3 * e->c_pad);
  if (rc) return rc;
  return tmap_weight(&e->tm_conv2_w, conv2_w, e->d, 3 * e->d);
}

int al_encoder_set_layer(al_encoder* e, int layer, const float* ln1_g, const float* ln1_b, const void* wqkv,
                         const float* bqkv, const void* wo, const float* bo, const float* ln2_g, const float* ln2_b,
                         const void* w1, const float* b1, const void* w2, const float* b2) {
  AL_REQUIRE(e && layer >= 0 && layer < e->L, "al_encoder_set_layer: layer %d out of range", layer);
  AL_REQUIRE(ln1_g && ln1_b && wqkv && bqkv && wo && bo && ln2_g && ln2_b && w1 && b1 && w2 && b2,
             "al_encoder_set_layer: NULL argument");
  LayerW& w = e->layers[layer];
  w.ln1_g = ln1_g; w.ln1_b = ln1_b; w.wqkv = wqkv; w.bqkv = bqkv; w.wo = wo; w.bo = bo;
  w.ln2_g = ln2_g; w.ln2_b = ln2_b; w.w1 = w1; w.b1 = b1; w.w2 = w2; w.b2 = b2;
  int rc;
  if ((rc = tmap_weight(&w.tm_wqkv, wqkv, 3 * e->d, e->d))) return rc;
  if ((rc = tmap_weight(&w.tm_wo, wo, e->d, e->d))) return rc;
  if ((rc = tmap_weight(&w.tm_w1, w1, e->ffn, e->d))) return rc;
  if ((rc = tmap_weight(&w.tm_w2, w2, e->d, e->ffn))) return rc;
  w.set = true;
  return 0;
}

static int prof_events(al_encoder* e, int kind, cudaEvent_t* e0, cudaEvent_t* e1) {
  if (e->ev_used + 2 > e->ev.size()) {
    if (e->ev.size() >= 2 * 65536) return -1;     // cap: stop recording rather than grow without bound
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t ev;
      if (cudaEventCreate(&ev) != cudaSuccess) return -1;
      e->ev.push_back(ev);
    }
    e->ev_kind.push_back(kind);
  }
  e->ev_kind[e->ev_used / 2] = kind;
  *e0 = e->ev[e->ev_used];
  *e1 = e->ev[e->ev_used + 1];
  e->ev_used += 2;
  return 0;
}

static int encoder_maps(al_encoder* e, int B) {
  if (e->maps_B == B) return 0;
  const uint64_t d = e->d, c = e->c_pad, f = e->ffn, rows = (uint64_t)B * 1500;
  int rc;
  // conv1: im2col row t = melT rows t..t+2 (3*c_pad contiguous), row stride c_pad, clip stride 3002*c_pad
  if ((rc = tmap_rows3d(&e->tm_melT_A, e->melT, 2, 3 * c, 3000, B, c, 3002 * c, 64, 128))) return rc;
  // conv1 output -> h1 rows 1..3000 of each clip
  if ((rc = tmap_rows3d(&e->tm_h1_O, (uint8_t*)e->h1 + d * 2, 2, d, 3000, B, d, 3002 * d, 64, 128))) return rc;
  // conv2 (stride 2): im2col row t2 = h1 rows 2*t2..2*t2+2, row stride 2*d
  if ((rc = tmap_rows3d(&e->tm_h1_A, e->h1, 2, 3 * d, 1500, B, 2 * d, 3002 * d, 64, 128))) return rc;
  if ((rc = tmap_rows3d(&e->tm_x_O3, e->x, 4, d, 1500, B, d, 1500 * d, 32, 128))) return rc;
  // flat [B*1500] activations
  if ((rc = tmap_rows3d(&e->tm_xn_A, e->xn, 2, d, rows, 1, d, rows * d, 64, 128))) return rc;
  if ((rc = tmap_rows3d(&e->tm_qkv_O, e->qkv, 2, 3 * d, rows, 1, 3 * d, rows * 3 * d, 64, 128))) return rc;
  if ((rc = tmap_rows3d(&e->tm_qkv_att, e->qkv, 2, 3 * d, 1500, B, 3 * d, 1500 * 3 * d, 64, 128))) return rc;
  if ((rc = tmap_rows3d(&e->tm_attn_A, e->attn, 2, d, rows, 1, d, rows * d, 64, 128))) return rc;
  if ((rc = tmap_rows3d(&e->tm_x_red, e->x, 4, d, rows, 1, d, rows * d, 32, 128))) return rc;
  if ((rc = tmap_rows3d(&e->tm_hff_O, e->hff, 2, f, rows, 1, f, rows * f, 64, 128))) return rc;
  if ((rc = tmap_rows3d(&e->tm_hff_A, e->hff, 2, f, rows, 1, f, rows * f, 64, 128))) return rc;
  e->maps_B = B;
  return 0;
}

int al_encoder_forward(al_encoder* e, const float* mel, int B, void* out, int out_dtype, int n_layers_run,
                       al_stream_t stream) {
  return al_encoder_forward_ex(e, mel, nullptr, B, out, out_dtype, n_layers_run, stream);
}

int al_encoder_forward_ex(al_encoder* e, const float* mel, const unsigned int* clip_max_ws, int B, void* out, int out_dtype,
                          int n_layers_run, al_stream_t stream) {
  AL_REQUIRE(e && mel && out, "al_encoder_forward: NULL argument");
  AL_REQUIRE(B > 0 && B <= e->max_batch, "al_encoder_forward: B=%d outside 1..max_batch=%d", B, e->max_batch);
  AL_REQUIRE(e->conv1_w != nullptr, "al_encoder_forward: stem weights not set");
  const int L = (n_layers_run < 0 || n_layers_run > e->L) ? e->L : n_layers_run;
  for (int l = 0; l < L; ++l) AL_REQUIRE(e->layers[l].set, "al_encoder_forward: layer %d weights not set", l);
  int rc = encoder_maps(e, B);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int nsm = num_sms();
  const int d = e->d, rows = B * 1500;
#define RUN(kind, expr)                                                   \
  do {                                                                    \
    cudaEvent_t _e0 = nullptr, _e1 = nullptr;                             \
    if (e->profiling && prof_events(e, kind, &_e0, &_e1) == 0) cudaEventRecord(_e0, st); \
    rc = (expr);                                                          \
    if (rc) return rc;                                                    \
    if (_e1) cudaEventRecord(_e1, st);                                    \
    g_launches += 1;                                                      \
  } while (0)
  RUN(AL_K_PACK, launch_pack_mel(mel, e->melT, B, e->n_mels, 3000, e->c_pad, clip_max_ws, st));
  {  // conv1 + GELU (modeling_whisper.py:619)
    GemmParams p{};
    p.m_per_batch = 3000; p.batch = B; p.N = d; p.K = 3 * e->c_pad; p.bias = e->conv1_b;
    RUN(AL_K_CONV1, launch_gemm(e->tm_melT_A, e->tm_conv1_w, e->tm_h1_O, p, EPI_GELU, nsm, st));
  }
  {  // conv2 (stride 2) + GELU + position table (:620-625) -> fp32 residual stream
    GemmParams p{};
    p.m_per_batch = 1500; p.batch = B; p.N = d; p.K = 3 * d; p.bias = e->conv2_b; p.aux = e->pos; p.aux_ld = d;
    RUN(AL_K_CONV2, launch_gemm(e->tm_h1_A, e->tm_conv2_w, e->tm_x_O3, p, EPI_OUT_F32 | EPI_GELU | EPI_ROWAUX, nsm, st));
  }
  for (int l = 0; l < L; ++l) {
    const LayerW& w = e->layers[l];
    RUN(AL_K_LN, launch_layernorm(e->x, w.ln1_g, w.ln1_b, e->xn, rows, d, 1e-5f, 0, d, rows, 0, 0, st));
    {
      GemmParams p{};
      p.m_per_batch = rows; p.batch = 1; p.N = 3 * d; p.K = d; p.bias = w.bqkv;
      RUN(AL_K_QKV, launch_gemm(e->tm_xn_A, w.tm_wqkv, e->tm_qkv_O, p, 0, nsm, st));
    }
    RUN(AL_K_ATTN, launch_attention(e->tm_qkv_att, e->qkv, e->attn, B, 1500, e->H, (e->att_flags & AL_ATT_Q_LOG2) ? 1 : 0, st));
    {
      GemmParams p{};
      p.m_per_batch = rows; p.batch = 1; p.N = d; p.K = d; p.bias = w.bo;
      // residual: TMA reduce-add into the fp32 stream (measured faster than reading x in the epilogue threads)
      RUN(AL_K_OPROJ, launch_gemm(e->tm_attn_A, w.tm_wo, e->tm_x_red, p, EPI_OUT_F32 | EPI_REDUCE_ADD, nsm, st));
    }
    RUN(AL_K_LN, launch_layernorm(e->x, w.ln2_g, w.ln2_b, e->xn, rows, d, 1e-5f, 0, d, rows, 0, 0, st));
    {
      GemmParams p{};
      p.m_per_batch = rows; p.batch = 1; p.N = e->ffn; p.K = d; p.bias = w.b1;
      RUN(AL_K_FC1, launch_gemm(e->tm_xn_A, w.tm_w1, e->tm_hff_O, p, EPI_GELU, nsm, st));
    }
    {
      GemmParams p{};
      p.m_per_batch = rows; p.batch = 1; p.N = d; p.K = e->ffn; p.bias = w.b2;
      RUN(AL_K_FC2, launch_gemm(e->tm_hff_A, w.tm_w2, e->tm_x_red, p, EPI_OUT_F32 | EPI_REDUCE_ADD, nsm, st));
    }
  }
  RUN(AL_K_LN, launch_layernorm(e->x, e->lnf_g, e->lnf_b, out, rows, d, 1e-5f, out_dtype, d, rows, 0, 0, st));
#undef RUN
  return 0;
}

float* al_encoder_hidden(al_encoder* e) { return e ? e->x : nullptr; }

int al_encoder_set_options(al_encoder* e, int attention_flags) {
  AL_REQUIRE(e != nullptr, "al_encoder_set_options: NULL plan");
  AL_REQUIRE((attention_flags & ~AL_ATT_Q_LOG2) == 0, "al_encoder_set_options: unknown attention flags 0x%x", attention_flags);
  e->att_flags = attention_flags;
  return 0;
}

int al_encoder_set_profiling(al_encoder* e, int on) {
  AL_REQUIRE(e != nullptr, "al_encoder_set_profiling: NULL plan");
  e->profiling = on != 0;
  e->ev_used = 0;
  return 0;
}

int al_encoder_profile_read(al_encoder* e, float* ms_by_kind_host, int* launches_by_kind_host) {
  AL_REQUIRE(e && ms_by_kind_host && launches_by_kind_host, "al_encoder_profile_read: NULL argument");
  for (int k = 0; k < AL_K_COUNT; ++k) {
    ms_by_kind_host[k] = 0.f;
    launches_by_kind_host[k] = 0;
  }
  for (size_t i = 0; i + 1 < e->ev_used; i += 2) {
    AL_CHECK_CUDA(cudaEventSynchronize(e->ev[i + 1]));
    float ms = 0.f;
    AL_CHECK_CUDA(cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]));
    const int k = e->ev_kind[i / 2];
    ms_by_kind_host[k] += ms;
    launches_by_kind_host[k] += 1;
  }
  e->ev_used = 0;
  return 0;
}

int al_encoder_destroy(al_encoder* e) {
  if (e)
    for (cudaEvent_t ev : e->ev) cudaEventDestroy(ev);
  delete e;
  return 0;
}

// ----------------------------------------------------------------------------- projector
int al_projector_forward(const void* x, int rows, int d_in, int hidden, int d_out, const void* W1, const float* b1,
                         const void* W2, const float* b2, const float* gamma, const float* beta, void* h_ws,
                         float* y_ws, void* out, int out_dtype, long long out_ld, int rows_per_group,
                         long long out_group_stride, long long out_row_offset, al_stream_t stream) {
  AL_REQUIRE(x && W1 && b1 && W2 && b2 && gamma && beta && h_ws && y_ws && out, "al_projector_forward: NULL argument");
  int rc = al_gemm_bf16(x, d_in, (long long)rows * d_in, rows, 1, W1, hidden, d_in, b1, h_ws, hidden,
                        (long long)rows * hidden, AL_EPI_GELU, nullptr, 0, nullptr, stream);
  if (rc) return rc;
  rc = al_gemm_bf16(h_ws, hidden, (long long)rows * hidden, rows, 1, W2, d_out, hidden, b2, y_ws, d_out,
                    (long long)rows * d_out, AL_EPI_OUT_F32, nullptr, 0, nullptr, stream);
  if (rc) return rc;
  return al_layernorm(y_ws, gamma, beta, out, rows, d_out, 1e-5f, out_dtype, out_ld, rows_per_group, out_group_stride,
                      out_row_offset, stream);
}

// ----------------------------------------------------------------------------- projector backward
static int gemm_plain(const void* A, long long lda, int M, const void* W, long long ldw, int N, int K, const float* bias,
                      void* out, long long ldo, int flags, const void* grad_in, long long grad_ld, int want_items,
                      cudaStream_t st) {
  // A [M][lda] (K valid columns), W [N][ldw], out [M][ldo]; optional split-K for small outputs
  CUtensorMap ta, tb, to;
  int rc;
  if ((rc = tmap_rows3d(&ta, A, 2, K, M, 1, lda, (uint64_t)M * lda, 64, 128))) return rc;
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t str[2] = {2, (uint64_t)ldw * 2};
    const uint32_t box[2] = {64, 128};
    if ((rc = make_tmap(&tb, W, 2, 2, dims, str, box, true))) return rc;
  }
  const int oe = (flags & EPI_OUT_F32) ? 4 : 2;
  if ((rc = tmap_rows3d(&to, out, oe, N, M, 1, ldo, (uint64_t)M * ldo, gemm_out_box_cols(flags), 128))) return rc;
  GemmParams p{};
  p.m_per_batch = M; p.batch = 1; p.N = N; p.K = K; p.bias = bias;
  p.grad_in = reinterpret_cast<const __nv_bfloat16*>(grad_in);
  p.grad_ld = grad_ld;
  if (want_items > 0 && (flags & EPI_REDUCE_ADD)) {
    const int out_tiles = ((M + 255) / 256) * ((N + 255) / 256);
    const int num_kb = (K + 63) / 64;
    int ks = (want_items + out_tiles - 1) / out_tiles;
    if (ks > num_kb) ks = num_kb;
    if (ks < 1) ks = 1;
    const int per = (num_kb + ks - 1) / ks;
    p.k_splits = (num_kb + per - 1) / per;          // every slice non-empty
  }
  if (flags & EPI_ADD_BF16) {                        // grad_in = the bf16 addend [M][grad_ld]: brought in by TMA (output map's box)
    CUtensorMap tadd = to;
    if (grad_in != out && (rc = tmap_rows3d(&tadd, grad_in, 2, N, M, 1, grad_ld, (uint64_t)M * grad_ld, gemm_out_box_cols(flags), 128)))
      return rc;
    p.K2 = 0;
    rc = launch_gemm2_add(ta, tb, to, ta, tb, tadd, p, flags, num_sms(), st);
  } else {
    rc = launch_gemm(ta, tb, to, p, flags, num_sms(), st);
  }
  if (rc == 0) g_launches += 1;
  return rc;
}

// out[M][N] (fp32) += A_src^T W_src with A_src [K][lda] (M valid columns) and W_src [K][ldw] (N valid columns), both
// row-major over the contraction index (weight gradients: the contraction runs over the token rows). The operands are
// staged MN-major (EPI_TN), so nothing is transposed in memory; split-K with TMA reduce-add into the zeroed output.
static int gemm_tn(const void* A_src, long long lda, int M, const void* W_src, long long ldw, int N, int K, float* out,
                   long long ldo, int want_items, cudaStream_t st) {
  CUtensorMap ta, tb, to;
  int rc;
  {
    const uint64_t dims[2] = {(uint64_t)M, (uint64_t)K};
    const uint64_t str[2] = {2, (uint64_t)lda * 2};
    const uint32_t box[2] = {64, 64};
    if ((rc = make_tmap(&ta, A_src, 2, 2, dims, str, box, true))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)K};
    const uint64_t str[2] = {2, (uint64_t)ldw * 2};
    const uint32_t box[2] = {64, 64};
    if ((rc = make_tmap(&tb, W_src, 2, 2, dims, str, box, true))) return rc;
  }
  const int flags = EPI_TN | EPI_OUT_F32 | EPI_REDUCE_ADD;
  if ((rc = tmap_rows3d(&to, out, 4, N, M, 1, ldo, (uint64_t)M * ldo, gemm_out_box_cols(flags), 128))) return rc;
  GemmParams p{};
  p.m_per_batch = M; p.batch = 1; p.N = N; p.K = K;
  const int out_tiles = ((M + 255) / 256) * ((N + 255) / 256);
  const int num_kb = (K + 63) / 64;
  int ks = (want_items + out_tiles - 1) / out_tiles;
  if (ks > num_kb) ks = num_kb;
  if (ks < 1) ks = 1;
  const int per = (num_kb + ks - 1) / ks;
  p.k_splits = (num_kb + per - 1) / per;
  rc = launch_gemm(ta, tb, to, p, flags, num_sms(), st);
  if (rc == 0) g_launches += 1;
  return rc;
}

int al_gemm_tn_accumulate(const void* A_src, long long lda, int M, const void* W_src, long long ldw, int N, int K, float* out,
                          long long ldo, al_stream_t stream) {
  AL_REQUIRE(A_src && W_src && out, "al_gemm_tn_accumulate: NULL argument");
  AL_REQUIRE(M > 0 && N > 0 && K > 0 && lda % 8 == 0 && ldw % 8 == 0 && ldo % 4 == 0 && lda >= M && ldw >= N && ldo >= N,
             "al_gemm_tn_accumulate: bad shape M=%d N=%d K=%d lda=%lld ldw=%lld ldo=%lld", M, N, K, lda, ldw, ldo);
  return gemm_tn(A_src, lda, M, W_src, ldw, N, K, out, ldo, 2 * num_sms(), (cudaStream_t)stream);
}

size_t al_projector_backward_workspace_bytes(int rows, int d_in, int hidden, int d_out) {
  const size_t rp = (size_t)(rows + 7) / 8 * 8;
  size_t b = 0;
  auto add = [&](size_t n) { b += (n + 1023) / 1024 * 1024; };
  add((size_t)rows * d_out * 2);   // dy
  add(rp * d_out * 2);             // dyT
  add(rp * hidden * 2);            // hT / daT
  add((size_t)hidden * d_out * 2); // W2T
  add((size_t)rows * hidden * 2);  // dh
  add((size_t)rows * hidden * 2);  // da
  add(rp * d_in * 2);              // xT
  return b;
}

int al_projector_backward(const void* x, int rows, int d_in, int hidden, int d_out, const void* W1, const float* b1,
                          const void* W2, const float* gamma, const void* h_saved, const float* y_saved,
                          const float* dout, void* workspace, float* dW1, float* db1, float* dW2, float* db2,
                          float* dgamma, float* dbeta, al_stream_t stream) {
  AL_REQUIRE(x && W1 && b1 && W2 && gamma && h_saved && y_saved && dout && workspace && dW1 && db1 && dW2 && db2 && dgamma && dbeta,
             "al_projector_backward: NULL argument");
  AL_REQUIRE(rows > 0 && d_in % 8 == 0 && hidden % 8 == 0 && d_out % 8 == 0, "al_projector_backward: bad shape");
  AL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "al_projector_backward: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int rp = (rows + 7) / 8 * 8;
  uint8_t* w = (uint8_t*)workspace;
  auto take = [&](size_t n) { void* p = w; w += (n + 1023) / 1024 * 1024; return p; };
  void* dy = take((size_t)rows * d_out * 2);
  (void)take((size_t)rp * d_out * 2);   // (formerly dy^T / h^T: the weight-gradient GEMMs stage their operands MN-major now;
  (void)take((size_t)rp * hidden * 2);  //  the workspace layout is kept so existing callers' sizes stay valid)
  void* W2T = take((size_t)hidden * d_out * 2);
  void* dh = take((size_t)rows * hidden * 2);
  void* da = take((size_t)rows * hidden * 2);
  (void)take((size_t)rp * d_in * 2);
  AL_CHECK_CUDA(cudaMemsetAsync(dW1, 0, (size_t)hidden * d_in * 4, st));
  AL_CHECK_CUDA(cudaMemsetAsync(db1, 0, (size_t)hidden * 4, st));
  AL_CHECK_CUDA(cudaMemsetAsync(dW2, 0, (size_t)d_out * hidden * 4, st));
  AL_CHECK_CUDA(cudaMemsetAsync(db2, 0, (size_t)d_out * 4, st));
  AL_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, (size_t)d_out * 4, st));
  AL_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, (size_t)d_out * 4, st));
  int rc;
#define STEP(expr) do { rc = (expr); if (rc) return rc; g_launches += 1; } while (0)
  // 1. LayerNorm backward: dy (bf16), dgamma, dbeta, db2 = colsum(dy)
  STEP(launch_layernorm_bwd(y_saved, dout, gamma, dy, dgamma, dbeta, db2, rows, d_out, 1e-5f, st));
  // 2. dW2 = dy^T h : contraction over the rows, operands staged MN-major (no transposes), split-K reduce-add
  if ((rc = gemm_tn(dy, d_out, d_out, h_saved, hidden, hidden, rows, dW2, hidden, 2 * num_sms(), st))) return rc;
  // 3. dh = dy W2
  STEP(launch_transpose_bf16(W2, W2T, d_out, hidden, d_out, st));
  if ((rc = gemm_plain(dy, d_out, rows, W2T, d_out, hidden, d_out, nullptr, dh, hidden, 0, nullptr, 0, 0, st))) return rc;
  // 4. da = dh * gelu'(x W1^T + b1)  (pre-activation recomputed inside the GEMM)
  if ((rc = gemm_plain(x, d_in, rows, W1, d_in, hidden, d_in, b1, da, hidden, EPI_GELU_GRAD, dh, hidden, 0, st))) return rc;
  // 5. db1 = colsum(da); dW1 = da^T x
  STEP(launch_colsum_bf16(da, db1, rows, hidden, st));
  if ((rc = gemm_tn(da, hidden, hidden, x, d_in, d_in, rows, dW1, d_in, 2 * num_sms(), st))) return rc;
#undef STEP
  return 0;
}

// ----------------------------------------------------------------------------- LoRA linear
int al_lora_pack(const float* lora_A, const float* lora_B, int rank, int in_dim, int out_dim, float scaling, void* a_pad,
                 void* b_scaled_pad, al_stream_t stream) {
  AL_REQUIRE(lora_A && lora_B && a_pad && b_scaled_pad, "al_lora_pack: NULL argument");
  AL_REQUIRE(rank > 0 && in_dim > 0 && out_dim > 0, "al_lora_pack: bad shape rank=%d in=%d out=%d", rank, in_dim, out_dim);
  int rc = launch_lora_pack(lora_A, lora_B, rank, (rank + 7) / 8 * 8, in_dim, out_dim, scaling, a_pad, b_scaled_pad,
                            (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

int al_linear_add_bf16(const void* x, int rows, int in_dim, int out_dim, const void* W, const float* bias, const void* addend,
                       void* out, al_stream_t stream) {
  AL_REQUIRE(x && W && out && addend, "al_linear_add_bf16: NULL argument");
  AL_REQUIRE(rows > 0 && in_dim % 8 == 0 && out_dim % 8 == 0, "al_linear_add_bf16: bad shape rows=%d in=%d out=%d", rows, in_dim, out_dim);
  return gemm_plain(x, in_dim, rows, W, in_dim, out_dim, in_dim, bias, out, out_dim, EPI_ADD_BF16, addend, out_dim, 0,
                    (cudaStream_t)stream);
}

int al_lora_linear_forward(const void* x, int rows, int in_dim, int out_dim, int rank, const void* W, const float* bias,
                           const void* lora_A, const void* lora_B_scaled, void* t_ws, void* out, int out_dtype,
                           al_stream_t stream) {
  return al_lora_linear_forward_ex(x, rows, in_dim, out_dim, rank, W, bias, lora_A, lora_B_scaled, t_ws, nullptr, out, out_dtype,
                                   stream);
}
int al_lora_linear_forward_ex(const void* x, int rows, int in_dim, int out_dim, int rank, const void* W, const float* bias,
                              const void* lora_A, const void* lora_B_scaled, void* t_ws, const void* addend, void* out,
                              int out_dtype, al_stream_t stream) {
  AL_REQUIRE(x && W && lora_A && lora_B_scaled && t_ws && out, "al_lora_linear_forward: NULL argument");
  AL_REQUIRE(addend == nullptr || (out_dtype == 0 && out_dim % 8 == 0),
             "al_lora_linear_forward_ex: an addend needs a bf16 output and out_dim %% 8 == 0 (out_dim=%d)", out_dim);
  AL_REQUIRE(rows > 0 && in_dim % 8 == 0 && rank % 8 == 0 && rank > 0 && out_dim > 0,
             "al_lora_linear_forward: bad shape rows=%d in=%d out=%d rank=%d", rows, in_dim, out_dim, rank);
  // 1. T = x A^T  [rows, rank] (bf16)
  int rc = al_gemm_bf16(x, in_dim, (long long)rows * in_dim, rows, 1, lora_A, rank, in_dim, nullptr, t_ws, rank,
                        (long long)rows * rank, 0, nullptr, 0, nullptr, stream);
  if (rc) return rc;
  // 2. out = x W^T + T (s B)^T + bias: the rank-r product rides in the frozen GEMM's TMEM accumulator
  CUtensorMap ta, tb, ta2, tb2, to;
  if ((rc = tmap_rows3d(&ta, x, 2, in_dim, rows, 1, in_dim, (uint64_t)rows * in_dim, 64, 128))) return rc;
  if ((rc = tmap_weight(&tb, W, out_dim, in_dim))) return rc;
  if ((rc = tmap_rows3d(&ta2, t_ws, 2, rank, rows, 1, rank, (uint64_t)rows * rank, 64, 128))) return rc;
  if ((rc = tmap_weight(&tb2, lora_B_scaled, out_dim, rank))) return rc;
  const int flags = out_dtype == 1 ? AL_EPI_OUT_F32 : (addend ? EPI_ADD_BF16 : 0);
  if ((rc = tmap_rows3d(&to, out, out_dtype == 1 ? 4 : 2, out_dim, rows, 1, out_dim, (uint64_t)rows * out_dim,
                        gemm_out_box_cols(flags), 128)))
    return rc;
  GemmParams p{};
  p.m_per_batch = rows; p.batch = 1; p.N = out_dim; p.K = in_dim; p.K2 = rank; p.bias = bias;
  CUtensorMap tadd = to;                             // the addend chunk arrives through TMA with the output map's box
  if (addend != nullptr && addend != out &&
      (rc = tmap_rows3d(&tadd, addend, 2, out_dim, rows, 1, out_dim, (uint64_t)rows * out_dim, gemm_out_box_cols(flags), 128)))
    return rc;
  rc = launch_gemm2_add(ta, tb, to, ta2, tb2, tadd, p, flags, num_sms(), (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}

size_t al_lora_linear_backward_workspace_bytes(int rows, int in_dim, int out_dim, int rank) {
  const size_t rp = (size_t)(rows + 7) / 8 * 8;
  size_t b = 0;
  auto add = [&](size_t n) { b += (n + 1023) / 1024 * 1024; };
  add((size_t)rank * out_dim * 2);   // (sB)^T
  add((size_t)in_dim * rank * 2);    // A^T
  add((size_t)rows * rank * 2);      // U = dy (sB)
  (void)rp;
  return b;
}

// Backward of out = x W^T + T (sB)^T, T = x A^T, W frozen:
//   U = dy (sB)              [rows][rank]
//   dx = dy W + U A          one GEMM, second operand pair in the K loop (same TMEM accumulator)
//   dA = U^T x               [rank][in]   fp32, split-K reduce-add
//   dB_raw = dy^T T          [out][rank]  fp32, split-K reduce-add (gradient of the UNSCALED B = scaling * dB_raw)
int al_lora_linear_backward(const void* x, const void* dy, int rows, int in_dim, int out_dim, int rank, const void* W_T,
                            const void* lora_A, const void* lora_B_scaled, const void* t_saved, void* workspace, void* dx,
                            float* dA, float* dB_raw, al_stream_t stream) {
  return al_lora_linear_backward_ex(x, dy, rows, in_dim, out_dim, rank, W_T, lora_A, lora_B_scaled, t_saved, workspace, nullptr, dx,
                                    dA, dB_raw, stream);
}
int al_lora_linear_backward_ex(const void* x, const void* dy, int rows, int in_dim, int out_dim, int rank, const void* W_T,
                               const void* lora_A, const void* lora_B_scaled, const void* t_saved, void* workspace,
                               const void* dx_addend, void* dx, float* dA, float* dB_raw, al_stream_t stream) {
  AL_REQUIRE(x && dy && lora_A && lora_B_scaled && t_saved && workspace && dA && dB_raw,
             "al_lora_linear_backward: NULL argument");
  AL_REQUIRE(dx == nullptr || W_T != nullptr, "al_lora_linear_backward: dx needs W_T ([in][out], the frozen weight transposed)");
  AL_REQUIRE(rows > 0 && in_dim % 8 == 0 && out_dim % 8 == 0 && rank % 8 == 0 && rank > 0,
             "al_lora_linear_backward: bad shape rows=%d in=%d out=%d rank=%d", rows, in_dim, out_dim, rank);
  AL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "al_lora_linear_backward: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = (uint8_t*)workspace;
  auto take = [&](size_t n) { void* p = w; w += (n + 1023) / 1024 * 1024; return p; };
  void* sBT = take((size_t)rank * out_dim * 2);
  void* AT = take((size_t)in_dim * rank * 2);
  void* U = take((size_t)rows * rank * 2);

  if (dB_raw == dA + (size_t)rank * in_dim) {          // one allocation holding both (audio_llama_b200/ops.py): one memset
    AL_CHECK_CUDA(cudaMemsetAsync(dA, 0, ((size_t)rank * in_dim + (size_t)out_dim * rank) * 4, st));
  } else {
    AL_CHECK_CUDA(cudaMemsetAsync(dA, 0, (size_t)rank * in_dim * 4, st));
    AL_CHECK_CUDA(cudaMemsetAsync(dB_raw, 0, (size_t)out_dim * rank * 4, st));
  }
  int rc;
#define STEP(expr) do { rc = (expr); if (rc) return rc; g_launches += 1; } while (0)
  // 1. U = dy (sB): the weight operand is (sB)^T [rank][out]
  STEP(launch_transpose_bf16(lora_B_scaled, sBT, out_dim, rank, out_dim, st));
  if ((rc = gemm_plain(dy, out_dim, rows, sBT, out_dim, rank, out_dim, nullptr, U, rank, 0, nullptr, 0, 0, st))) return rc;
  // 2. dx = dy W + U A
  if (dx != nullptr) {
    STEP(launch_transpose_bf16(lora_A, AT, rank, in_dim, rank, st));
    CUtensorMap ta, tb, ta2, tb2, to;
    if ((rc = tmap_rows3d(&ta, dy, 2, out_dim, rows, 1, out_dim, (uint64_t)rows * out_dim, 64, 128))) return rc;
    if ((rc = tmap_weight(&tb, W_T, in_dim, out_dim))) return rc;
    if ((rc = tmap_rows3d(&ta2, U, 2, rank, rows, 1, rank, (uint64_t)rows * rank, 64, 128))) return rc;
    if ((rc = tmap_weight(&tb2, AT, in_dim, rank))) return rc;
    if ((rc = tmap_rows3d(&to, dx, 2, in_dim, rows, 1, in_dim, (uint64_t)rows * in_dim, gemm_out_box_cols(0), 128))) return rc;
    GemmParams p{};
    p.m_per_batch = rows; p.batch = 1; p.N = in_dim; p.K = out_dim; p.K2 = rank;
    CUtensorMap tadd = to;                           // dx = ... + dx_addend (may be dx itself: then the output map serves)
    if (dx_addend != nullptr && dx_addend != dx &&
        (rc = tmap_rows3d(&tadd, dx_addend, 2, in_dim, rows, 1, in_dim, (uint64_t)rows * in_dim, gemm_out_box_cols(0), 128)))
      return rc;
    rc = launch_gemm2_add(ta, tb, to, ta2, tb2, tadd, p, dx_addend ? EPI_ADD_BF16 : 0, num_sms(), st);
    if (rc) return rc;
    g_launches += 1;
  }
  // 3. dA = U^T x  (contraction over the rows: operands staged MN-major, nothing transposed; split-K reduce-add)
  if ((rc = gemm_tn(U, rank, rank, x, in_dim, in_dim, rows, dA, in_dim, 2 * num_sms(), st))) return rc;
  // 4. dB_raw = dy^T T
  if ((rc = gemm_tn(dy, out_dim, out_dim, t_saved, rank, rank, rows, dB_raw, rank, 2 * num_sms(), st))) return rc;
#undef STEP
  return 0;
}

// ----------------------------------------------------------------------------- LLaMA-side row kernels (§8f-1 slice)
int al_rmsnorm_forward(const void* x, const void* weight, void* y, float* rstd, int rows, int d, float eps, al_stream_t stream) {
  AL_REQUIRE(x && weight && y && rows >= 0, "al_rmsnorm_forward: NULL argument");
  int rc = launch_rmsnorm(x, weight, y, rstd, nullptr, nullptr, nullptr, rows, d, eps, false, (cudaStream_t)stream);
  if (rc == 0 && rows > 0) g_launches += 1;
  return rc;
}
int al_rmsnorm_backward(const void* x, const void* weight, const float* rstd, const void* dy, void* dx, int rows, int d,
                        al_stream_t stream) {
  return al_rmsnorm_backward_ex(x, weight, rstd, dy, nullptr, dx, rows, d, stream);
}
int al_rmsnorm_backward_ex(const void* x, const void* weight, const float* rstd, const void* dy, const void* dx_addend, void* dx,
                           int rows, int d, al_stream_t stream) {
  AL_REQUIRE(x && weight && rstd && dy && dx && rows >= 0, "al_rmsnorm_backward: NULL argument");
  int rc = launch_rmsnorm(x, weight, nullptr, const_cast<float*>(rstd), dy, dx_addend, dx, rows, d, 0.f, true, (cudaStream_t)stream);
  if (rc == 0 && rows > 0) g_launches += 1;
  return rc;
}
int al_swiglu_forward(const void* gate, const void* up, void* h, long long n, al_stream_t stream) {
  AL_REQUIRE(gate && up && h && n >= 0, "al_swiglu_forward: NULL argument");
  int rc = launch_swiglu(gate, up, nullptr, h, nullptr, n, false, num_sms(), (cudaStream_t)stream);
  if (rc == 0 && n > 0) g_launches += 1;
  return rc;
}
int al_swiglu_backward(const void* gate, const void* up, const void* dh, void* dgate, void* dup, long long n, al_stream_t stream) {
  AL_REQUIRE(gate && up && dh && dgate && dup && n >= 0, "al_swiglu_backward: NULL argument");
  int rc = launch_swiglu(gate, up, dh, dgate, dup, n, true, num_sms(), (cudaStream_t)stream);
  if (rc == 0 && n > 0) g_launches += 1;
  return rc;
}
int al_rope(const void* x, const void* cos, const void* sin, void* out, int B, int S, int H, int head_dim, int cos_batch,
            int backward, al_stream_t stream) {
  AL_REQUIRE(x && cos && sin && out, "al_rope: NULL argument");
  int rc = launch_rope(x, cos, sin, out, B, S, H, head_dim, cos_batch, backward, num_sms(), (cudaStream_t)stream);
  if (rc == 0) g_launches += 1;
  return rc;
}
int al_cross_entropy_inplace(void* logits, const long long* labels, int rows, int vocab, long long ld, float grad_scale,
                             float* loss_sum, al_stream_t stream) {
  AL_REQUIRE(logits && labels && loss_sum && rows >= 0 && vocab > 0, "al_cross_entropy_inplace: bad argument");
  int rc = launch_ce_inplace(logits, labels, rows, vocab, ld, grad_scale, loss_sum, (cudaStream_t)stream);
  if (rc == 0 && rows > 0) g_launches += 1;
  return rc;
}

size_t al_linear_ce_workspace_bytes(int chunk_rows, int vocab) {
  const size_t ldv = ((size_t)vocab + 7) / 8 * 8;
  return (size_t)chunk_rows * ldv * 2 + 1024;
}
// lm_head + cross-entropy without materialising the [rows][vocab] logits: per chunk of rows, logits = h W^T (bf16),
// the cross-entropy overwrites them with (softmax - onehot) * grad_scale, and dh = dlogits W follows at once.
int al_linear_ce(const void* h, const void* W, const void* W_T, const long long* labels, int rows, int d, int vocab,
                 float grad_scale, int chunk_rows, void* workspace, float* loss_sum, void* dh, al_stream_t stream) {
  AL_REQUIRE(h && W && labels && workspace && loss_sum, "al_linear_ce: NULL argument");
  AL_REQUIRE(dh == nullptr || W_T != nullptr, "al_linear_ce: dh needs W_T ([d][round8(vocab)], lm_head transposed)");
  AL_REQUIRE(rows >= 0 && d % 8 == 0 && vocab > 0 && chunk_rows > 0, "al_linear_ce: bad shape rows=%d d=%d vocab=%d", rows, d, vocab);
  AL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "al_linear_ce: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long ldv = ((long long)vocab + 7) / 8 * 8;
  AL_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(float), st));
  int rc;
  for (int r0 = 0; r0 < rows; r0 += chunk_rows) {
    const int m = rows - r0 < chunk_rows ? rows - r0 : chunk_rows;
    const uint8_t* hc = (const uint8_t*)h + (size_t)r0 * d * 2;
    if ((rc = gemm_plain(hc, d, m, W, d, vocab, d, nullptr, workspace, ldv, 0, nullptr, 0, 0, st))) return rc;
    if ((rc = launch_ce_inplace(workspace, labels + r0, m, vocab, ldv, grad_scale, loss_sum, st))) return rc;
    g_launches += 1;
    if (dh != nullptr) {
      uint8_t* dc = (uint8_t*)dh + (size_t)r0 * d * 2;
      if ((rc = gemm_plain(workspace, ldv, m, W_T, ldv, d, vocab, nullptr, dc, d, 0, nullptr, 0, 0, st))) return rc;
    }
  }
  return 0;
}

// ----------------------------------------------------------------------------- splice
int al_splice(const void* table, int elem_bytes, int d, const long long* input_ids, const long long* attn_mask,
              const long long* labels, int B, int t_txt, int n_audio, long long start_id, long long end_id,
              const void* audio_rows, void* out, float* mask_out, long long* labels_out, long long vocab,
              int* bad_id_flag, al_stream_t stream) {
  AL_REQUIRE(table && input_ids && out, "al_splice: NULL argument");
  AL_REQUIRE(vocab > 0, "al_splice: vocab must be positive (rows of the embedding table)");
  AL_REQUIRE(start_id < vocab && end_id < vocab, "Token IDs %lld, %lld are outside vocabulary size %lld", start_id, end_id, vocab);
  AL_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "al_splice: elem_bytes must be 2 or 4");
  AL_REQUIRE(B >= 0 && t_txt >= 0 && n_audio >= 0, "al_splice: negative size");
  AL_REQUIRE(start_id >= 0 && end_id >= 0, "al_splice: negative delimiter id");
  int rc = launch_splice(table, elem_bytes, d, input_ids, attn_mask, labels, B, t_txt, n_audio, start_id, end_id,
                         audio_rows, out, mask_out, labels_out, vocab, bad_id_flag, (cudaStream_t)stream);
  if (rc == 0 && B > 0) g_launches += 1;
  return rc;
}

int al_splice_ragged(const void* table, int elem_bytes, int d, const long long* input_ids,
                     const long long* attn_mask, const long long* labels, int B, int t_txt, int S_out,
                     const int* span_rows, const int* span_src_row, const int* n_spans, int max_spans,
                     const void* audio_rows, long long start_id, long long end_id, void* out, float* mask_out,
                     long long* labels_out, int* span_start_out, long long vocab, int* bad_id_flag,
                     al_stream_t stream) {
  AL_REQUIRE(table && input_ids && out && span_rows && span_src_row && n_spans && audio_rows,
             "al_splice_ragged: NULL argument");
  AL_REQUIRE(vocab > 0 && start_id >= 0 && end_id >= 0 && start_id < vocab && end_id < vocab,
             "Token IDs %lld, %lld are outside vocabulary size %lld", start_id, end_id, vocab);
  AL_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "al_splice_ragged: elem_bytes must be 2 or 4");
  int rc = launch_splice_ragged(table, elem_bytes, d, input_ids, attn_mask, labels, B, t_txt, S_out, span_rows,
                                span_src_row, n_spans, max_spans, audio_rows, start_id, end_id, out, mask_out,
                                labels_out, span_start_out, vocab, bad_id_flag, (cudaStream_t)stream);
  if (rc == 0 && B > 0) g_launches += 1;
  return rc;
}

}  // extern "C"
