// Internal declarations shared by the kernels behind include/mintime_b200.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mintime_b200.h"

struct CUtensorMap_st;

namespace mt {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------ host-side error reporting
void set_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);

#define MT_REQUIRE(cond, ...)     \
  do {                            \
    if (!(cond)) {                \
      mt::set_error(__VA_ARGS__); \
      return MT_ERR_ARG;          \
    }                             \
  } while (0)

// Function attributes (dynamic shared-memory opt-in, carve-out) are PER DEVICE: true the first time `key` (a kernel's
// address) is seen on the calling thread's current device, so a process that touches a second GPU opts in there too.
bool first_use_on_device(const void* key);
int current_sms();     // multiprocessor count of the current device

// ------------------------------------------------------------------ diagnostics (mt_prof_* in the C ABI)
void count_launch();
// When profiling is enabled, brackets the launches issued in its scope with CUDA events on `stream`
// and books the elapsed time under `name` with the algorithmic flops / bytes of that launch.
struct ProfScope {
  ProfScope(cudaStream_t stream, double flops, double bytes, const char* fmt, ...);
  ~ProfScope();
  int slot;
  cudaStream_t st;
};

// effnet.cu: plain 3x3 s1 p1 depthwise convolution (bf16 NHWC, optional ReLU on the input) on the TMA-staged FFMA2 kernel
int launch_dw_plain_bf16(const void* in, const float* w, void* out, int n_img, int H, int C, int relu_in, cudaStream_t st);

#define MT_LAUNCH_CHECK(what)                                   \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return mt::cuda_status(e__, what);  \
    mt::count_launch();                                         \
  } while (0)

// ------------------------------------------------------------------ element helpers
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// swish / SiLU: x * sigmoid(x)  (reference utils.py:66-69).  The exact path uses expf, the bf16 path
// the fast intrinsic (its error is far below bf16 resolution).
// bf16 path: x*sigmoid(x) = h + h*tanh(h), h = x/2 -- one MUFU (tanh.approx, rel. error 2^-11) and two
// FMA-pipe ops instead of ex2 + rcp; the extractor evaluates ~3e9 swishes per 32-clip step, which
// would otherwise be bound by the 16-lane MUFU pipe.
template <bool kExact>
__device__ __forceinline__ float silu(float x) {
  if (kExact) return x / (1.0f + expf(-x));
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// bf16-path swish of (acc + shift) in the ONE form every epilogue of the library uses (scalar or packed FFMA2), so
// that the kernel chosen for a shape (single-CTA / CTA-pair GEMM, fused or stand-alone depthwise) never changes a
// bit of the result: h = fma(acc, 0.5, shift/2), h + h*tanh(h).
__device__ __forceinline__ float swish_shift_fast(float acc, float shift) {
  const float h = fmaf(acc, 0.5f, 0.5f * shift);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
template <bool kExact>
__device__ __forceinline__ float sigmoidf_(float x) {
  if (kExact) return 1.0f / (1.0f + expf(-x));
  return __fdividef(1.0f, 1.0f + __expf(-x));
}
// F.gelu default (erf form), size_invariant_timesformer.py:63
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

// Packed GEGLU for a pair of columns: u * gelu_erf(g), the Abramowitz-Stegun 7.1.26 erf (|error| < 1.5e-7) on
// float2 operands (FFMA2 / FMUL2: half the issue slots of the scalar form; the MUFU count -- one rcp and one ex2
// per gate value -- is unchanged).  The polynomial carries the minus sign so that r = 1 + (-p t) e is one FFMA2.
__device__ __forceinline__ float2 geglu2(float2 u, float2 g) {
  const float2 z = __fmul2_rn(g, make_float2(0.70710678118654752f, 0.70710678118654752f));
  const float2 az = make_float2(fabsf(z.x), fabsf(z.y));
  const float2 den = __ffma2_rn(az, make_float2(0.3275911f, 0.3275911f), make_float2(1.0f, 1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  float2 p = __ffma2_rn(make_float2(-1.061405429f, -1.061405429f), t, make_float2(1.453152027f, 1.453152027f));
  p = __ffma2_rn(p, t, make_float2(-1.421413741f, -1.421413741f));
  p = __ffma2_rn(p, t, make_float2(0.284496736f, 0.284496736f));
  p = __ffma2_rn(p, t, make_float2(-0.254829592f, -0.254829592f));
  const float2 npt = __fmul2_rn(p, t);                                           // -(poly * t)
  const float2 q = __fmul2_rn(__fmul2_rn(az, make_float2(-1.4426950408889634f, -1.4426950408889634f)), az);
  float2 ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(q.x));                     // exp(-z^2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(q.y));
  const float2 r = __ffma2_rn(npt, ex, make_float2(1.0f, 1.0f));                 // erf(|z|)
  const float2 erf2 = make_float2(copysignf(r.x, z.x), copysignf(r.y, z.y));
  const float2 h = __fmul2_rn(g, make_float2(0.5f, 0.5f));
  return __fmul2_rn(u, __ffma2_rn(h, erf2, h));                                  // u * 0.5 g (1 + erf)
}

// Packed derivative pieces of GELU for the GEGLU backward (bf16 path): cdf = Phi(g) = (1 + erf(g / sqrt 2)) / 2 and
// pdf = phi(g) = exp(-g^2 / 2) / sqrt(2 pi) from ONE evaluation of the same Abramowitz-Stegun erf as geglu2 -- its exponential
// exp(-z^2), z = g / sqrt 2, is exp(-g^2 / 2).
__device__ __forceinline__ void gelu_cdf_pdf2(float2 g, float2& cdf, float2& pdf) {
  const float2 z = __fmul2_rn(g, make_float2(0.70710678118654752f, 0.70710678118654752f));
  const float2 az = make_float2(fabsf(z.x), fabsf(z.y));
  const float2 den = __ffma2_rn(az, make_float2(0.3275911f, 0.3275911f), make_float2(1.0f, 1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  float2 p = __ffma2_rn(make_float2(-1.061405429f, -1.061405429f), t, make_float2(1.453152027f, 1.453152027f));
  p = __ffma2_rn(p, t, make_float2(-1.421413741f, -1.421413741f));
  p = __ffma2_rn(p, t, make_float2(0.284496736f, 0.284496736f));
  p = __ffma2_rn(p, t, make_float2(-0.254829592f, -0.254829592f));
  const float2 npt = __fmul2_rn(p, t);
  const float2 q = __fmul2_rn(__fmul2_rn(az, make_float2(-1.4426950408889634f, -1.4426950408889634f)), az);
  float2 ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(q.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(q.y));
  const float2 r = __ffma2_rn(npt, ex, make_float2(1.0f, 1.0f));                 // erf(|z|)
  const float2 erf2 = make_float2(copysignf(r.x, z.x), copysignf(r.y, z.y));
  cdf = __ffma2_rn(erf2, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
  pdf = __fmul2_rn(ex, make_float2(0.39894228040143268f, 0.39894228040143268f));
}

// 8 consecutive elements <-> registers
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ------------------------------------------------------------------ GEMM epilogues
enum EpiKind { EPI_STORE = 0, EPI_RESID_F32 = 1, EPI_GEGLU = 2, EPI_PATCH_EMBED = 3 };

struct EpiParams {
  int kind;
  int M, N;            // GEMM extents (N counts weight rows, i.e. 2x the output width for GEGLU)
  const float* bias;   // [N] or null
  int act;             // EPI_STORE: 0 none, 1 swish
  const void* resid;   // EPI_STORE: T [M][ldo] or null (added after act)
  void* out;           // T* (STORE, GEGLU) or float* (RESID_F32, PATCH_EMBED)
  int ldo;             // output row stride (elements)
  // EPI_PATCH_EMBED
  int rows_per_batch;  // f * n
  int n_patches;       // n
  int frames;          // f
  const float* pos_tab;
  const float* size_tab;
  const long long* positions;  // [B][rows_per_batch + 1] or null (-> arange)
  const int* size_idx;         // [B][f]
  int table_rows;              // rows of pos_tab / size_tab: an index outside [0, table_rows) traps the kernel (nn.Embedding
                               // raises IndexError in the reference; silently reading out of bounds is not an option)
};

struct GemmArgs {
  const void* a;   // T [M][K]
  const void* w;   // T [N][K]
  int M, N, K;
  const float* gate;  // f32 [M / rows_per_gate][K] or null
  int rows_per_gate;
  EpiParams epi;
  int splits;         // > 1: split-K request (honoured for EPI_RESID_F32 without bias on the tensor-core path)
  int mn_major;       // 1: a is T [K][M] and w is T [K][N] (contraction index slow): dW = dY^T X without transposed copies
};

// One group of 8 consecutive output columns [col, col+8) of row `row` (both already bounds-checked).
template <typename T, int KIND>
__device__ __forceinline__ void epi_store8(const EpiParams& p, int row, int col, float (&v)[8]) {
  constexpr bool kExact = sizeof(T) == 4;
  float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (p.bias) load8(p.bias + col, b);
  if (KIND == EPI_STORE && p.act == 1 && !kExact) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = swish_shift_fast(v[i], b[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += b[i];
  }
  if (KIND == EPI_STORE) {
    if (p.act == 1 && kExact) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = silu<true>(v[i]);
    }
    const size_t off = (size_t)row * p.ldo + col;
    if (p.resid) {
      float r[8];
      load8(reinterpret_cast<const T*>(p.resid) + off, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += r[i];
    }
    store8(reinterpret_cast<T*>(p.out) + off, v);
  } else if (KIND == EPI_RESID_F32) {
    float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col;
    float r[8];
    load8(o, r);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += r[i];
    store8(o, v);
  } else if (KIND == EPI_PATCH_EMBED) {
    const int b = row / p.rows_per_batch, t = row - b * p.rows_per_batch;
    const size_t orow = (size_t)row + b + 1;
    const long long pos = p.positions ? p.positions[(size_t)b * (p.rows_per_batch + 1) + 1 + t] : (long long)(1 + t);
    if ((unsigned long long)pos >= (unsigned long long)p.table_rows) __trap();
    float e[8];
    load8(p.pos_tab + (size_t)pos * p.ldo + col, e);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += e[i];
    if (p.size_tab) {
      const int si = p.size_idx[b * p.frames + t / p.n_patches];
      if ((unsigned)si >= (unsigned)p.table_rows) __trap();
      load8(p.size_tab + (size_t)si * p.ldo + col, e);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += e[i];
    }
    store8(reinterpret_cast<float*>(p.out) + orow * p.ldo + col, v);
  }
}

// GEGLU: u = 8 packed columns [col, col+8), g = their gate columns [col+32, col+40) of the same
// 64-wide interleave block; writes out[row][ (col/64)*32 + col%64 .. +8 ).
template <typename T>
__device__ __forceinline__ void epi_geglu8(const EpiParams& p, int row, int col, float (&u)[8], float (&g)[8]) {
  if (p.bias) {
    float b[8];
    load8(p.bias + col, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] += b[i];
    load8(p.bias + col + 32, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] += b[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] *= gelu_erf(g[i]);
  const int ocol = (col >> 6) * 32 + (col & 63);
  store8(reinterpret_cast<T*>(p.out) + (size_t)row * p.ldo + ocol, u);
}

// ------------------------------------------------------------------ internal launchers
int launch_gemm(int precision, const GemmArgs& g, cudaStream_t stream);
// 4-D TMA descriptor over a bf16 NHWC tensor: box = 64 channels x box_w x box_h x 1 image, 128-byte swizzle
int make_tmap_nhwc_bf16(CUtensorMap_st* m, const void* base, int n, int h, int w, int c, int box_w, int box_h);
int make_tmap_nhwc_bf16_kmajor(CUtensorMap_st* m, const void* base, int n, int h, int w, int c, int box_c, int box_w,
                               int box_h);
int make_tmap_weights_kmajor(CUtensorMap_st* m, const void* base, int rows, int cols, int box_rows, int box_cols);
int make_tmap_4d_bf16_sw128(CUtensorMap_st* m, const void* base, const unsigned long long dims[4],
                            const unsigned long long strides[3], const unsigned box[4]);

}  // namespace mt
