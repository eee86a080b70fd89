// EfficientNet-B0 feature extractor, eval mode (reference models/efficientnet/efficientnet_pytorch/
// model.py:267-288): stem -> 16 MBConv blocks -> head, all activations NHWC.
// The 1x1 convolutions run in gemm.cu; this file holds the stencil / reduction kernels
// (stem 3x3, depthwise kxk + BN + swish + SE squeeze, SE excitation) and the layer schedule.
#include <float.h>

#include "common.cuh"

namespace mt {
namespace {

// B0 stage table (reference utils.py:502-510 expanded by model.py:171-191): kernel, stride, expand,
// cin, cout, input side.  Mirrors spec.py; tests/test_host_logic.py checks the two agree.
struct BlockSpec { int k, s, e, cin, cout, hw; };
const BlockSpec kBlocks[16] = {
    {3, 1, 1, 32, 16, 112},  {3, 2, 6, 16, 24, 112},  {3, 1, 6, 24, 24, 56},   {5, 2, 6, 24, 40, 56},
    {5, 1, 6, 40, 40, 28},   {3, 2, 6, 40, 80, 28},   {3, 1, 6, 80, 80, 14},   {3, 1, 6, 80, 80, 14},
    {5, 1, 6, 80, 112, 14},  {5, 1, 6, 112, 112, 14}, {5, 1, 6, 112, 112, 14}, {5, 2, 6, 112, 192, 14},
    {5, 1, 6, 192, 192, 7},  {5, 1, 6, 192, 192, 7},  {5, 1, 6, 192, 192, 7},  {3, 1, 6, 192, 320, 7}};

__host__ __device__ inline int same_pad_lo(int in, int k, int s) {
  // Conv2dStaticSamePadding (utils.py:254-269): total = max((ceil(in/s)-1)*s + k - in, 0), low side gets total/2
  const int out = (in + s - 1) / s;
  int total = (out - 1) * s + k - in;
  if (total < 0) total = 0;
  return total / 2;
}

// ---------------------------------------------------------------------------------------------------
// stem: pad(0,1,0,1) + conv3x3 s2 (3 -> 32) + BN + swish   (utils.py:273-276, model.py:276)
// one thread = one output pixel x 8 output channels
// ---------------------------------------------------------------------------------------------------
template <typename T, typename TIN>
__global__ void __launch_bounds__(256) stem_kernel(const TIN* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ shift, T* __restrict__ out, int n_img,
                                                   int H, int W, int Ho, int Wo, int pad_lo) {
  __shared__ float ws[27 * 32];
  __shared__ float sh[32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < 32) sh[threadIdx.x] = shift[threadIdx.x];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n_img * Ho * Wo * 4;
  if (idx >= total) return;
  const int oct = (int)(idx & 3);
  long long pix = idx >> 2;
  const int ox = (int)(pix % Wo);
  pix /= Wo;
  const int oy = (int)(pix % Ho);
  const int img = (int)(pix / Ho);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const TIN* xi = x + (size_t)img * H * W * 3;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - pad_lo;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 + kx - pad_lo;
      if (ix < 0 || ix >= W) continue;
      const TIN* px = xi + ((size_t)iy * W + ix) * 3;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        const float v = (float)px[ci];
        const float* wr = ws + ((ky * 3 + kx) * 3 + ci) * 32 + oct * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(v, wr[i], acc[i]);
      }
    }
  }
  constexpr bool kExact = sizeof(T) == 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = silu<kExact>(acc[i] + sh[oct * 8 + i]);
  store8(out + ((size_t)(img * Ho + oy) * Wo + ox) * 32 + oct * 8, acc);
}

// ---------------------------------------------------------------------------------------------------
// depthwise kxk stride s + BN + swish, and per-(image, pixel chunk, channel) partial sums for the
// squeeze-excite average pool (model.py:105-107,110).  grid = (pixel chunks, 64-channel chunks, images); a warp walks pixels,
// its lanes hold channel pairs so every tap is one coalesced 128-byte (bf16) row segment.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 load2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load2(const bf16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store2(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

constexpr int kDwPixPerBlock = 256;

template <typename T, int K, int S>
__global__ void __launch_bounds__(256) dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                     const float* __restrict__ shift, T* __restrict__ out,
                                                     float* __restrict__ pool_part, int H, int W, int Ho, int Wo, int C,
                                                     int pad_lo) {
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int img = blockIdx.z;
  const int c = blockIdx.y * 64 + lane * 2;
  const bool active = c < C;
  float wk[K * K][2];
  float sh0 = 0.f, sh1 = 0.f;
  if (active) {
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
      const float2 ww = load2(w + (size_t)t * C + c);
      wk[t][0] = ww.x; wk[t][1] = ww.y;
    }
    sh0 = shift[c]; sh1 = shift[c + 1];
  }
  const T* in_img = in + (size_t)img * H * W * C;
  T* out_img = out + (size_t)img * Ho * Wo * C;
  const int p_begin = blockIdx.x * kDwPixPerBlock;
  const int p_end = min(p_begin + kDwPixPerBlock, Ho * Wo);
  float ps0 = 0.f, ps1 = 0.f;
  constexpr bool kExact = sizeof(T) == 4;
  if (active) {
    for (int p = p_begin + warp; p < p_end; p += 8) {
      const int oy = p / Wo, ox = p - oy * Wo;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * S + ky - pad_lo;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int ix = ox * S + kx - pad_lo;
          if (ix < 0 || ix >= W) continue;
          const float2 v = load2(in_img + ((size_t)iy * W + ix) * C + c);
          a0 = fmaf(v.x, wk[ky * K + kx][0], a0);
          a1 = fmaf(v.y, wk[ky * K + kx][1], a1);
        }
      }
      a0 = silu<kExact>(a0 + sh0);
      a1 = silu<kExact>(a1 + sh1);
      store2(out_img + (size_t)p * C + c, a0, a1);
      ps0 += a0; ps1 += a1;
    }
  }
  red[warp][lane * 2] = ps0;
  red[warp][lane * 2 + 1] = ps1;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int cc = blockIdx.y * 64 + threadIdx.x;
    if (cc < C) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
      // one writer per (image, pixel chunk, channel): deterministic, no zero-init needed
      pool_part[((size_t)img * gridDim.x + blockIdx.x) * C + cc] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// SE excitation (model.py:111-115): gate = sigmoid(We * swish(Wr * mean + br) + be); one block per image
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) se_gate_kernel(const float* __restrict__ pool_part, int n_chunks, float inv_hw,
                                                      const float* __restrict__ wr, const float* __restrict__ br,
                                                      const float* __restrict__ we, const float* __restrict__ be,
                                                      float* __restrict__ gate, int C, int SQ) {
  extern __shared__ float sm[];
  float* mean = sm;        // [C]
  float* sq = sm + C;      // [SQ]
  const int img = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double acc = 0.0;
    for (int j = 0; j < n_chunks; ++j) acc += (double)pool_part[((size_t)img * n_chunks + j) * C + c];
    mean[c] = (float)(acc * (double)inv_hw);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int j = warp; j < SQ; j += nwarps) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(wr[(size_t)j * C + c], mean[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sq[j] = silu<true>(s + br[j]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = be[c];
    for (int j = 0; j < SQ; ++j) s = fmaf(we[(size_t)c * SQ + j], sq[j], s);
    gate[(size_t)img * C + c] = sigmoidf_<true>(s);
  }
}

template <typename T, typename TIN>
int launch_stem_t(const void* x, const float* w, const float* shift, void* out, int n_img, int H, int W,
                  cudaStream_t st) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = (long long)n_img * Ho * Wo * 4;
  const int grid = (int)((total + 255) / 256);
  ProfScope prof(st, 2.0 * 27 * 32 * (double)n_img * Ho * Wo,
                 (double)n_img * ((double)H * W * 3 * sizeof(TIN) + (double)Ho * Wo * 32 * sizeof(T)), "stem");
  stem_kernel<T, TIN><<<grid, 256, 0, st>>>(reinterpret_cast<const TIN*>(x), w, shift, reinterpret_cast<T*>(out),
                                            n_img, H, W, Ho, Wo, same_pad_lo(H, 3, 2));
  MT_LAUNCH_CHECK("stem_kernel");
  return MT_OK;
}

template <typename T>
int launch_dw_t(const void* in, const float* w, const float* shift, void* out, float* pool, int n_img, int H, int W,
                int C, int k, int s, cudaStream_t st) {
  const int Ho = (H + s - 1) / s, Wo = (W + s - 1) / s;
  dim3 grid((Ho * Wo + kDwPixPerBlock - 1) / kDwPixPerBlock, (C + 63) / 64, n_img);
  const int pad = same_pad_lo(H, k, s);
  const T* i = reinterpret_cast<const T*>(in);
  T* o = reinterpret_cast<T*>(out);
  ProfScope prof(st, 2.0 * k * k * (double)n_img * Ho * Wo * C,
                 (double)n_img * C * ((double)H * W + (double)Ho * Wo) * sizeof(T), "dwconv k%d s%d C%d H%d", k, s, C, H);
  if (k == 3 && s == 1) dwconv_kernel<T, 3, 1><<<grid, 256, 0, st>>>(i, w, shift, o, pool, H, W, Ho, Wo, C, pad);
  else if (k == 3 && s == 2) dwconv_kernel<T, 3, 2><<<grid, 256, 0, st>>>(i, w, shift, o, pool, H, W, Ho, Wo, C, pad);
  else if (k == 5 && s == 1) dwconv_kernel<T, 5, 1><<<grid, 256, 0, st>>>(i, w, shift, o, pool, H, W, Ho, Wo, C, pad);
  else if (k == 5 && s == 2) dwconv_kernel<T, 5, 2><<<grid, 256, 0, st>>>(i, w, shift, o, pool, H, W, Ho, Wo, C, pad);
  else {
    set_error("dwconv: unsupported kernel %d / stride %d", k, s);
    return MT_ERR_UNSUPPORTED;
  }
  MT_LAUNCH_CHECK("dwconv_kernel");
  return MT_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int dw_chunks(int hw_out) { return (hw_out * hw_out + kDwPixPerBlock - 1) / kDwPixPerBlock; }

struct BlockWs { size_t exp, dw, pool, gate, total; };
BlockWs block_ws_layout(const mt_mbconv_spec_t& b, int n_img, int precision) {
  const size_t es = precision == MT_PREC_FP32 ? 4 : 2;
  const size_t ho = (b.hw_in + b.stride - 1) / b.stride, cexp = (size_t)b.cin * b.expand;
  BlockWs l;
  size_t off = 0;
  l.exp = off;  off += b.expand != 1 ? align_up((size_t)b.hw_in * b.hw_in * cexp * n_img * es, 1024) : 0;
  l.dw = off;   off += align_up(ho * ho * cexp * n_img * es, 1024);
  l.pool = off; off += align_up((size_t)dw_chunks((int)ho) * cexp * n_img * 4, 1024);
  l.gate = off; off += align_up(cexp * n_img * 4, 1024);
  l.total = off;
  return l;
}

mt_mbconv_spec_t spec_of(int i) {
  const BlockSpec& b = kBlocks[i];
  mt_mbconv_spec_t s;
  s.kernel = b.k; s.stride = b.s; s.expand = b.e; s.cin = b.cin; s.cout = b.cout; s.hw_in = b.hw;
  return s;
}

struct EffnetWs { size_t act_a, act_b, block, total; };
EffnetWs effnet_ws_layout(int n_img, int precision) {
  const size_t es = precision == MT_PREC_FP32 ? 4 : 2;
  size_t max_io = (size_t)112 * 112 * 32, max_block = 0;
  for (int i = 0; i < 16; ++i) {
    const mt_mbconv_spec_t b = spec_of(i);
    const size_t ho = (b.hw_in + b.stride - 1) / b.stride;
    max_io = std::max(max_io, ho * ho * b.cout);
    max_block = std::max(max_block, block_ws_layout(b, n_img, precision).total);
  }
  EffnetWs l;
  size_t off = 0;
  l.act_a = off; off += align_up(max_io * n_img * es, 1024);
  l.act_b = off; off += align_up(max_io * n_img * es, 1024);
  l.block = off; off += max_block;
  l.total = off;
  return l;
}

}  // namespace
}  // namespace mt

using namespace mt;

extern "C" int mt_stem_fwd(int precision, const void* x, int x_dtype, const float* w, const float* shift, void* out,
                           int n_img, int h, int w_, void* stream) {
  MT_REQUIRE(x && w && shift && out && n_img > 0 && h > 0 && w_ > 0, "stem: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (precision == MT_PREC_FP32)
    return x_dtype == MT_IN_U8 ? launch_stem_t<float, uint8_t>(x, w, shift, out, n_img, h, w_, st)
                               : launch_stem_t<float, float>(x, w, shift, out, n_img, h, w_, st);
  if (precision == MT_PREC_BF16)
    return x_dtype == MT_IN_U8 ? launch_stem_t<bf16, uint8_t>(x, w, shift, out, n_img, h, w_, st)
                               : launch_stem_t<bf16, float>(x, w, shift, out, n_img, h, w_, st);
  set_error("stem: unknown precision %d", precision);
  return MT_ERR_ARG;
}

extern "C" int mt_dwconv_chunks(int h, int w_, int s) {
  if (h <= 0 || w_ <= 0 || s <= 0) return 0;
  return (((h + s - 1) / s) * ((w_ + s - 1) / s) + kDwPixPerBlock - 1) / kDwPixPerBlock;
}

extern "C" int mt_dwconv_fwd(int precision, const void* in, const float* w, const float* shift, void* out,
                             float* pool_part, int n_img, int h, int w_, int c, int k, int s, void* stream) {
  MT_REQUIRE(in && w && shift && out && pool_part, "dwconv: null pointer");
  MT_REQUIRE(n_img > 0 && h > 0 && w_ > 0 && c > 0 && c % 2 == 0, "dwconv: bad shape n=%d h=%d w=%d c=%d", n_img, h, w_, c);
  MT_REQUIRE(n_img <= 65535, "dwconv: at most 65535 images per call (got %d)", n_img);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (precision == MT_PREC_FP32) return launch_dw_t<float>(in, w, shift, out, pool_part, n_img, h, w_, c, k, s, st);
  if (precision == MT_PREC_BF16) return launch_dw_t<bf16>(in, w, shift, out, pool_part, n_img, h, w_, c, k, s, st);
  set_error("dwconv: unknown precision %d", precision);
  return MT_ERR_ARG;
}

extern "C" int mt_se_gate_fwd(const float* pool_part, int n_chunks, int hw, const float* wr, const float* br,
                              const float* we, const float* be, float* gate, int n_img, int c, int sq, void* stream) {
  MT_REQUIRE(pool_part && wr && br && we && be && gate, "se_gate: null pointer");
  MT_REQUIRE(n_img > 0 && c > 0 && sq > 0 && hw > 0 && n_chunks > 0 && (size_t)(c + sq) * 4 <= 48 * 1024,
             "se_gate: bad shape");
  ProfScope prof(reinterpret_cast<cudaStream_t>(stream), 4.0 * n_img * (double)c * sq,
                 (double)n_img * c * 4 * (n_chunks + 1), "se_gate");
  se_gate_kernel<<<n_img, 256, (size_t)(c + sq) * 4, reinterpret_cast<cudaStream_t>(stream)>>>(
      pool_part, n_chunks, 1.0f / (float)hw, wr, br, we, be, gate, c, sq);
  MT_LAUNCH_CHECK("se_gate_kernel");
  return MT_OK;
}

extern "C" int mt_pointwise_fwd(int precision, const void* a, const void* w, const float* shift, const float* gate,
                                int rows_per_gate, const void* residual, int act, void* out, int m, int n, int k,
                                void* stream) {
  GemmArgs g{};
  g.a = a; g.w = w; g.M = m; g.N = n; g.K = k;
  g.gate = gate; g.rows_per_gate = rows_per_gate;
  g.epi.kind = EPI_STORE; g.epi.M = m; g.epi.N = n;
  g.epi.bias = shift; g.epi.act = act; g.epi.resid = residual; g.epi.out = out; g.epi.ldo = n;
  return launch_gemm(precision, g, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mt_effnet_b0_block_spec(int index, mt_mbconv_spec_t* spec) {
  MT_REQUIRE(spec && index >= 0 && index < 16, "block_spec: index must be in 0..15");
  *spec = spec_of(index);
  return MT_OK;
}

extern "C" size_t mt_mbconv_workspace_bytes(const mt_mbconv_spec_t* spec, int n_img, int precision) {
  if (!spec || n_img <= 0) return 0;
  return block_ws_layout(*spec, n_img, precision).total;
}

// MBConvBlock.forward, eval mode (model.py:89-128):
//   [expand 1x1 + BN + swish] -> depthwise + BN + swish (+ SE squeeze partials) -> SE gate ->
//   project 1x1 (SE gate applied to its input) + BN (+ skip)
extern "C" int mt_mbconv_fwd(int precision, const mt_mbconv_spec_t* spec, const mt_mbconv_t* w, const void* in,
                             void* out, int n_img, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(spec && w && in && out && workspace && n_img > 0, "mbconv: bad argument");
  MT_REQUIRE(precision == MT_PREC_FP32 || precision == MT_PREC_BF16, "mbconv: unknown precision %d", precision);
  const mt_mbconv_spec_t& b = *spec;
  MT_REQUIRE((b.kernel == 3 || b.kernel == 5) && (b.stride == 1 || b.stride == 2) && b.expand >= 1 && b.cin % 8 == 0 &&
                 b.cout % 8 == 0 && b.hw_in > 0,
             "mbconv: unsupported block k=%d s=%d e=%d cin=%d cout=%d", b.kernel, b.stride, b.expand, b.cin, b.cout);
  const BlockWs l = block_ws_layout(b, n_img, precision);
  if (workspace_bytes < l.total) {
    set_error("mbconv: workspace too small (%zu < %zu)", workspace_bytes, l.total);
    return MT_ERR_WORKSPACE;
  }
  MT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "mbconv: workspace must be 1024-byte aligned");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* bexp = ws + l.exp;
  void* bdw = ws + l.dw;
  float* pool = reinterpret_cast<float*>(ws + l.pool);
  float* gate = reinterpret_cast<float*>(ws + l.gate);
  const int cexp = b.cin * b.expand;
  const int ho = (b.hw_in + b.stride - 1) / b.stride;
  const int sq = std::max(1, b.cin / 4);     // max(1, int(cin * 0.25)), model.py:78
  const void* dw_in = in;
  int rc;
  if (b.expand != 1) {
    MT_REQUIRE(w->expand.w, "mbconv: expand weights missing");
    rc = mt_pointwise_fwd(precision, in, w->expand.w, w->expand.shift, nullptr, 0, nullptr, 1, bexp,
                          n_img * b.hw_in * b.hw_in, cexp, b.cin, stream);
    if (rc) return rc;
    dw_in = bexp;
  }
  rc = mt_dwconv_fwd(precision, dw_in, w->dw_w, w->dw_shift, bdw, pool, n_img, b.hw_in, b.hw_in, cexp, b.kernel,
                     b.stride, stream);
  if (rc) return rc;
  rc = mt_se_gate_fwd(pool, dw_chunks(ho), ho * ho, w->se_reduce_w, w->se_reduce_b, w->se_expand_w, w->se_expand_b,
                      gate, n_img, cexp, sq, stream);
  if (rc) return rc;
  const bool skip = b.stride == 1 && b.cin == b.cout;   // model.py:123
  return mt_pointwise_fwd(precision, bdw, w->project.w, w->project.shift, gate, ho * ho, skip ? in : nullptr, 0, out,
                          n_img * ho * ho, b.cout, cexp, stream);
}

extern "C" size_t mt_effnet_b0_workspace_bytes(int n_img, int precision) {
  if (n_img <= 0) return 0;
  return effnet_ws_layout(n_img, precision).total;
}

extern "C" int mt_effnet_b0_fwd(const mt_effnet_b0_weights_t* w, const void* x, int x_dtype, void* feats, int n_img,
                                int precision, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(w && x && feats && workspace, "effnet_b0_fwd: null pointer");
  MT_REQUIRE(n_img > 0 && n_img <= 65535, "effnet_b0_fwd: n_img must be in 1..65535 (got %d)", n_img);
  MT_REQUIRE(precision == MT_PREC_FP32 || precision == MT_PREC_BF16, "effnet_b0_fwd: unknown precision %d", precision);
  const EffnetWs l = effnet_ws_layout(n_img, precision);
  if (workspace_bytes < l.total) {
    set_error("effnet_b0_fwd: workspace too small (%zu < %zu)", workspace_bytes, l.total);
    return MT_ERR_WORKSPACE;
  }
  MT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "effnet_b0_fwd: workspace must be 1024-byte aligned");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* cur = ws + l.act_a;
  void* nxt = ws + l.act_b;
  int rc = mt_stem_fwd(precision, x, x_dtype, w->stem_w, w->stem_shift, cur, n_img, 224, 224, stream);
  if (rc) return rc;
  for (int i = 0; i < 16; ++i) {
    const mt_mbconv_spec_t b = spec_of(i);
    rc = mt_mbconv_fwd(precision, &b, &w->blocks[i], cur, nxt, n_img, ws + l.block, l.total - l.block, stream);
    if (rc) return rc;
    std::swap(cur, nxt);
  }
  // head 1x1 (320 -> 1280) + BN + swish, written straight into the token layout the patch embedding reads
  return mt_pointwise_fwd(precision, cur, w->head.w, w->head.shift, nullptr, 0, nullptr, 1, feats, n_img * 49, 1280,
                          320, stream);
}
