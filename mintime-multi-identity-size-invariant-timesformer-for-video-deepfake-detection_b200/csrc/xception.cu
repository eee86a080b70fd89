// Xception feature extractor, eval mode (reference models/xception.py:93-183 `Xception.features`, the alternative
// 2048-channel extractor of train.py:129-133 / predict.py:365-369, `--extractor_model 1`).
//
// NHWC throughout, element type T (fp32 = exact path, bf16 = tensor-core path).  Every 1x1 convolution -- the pointwise
// half of the 34 separable convolutions, the four strided skip projections -- and, through an im2col staging pass, the two
// dense 3x3 stem convolutions run on the library's GEMM (mt_pointwise_fwd: tcgen05 on the bf16 path) with the following
// BatchNorm folded into the weight rows and the shift.  What this file adds are the memory-bound pieces around them:
//   * im2col of a 3x3 VALID convolution (xception.py:105,109: conv1 stride 2, conv2 stride 1, padding 0)
//   * depthwise 3x3, padding 1, no bias (SeparableConv2d.conv1, xception.py:20) with the preceding ReLU folded into its loads:
//     bf16 on the MBConv stack's TMA-staged FFMA2 kernel (dwconv_simt.cuh MODE 1, via launch_dw_plain_bf16), fp32 on the simple
//     kernel below
//   * MaxPool2d(3, 2, 1) (xception.py:63)
//   * the stride-2 pixel gather in front of a skip projection (xception.py:35)
// ReLUs never run as kernels: each one sits in front of exactly one consumer (a depthwise conv, the im2col of conv2, or
// block 1's skip gather), which applies it while loading.
#include <cfloat>

#include "common.cuh"

namespace mt {
namespace {

template <typename T>
__device__ __forceinline__ float ld_in(const T* p) { return to_f(*p); }
__device__ __forceinline__ float ld_in(const uint8_t* p) { return (float)*p; }

// A [n*ho*wo][kp] = patches of in [n][h][w][c] (3x3, VALID, stride s), column (ky*3+kx)*c + ci; columns >= 9c are zero.
template <typename TI, typename T>
__global__ void __launch_bounds__(256) xc_im2col3x3_kernel(const TI* __restrict__ in, T* __restrict__ out, int n, int h, int w,
                                                           int c, int s, int ho, int wo, int kp, int relu_in) {
  const size_t total = (size_t)n * ho * wo * kp;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const int col = (int)(idx % kp);
    const size_t pix = idx / kp;
    float v = 0.f;
    if (col < 9 * c) {
      const int tap = col / c, ci = col - tap * c;
      const int ky = tap / 3, kx = tap - ky * 3;
      const int ox = (int)(pix % wo);
      const size_t t = pix / wo;
      const int oy = (int)(t % ho);
      const int img = (int)(t / ho);
      v = ld_in(in + (((size_t)img * h + oy * s + ky) * w + ox * s + kx) * c + ci);
      if (relu_in) v = fmaxf(v, 0.f);
    }
    out[idx] = from_f<T>(v);
  }
}

// out[n][h][w][c] = sum_{ky,kx} relu?(in[n][y+ky-1][x+kx-1][c]) * wt[ky*3+kx][c]   (zero padding 1, stride 1)
// A thread owns 8 consecutive channels of one pixel (one 16-byte load per tap on the bf16 path); lanes run over channel
// groups, so a warp reads whole contiguous channel rows; the 9x re-read of the input is served by L1/L2.  (A sliding-window
// variant -- 8 output pixels per thread, 3 loads per output -- measured 2x slower: 176 registers, one block per SM.)
// c is a multiple of 8 for every Xception layer.
template <typename T>
__global__ void __launch_bounds__(256) xc_dw3x3_kernel(const T* __restrict__ in, const float* __restrict__ wt,
                                                       T* __restrict__ out, int n, int h, int w, int c, int relu_in) {
  const int c8 = c >> 3;
  const size_t total = (size_t)n * h * w * c8;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const int cg = (int)(idx % c8) * 8;
    const size_t pix = idx / c8;
    const int x = (int)(pix % w);
    const size_t t = pix / w;
    const int y = (int)(t % h);
    const int img = (int)(t / h);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = y + ky - 1;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = x + kx - 1;
        if (ix < 0 || ix >= w) continue;
        float v[8], wv[8];
        load8(in + (((size_t)img * h + iy) * w + ix) * c + cg, v);
        load8(wt + (size_t)(ky * 3 + kx) * c + cg, wv);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(relu_in ? fmaxf(v[i], 0.f) : v[i], wv[i], acc[i]);
      }
    }
    store8(out + pix * c + cg, acc);
  }
}

// MaxPool2d(kernel 3, stride 2, padding 1): ho = (h - 1) / 2 + 1; padded positions never win (-inf).  8 channels per thread.
template <typename T>
__global__ void __launch_bounds__(256) xc_maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int h, int w,
                                                         int c, int ho, int wo) {
  const int c8 = c >> 3;
  const size_t total = (size_t)n * ho * wo * c8;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const int cg = (int)(idx % c8) * 8;
    const size_t pix = idx / c8;
    const int ox = (int)(pix % wo);
    const size_t t = pix / wo;
    const int oy = (int)(t % ho);
    const int img = (int)(t / ho);
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -FLT_MAX;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 + ky - 1;
      if (iy < 0 || iy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 + kx - 1;
        if (ix < 0 || ix >= w) continue;
        float v[8];
        load8(in + (((size_t)img * h + iy) * w + ix) * c + cg, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
      }
    }
    store8(out + pix * c + cg, m);
  }
}

// out[n][ho][wo][c] = relu?(in[n][2*oy][2*ox][c]): the pixels a 1x1 stride-2 convolution reads (xception.py:35)
template <typename T>
__global__ void __launch_bounds__(256) xc_gather2_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int h, int w,
                                                         int c, int ho, int wo, int relu_in) {
  const int c8 = c >> 3;
  const size_t total = (size_t)n * ho * wo * c8;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const int cg = (int)(idx % c8) * 8;
    const size_t pix = idx / c8;
    const int ox = (int)(pix % wo);
    const size_t t = pix / wo;
    const int oy = (int)(t % ho);
    const int img = (int)(t / ho);
    float v[8];
    load8(in + (((size_t)img * h + 2 * oy) * w + 2 * ox) * c + cg, v);
    if (relu_in) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    store8(out + pix * c + cg, v);
  }
}

// im2col of a 3x3 VALID convolution whose input has a multiple of 8 channels (conv2): one thread per (pixel, tap, 8 channels)
template <typename T>
__global__ void __launch_bounds__(256) xc_im2col3x3_c8_kernel(const T* __restrict__ in, T* __restrict__ out, int n, int h, int w,
                                                              int c, int s, int ho, int wo, int relu_in) {
  const int c8 = c >> 3, per_pix = 9 * c8;
  const size_t total = (size_t)n * ho * wo * per_pix;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const int q = (int)(idx % per_pix);
    const size_t pix = idx / per_pix;
    const int tap = q / c8, cg = (q - tap * c8) * 8;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int ox = (int)(pix % wo);
    const size_t t = pix / wo;
    const int oy = (int)(t % ho);
    const int img = (int)(t / ho);
    float v[8];
    load8(in + (((size_t)img * h + oy * s + ky) * w + ox * s + kx) * c + cg, v);
    if (relu_in) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    store8(out + pix * (size_t)(9 * c) + (size_t)tap * c + cg, v);
  }
}

inline unsigned grid_for(size_t total);

// depthwise 3x3 of one separable convolution: bf16 on the TMA-staged FFMA2 kernel of the MBConv stack (effnet.cu, MODE 1),
// fp32 (exact path) on the simple kernel above
template <typename T>
int xc_dw(const T* in, const float* w, T* out, int n, int h, int c, int relu_in, cudaStream_t st);

inline unsigned grid_for(size_t total) {
  const size_t blocks = (total + 255) / 256;
  return (unsigned)std::min<size_t>(blocks, (size_t)current_sms() * 32);
}

template <>
int xc_dw<bf16>(const bf16* in, const float* w, bf16* out, int n, int h, int c, int relu_in, cudaStream_t st) {
  return launch_dw_plain_bf16(in, w, out, n, h, c, relu_in, st);
}
template <>
int xc_dw<float>(const float* in, const float* w, float* out, int n, int h, int c, int relu_in, cudaStream_t st) {
  const size_t total = (size_t)n * h * h * (c / 8);
  ProfScope ps(st, 18.0 * (double)n * h * h * c, 2.0 * (double)n * h * h * c * 4.0, "xc_dw3x3 C%d H%d", c, h);
  xc_dw3x3_kernel<float><<<grid_for(total), 256, 0, st>>>(in, w, out, n, h, h, c, relu_in);
  MT_LAUNCH_CHECK("xc_dw3x3_kernel");
  return MT_OK;
}

// geometry of the network at a given input size (224 in MINTIME): xception.py:105-131
struct XcGeom {
  int h0, h1, h2;          // input, after conv1 (3x3 s2 p0), after conv2 (3x3 s1 p0)
  int hb[13];              // hb[i] = spatial size of block i's output (hb[0] = h2)
};
inline XcGeom xc_geom(int h) {
  XcGeom g{};
  g.h0 = h;
  g.h1 = (h - 3) / 2 + 1;
  g.h2 = g.h1 - 2;
  g.hb[0] = g.h2;
  for (int b = 1; b <= 12; ++b) {
    const bool strided = b <= 3 || b == 12;
    g.hb[b] = strided ? (g.hb[b - 1] - 1) / 2 + 1 : g.hb[b - 1];
  }
  return g;
}

// blocks 1..12: (cin, cout, reps, stride, start_with_relu, grow_first)   xception.py:113-129
struct XcBlock { int cin, cout, reps, stride, relu0, grow_first; };
inline XcBlock xc_block(int b) {
  switch (b) {
    case 1: return {64, 128, 2, 2, 0, 1};
    case 2: return {128, 256, 2, 2, 1, 1};
    case 3: return {256, 728, 2, 2, 1, 1};
    case 12: return {728, 1024, 2, 2, 1, 0};
    default: return {728, 728, 3, 1, 1, 1};
  }
}

struct XcWs { size_t a, b, c, col, total; };
inline XcWs xc_ws(int n_img, int precision, int h) {
  const XcGeom g = xc_geom(h);
  const size_t es = precision == MT_PREC_FP32 ? 4 : 2;
  auto al = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
  // largest activation: conv2's output (h2 x h2 x 64) or block 1's 128-channel tensors at h2; im2col of conv2: h2^2 x 288
  size_t act = (size_t)g.h1 * g.h1 * 32;
  act = std::max(act, (size_t)g.h2 * g.h2 * 128);
  act = std::max(act, (size_t)g.hb[12] * g.hb[12] * 2048);
  const size_t col = std::max((size_t)g.h1 * g.h1 * 32, (size_t)g.h2 * g.h2 * 288);
  XcWs w{};
  size_t off = 0;
  w.a = off; off += al(act * n_img * es);
  w.b = off; off += al(act * n_img * es);
  w.c = off; off += al(act * n_img * es);
  w.col = off; off += al(col * n_img * es);
  w.total = off;
  return w;
}

template <typename T>
int xc_forward(const mt_xception_weights_t* w, const void* x, int x_dtype, void* feats, int n, int h, int precision,
               uint8_t* ws, cudaStream_t st) {
  const XcGeom g = xc_geom(h);
  const XcWs l = xc_ws(n, precision, h);
  T* A = reinterpret_cast<T*>(ws + l.a);
  T* B = reinterpret_cast<T*>(ws + l.b);
  T* C = reinterpret_cast<T*>(ws + l.c);
  T* col = reinterpret_cast<T*>(ws + l.col);
  void* stream = st;
  int rc;
  // ---- conv1 3x3 s2 p0 (3 -> 32) + bn1 [+ relu, applied by conv2's im2col]            xception.py:105-107,148-150
  {
    const size_t total = (size_t)n * g.h1 * g.h1 * 32;
    if (x_dtype == MT_IN_U8)
      xc_im2col3x3_kernel<uint8_t, T><<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(x), col, n, h, h, 3, 2,
                                                                       g.h1, g.h1, 32, 0);
    else
      xc_im2col3x3_kernel<float, T><<<grid_for(total), 256, 0, st>>>(reinterpret_cast<const float*>(x), col, n, h, h, 3, 2,
                                                                     g.h1, g.h1, 32, 0);
    MT_LAUNCH_CHECK("xc_im2col3x3_kernel(conv1)");
    rc = mt_pointwise_fwd(precision, col, w->conv1.w, w->conv1.shift, nullptr, 0, nullptr, 0, A, n * g.h1 * g.h1, 32, 32, stream);
    if (rc) return rc;
  }
  // ---- conv2 3x3 s1 p0 (32 -> 64) + bn2 [+ relu, applied by block 1's consumers]       xception.py:109-111,152-154
  {
    const size_t total = (size_t)n * g.h2 * g.h2 * 36;
    {
      ProfScope ps(st, 0.0, ((double)n * g.h1 * g.h1 * 32 + (double)n * g.h2 * g.h2 * 288) * (double)sizeof(T), "xc_im2col conv2");
      xc_im2col3x3_c8_kernel<T><<<grid_for(total), 256, 0, st>>>(A, col, n, g.h1, g.h1, 32, 1, g.h2, g.h2, 1);
      MT_LAUNCH_CHECK("xc_im2col3x3_kernel(conv2)");
    }
    rc = mt_pointwise_fwd(precision, col, w->conv2.w, w->conv2.shift, nullptr, 0, nullptr, 0, B, n * g.h2 * g.h2, 64, 288, stream);
    if (rc) return rc;
  }
  // cur = block input (for block 1: bn2's output BEFORE the relu of xception.py:154, which both of block 1's consumers apply)
  T* cur = B;
  T* t0 = A;
  T* t1 = C;
  int unit = 0;
  for (int b = 1; b <= 12; ++b) {
    const XcBlock bk = xc_block(b);
    const int hin = g.hb[b - 1], hout = g.hb[b];
    const int rows_in = n * hin * hin;
    // the separable convolutions of `rep` (xception.py:43-58): [relu] dw3x3 -> 1x1 + BN
    T* src = cur;
    int ch = bk.cin;
    for (int r = 0; r < bk.reps; ++r, ++unit) {
      int cout;
      if (bk.grow_first) cout = bk.cout;
      else cout = (r == bk.reps - 1) ? bk.cout : bk.cin;
      const int relu_in = (b == 1) ? 1 : ((r == 0) ? bk.relu0 : 1);   // block 1: its input is relu(bn2(.)) (xception.py:154)
      const mt_xc_sep_t& u = w->sep[unit];
      T* dwo = (src == t0) ? t1 : t0;
      rc = xc_dw<T>(src, u.dw_w, dwo, n, hin, ch, relu_in, st);
      if (rc) return rc;
      T* pwo = (dwo == t0) ? t1 : t0;
      // identity-skip blocks add their input in the last unit's epilogue (xception.py:73-75)
      const bool last = r == bk.reps - 1;
      const void* resid = (last && bk.stride == 1 && bk.cin == bk.cout) ? cur : nullptr;
      rc = mt_pointwise_fwd(precision, dwo, u.pw.w, u.pw.shift, nullptr, 0, resid, 0, pwo, rows_in, cout, ch, stream);
      if (rc) return rc;
      src = pwo;
      ch = cout;
    }
    T* outp;
    if (bk.stride != 1) {
      // MaxPool2d(3, 2, 1) on the main path, 1x1 stride-2 projection + BN of the block INPUT on the skip path, summed in the
      // projection's epilogue (xception.py:62-63, 68-76)
      T* pooled = (src == t0) ? t1 : t0;
      const size_t tp = (size_t)n * hout * hout * (ch / 8);
      {
        ProfScope ps(st, 0.0, ((double)rows_in + (double)n * hout * hout) * ch * (double)sizeof(T), "xc_maxpool C%d H%d", ch, hin);
        xc_maxpool_kernel<T><<<grid_for(tp), 256, 0, st>>>(src, pooled, n, hin, hin, ch, hout, hout);
        MT_LAUNCH_CHECK("xc_maxpool_kernel");
      }
      const size_t tg = (size_t)n * hout * hout * (bk.cin / 8);
      xc_gather2_kernel<T><<<grid_for(tg), 256, 0, st>>>(cur, col, n, hin, hin, bk.cin, hout, hout, b == 1 ? 1 : 0);
      MT_LAUNCH_CHECK("xc_gather2_kernel");
      const int si = b <= 3 ? b - 1 : 3;
      T* dst = (pooled == t0) ? t1 : t0;
      rc = mt_pointwise_fwd(precision, col, w->skip[si].w, w->skip[si].shift, nullptr, 0, pooled, 0, dst, n * hout * hout, bk.cout,
                            bk.cin, stream);
      if (rc) return rc;
      outp = dst;
    } else {
      outp = src;
    }
    // rotate buffers: the block output becomes `cur`; the other two are scratch
    T* bufs[3] = {A, B, C};
    int k = 0;
    T* free2[2];
    for (int i = 0; i < 3; ++i)
      if (bufs[i] != outp) free2[k++] = bufs[i];
    cur = outp; t0 = free2[0]; t1 = free2[1];
  }
  // ---- conv3 (sep 1024 -> 1536) + bn3 [+ relu by conv4's depthwise], conv4 (sep 1536 -> 2048) + bn4   xception.py:131-137,176-183
  {
    const int hh = g.hb[12], rows = n * hh * hh;
    const mt_xc_sep_t& u3 = w->sep[unit];
    const mt_xc_sep_t& u4 = w->sep[unit + 1];
    rc = xc_dw<T>(cur, u3.dw_w, t0, n, hh, 1024, 0, st);
    if (rc) return rc;
    rc = mt_pointwise_fwd(precision, t0, u3.pw.w, u3.pw.shift, nullptr, 0, nullptr, 0, t1, rows, 1536, 1024, stream);
    if (rc) return rc;
    rc = xc_dw<T>(t1, u4.dw_w, t0, n, hh, 1536, 1, st);
    if (rc) return rc;
    rc = mt_pointwise_fwd(precision, t0, u4.pw.w, u4.pw.shift, nullptr, 0, nullptr, 0, feats, rows, 2048, 1536, stream);
    if (rc) return rc;
  }
  return MT_OK;
}

}  // namespace
}  // namespace mt

using namespace mt;

extern "C" int mt_xception_out_hw(int in_hw) {
  if (in_hw < 35) return 0;
  return xc_geom(in_hw).hb[12];
}

extern "C" size_t mt_xception_workspace_bytes(int n_img, int in_hw, int precision) {
  if (n_img <= 0 || in_hw < 35) return 0;
  return xc_ws(n_img, precision, in_hw).total;
}

extern "C" int mt_xception_fwd(const mt_xception_weights_t* w, const void* x, int x_dtype, void* feats, int n_img, int in_hw,
                               int precision, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(w && x && feats && n_img > 0 && in_hw >= 35, "xception: bad argument");
  MT_REQUIRE(x_dtype == MT_IN_F32 || x_dtype == MT_IN_U8, "xception: x_dtype must be MT_IN_F32 or MT_IN_U8");
  MT_REQUIRE(precision == MT_PREC_FP32 || precision == MT_PREC_BF16, "xception: unknown precision %d", precision);
  if (!workspace || workspace_bytes < mt_xception_workspace_bytes(n_img, in_hw, precision) ||
      (reinterpret_cast<uintptr_t>(workspace) & 1023)) {
    set_error("xception: workspace missing, too small or not 1024-byte aligned");
    return MT_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  return precision == MT_PREC_FP32 ? xc_forward<float>(w, x, x_dtype, feats, n_img, in_hw, precision, ws, st)
                                   : xc_forward<bf16>(w, x, x_dtype, feats, n_img, in_hw, precision, ws, st);
}
