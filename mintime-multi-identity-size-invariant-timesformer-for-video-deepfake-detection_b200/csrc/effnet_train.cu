// EfficientNet-B0 in TRAIN mode (reference train.py:153-170: the extractor is in .train() unless --freeze_backbone):
// the building blocks of MBConvBlock.forward / backward that the eval path does not have, fp32 NHWC, through the C ABI:
//   * BatchNorm with BATCH statistics + running-stat update (utils.py:520-521: momentum 0.01, eps 1e-3), its backward
//   * raw convolutions (no folded BN): stem 3x3 s2 and depthwise kxk with TF-"SAME" padding (utils.py:248-276),
//     depthwise data / weight gradients, stem weight gradient
//   * swish backward sigma(x)(1 + x(1 - sigma(x))) (utils.py:71-80) fused into the BatchNorm backward
//   * squeeze-excite: pooled mean, the two small FC layers forward / backward, gate multiply backward (model.py:110-115)
//   * drop-connect scale + skip add (utils.py:129-154, model.py:123-127)
// The 1x1 convolutions are GEMMs: mt_pointwise_fwd (forward and data gradient on W^T), mt_grad_prep + mt_linear_wgrad
// (weight gradient).  Orchestrated by mintime_b200/efficientnet_train.py (one torch.autograd.Function).
// These kernels favour exactness and determinism (fixed-order reductions, fp64 accumulation of the statistics) over
// speed: the training-mode extractor is new in round 2; its bf16 tensor-core variants are the next step.
#include <float.h>

#include <algorithm>

#include "common.cuh"

namespace mt {
namespace {

__host__ __device__ inline int same_pad_lo_t(int in, int k, int s) {
  const int out = (in + s - 1) / s;
  int total = (out - 1) * s + k - in;
  if (total < 0) total = 0;
  return total / 2;
}

__device__ __forceinline__ float swish_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float swish_grad(float x) {           // utils.py:76-80
  const float s = 1.0f / (1.0f + expf(-x));
  return s * (1.0f + x * (1.0f - s));
}

constexpr int kRedRows = 8;        // row lanes per block in the per-channel reductions (block = 32 channels x 8 lanes)

// ---- per-channel sums over the rows of x [rows][C]: partial[blockIdx.y][c][{0,1}] = (sum v0, sum v1) in fp64.
// MODE 0: v0 = x, v1 = x^2 (BatchNorm statistics).
// MODE 1: v0 = dz, v1 = dz * xhat with dz = dy * act'(z), z = gamma*xhat + beta (BatchNorm backward).
template <int MODE>
__global__ void __launch_bounds__(256) chan_sums_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                        const float* __restrict__ mean, const float* __restrict__ var,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        int act, float eps, double* __restrict__ partial, int rows, int C,
                                                        int rows_per_block) {
  __shared__ double s0[kRedRows][33], s1[kRedRows][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  double a0 = 0.0, a1 = 0.0;
  if (c < C) {
    float mu = 0.f, rstd = 0.f, g = 0.f, b = 0.f;
    if (MODE == 1) { mu = mean[c]; rstd = 1.0f / sqrtf(var[c] + eps); g = gamma[c]; b = beta[c]; }
    for (int r = r0 + rl; r < r1; r += kRedRows) {
      const float v = x[(size_t)r * C + c];
      if (MODE == 0) {
        a0 += (double)v;
        a1 += (double)v * (double)v;
      } else {
        const float xh = (v - mu) * rstd;
        float dz = dy[(size_t)r * C + c];
        if (act == 1) dz *= swish_grad(fmaf(g, xh, b));
        a0 += (double)dz;
        a1 += (double)dz * (double)xh;
      }
    }
  }
  s0[rl][cl] = a0; s1[rl][cl] = a1;
  __syncthreads();
  if (rl == 0 && c < C) {
    for (int i = 1; i < kRedRows; ++i) { a0 += s0[i][cl]; a1 += s1[i][cl]; }     // fixed order
    partial[((size_t)blockIdx.y * C + c) * 2] = a0;
    partial[((size_t)blockIdx.y * C + c) * 2 + 1] = a1;
  }
}

// BatchNorm statistics from the partial sums (fixed order over the row blocks) + running-stat update.
__global__ void bn_stats_finish_kernel(const double* __restrict__ partial, int nblk, int rows, int C, float* __restrict__ mean,
                                       float* __restrict__ var, float* running_mean, float* running_var, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = 0; b < nblk; ++b) { s += partial[((size_t)b * C + c) * 2]; q += partial[((size_t)b * C + c) * 2 + 1]; }
  const double m = s / rows;
  double v = q / rows - m * m;                      // biased variance (what F.batch_norm normalises with)
  if (v < 0.0) v = 0.0;
  mean[c] = (float)m;
  var[c] = (float)v;
  if (running_mean) {                               // nn.BatchNorm2d: running stats move towards mean / UNBIASED variance
    const double vu = rows > 1 ? v * rows / (rows - 1) : 0.0;
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)vu;
  }
}

// y = act(gamma * (x - mean) * rstd + beta)
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                         const float* __restrict__ var, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int act, float eps,
                                                         float* __restrict__ out, size_t total, int C) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float z = fmaf(gamma[c], (x[i] - mean[c]) * (1.0f / sqrtf(var[c] + eps)), beta[c]);
    out[i] = act == 1 ? swish_f(z) : z;
  }
}

// dgamma = sum dz*xhat, dbeta = sum dz (from the partial sums, fixed order); sums[c] = (sum dz, sum dz*xhat) as fp32
__global__ void bn_bwd_finish_kernel(const double* __restrict__ partial, int nblk, int C, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, float* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int b = 0; b < nblk; ++b) { s += partial[((size_t)b * C + c) * 2]; q += partial[((size_t)b * C + c) * 2 + 1]; }
  dbeta[c] = (float)s;
  dgamma[c] = (float)q;
  sums[2 * c] = (float)s;
  sums[2 * c + 1] = (float)q;
}

// dx = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat))
__global__ void __launch_bounds__(256) bn_act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                         const float* __restrict__ mean, const float* __restrict__ var,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ sums, int act, float eps,
                                                         float* __restrict__ dx, size_t total, int C, float inv_rows) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float rstd = 1.0f / sqrtf(var[c] + eps), g = gamma[c];
    const float xh = (x[i] - mean[c]) * rstd;
    float dz = dy[i];
    if (act == 1) dz *= swish_grad(fmaf(g, xh, beta[c]));
    dx[i] = g * rstd * (dz - sums[2 * c] * inv_rows - xh * sums[2 * c + 1] * inv_rows);
  }
}

// ---- stem: raw 3x3 stride-2 convolution 3 -> 32, TF-SAME padding; w tap-major [(ky,kx,ci)][32]
__global__ void __launch_bounds__(256) stem_raw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           float* __restrict__ out, int n, int H, int W, int Ho, int Wo,
                                                           int pad) {
  const size_t total = (size_t)n * Ho * Wo * 32;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i & 31);
    size_t p = i >> 5;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int img = (int)(p / Ho);
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 + ky - pad;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 + kx - pad;
        if (ix < 0 || ix >= W) continue;
        const float* px = x + (((size_t)img * H + iy) * W + ix) * 3;
        const float* wr = w + ((ky * 3 + kx) * 3) * 32 + co;
        acc = fmaf(px[0], wr[0], fmaf(px[1], wr[32], fmaf(px[2], wr[64], acc)));
      }
    }
    out[i] = acc;
  }
}

// dw[(ky,kx,ci)][co] partial per block of output rows: thread = (tap 0..26, co); fp64 accumulation
__global__ void __launch_bounds__(864) stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         double* __restrict__ partial, int n, int H, int W, int Ho, int Wo,
                                                         int pad, int rows_per_block) {
  const int co = threadIdx.x & 31, tap = threadIdx.x >> 5;      // 27 taps x 32 outputs
  const int ci = tap % 3, kx = (tap / 3) % 3, ky = tap / 9;
  const long long rows = (long long)n * Ho * Wo;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  double acc = 0.0;
  // (ox, oy, img) walk with the row index: one decode per block instead of three 64-bit divisions per row
  int ox = (int)(r0 % Wo), oy = (int)((r0 / Wo) % Ho), img = (int)(r0 / ((long long)Wo * Ho));
  for (long long r = r0; r < r1; ++r) {
    const int iy = oy * 2 + ky - pad, ix = ox * 2 + kx - pad;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      acc += (double)dy[(size_t)r * 32 + co] * (double)x[(((size_t)img * H + iy) * W + ix) * 3 + ci];
    if (++ox == Wo) {
      ox = 0;
      if (++oy == Ho) { oy = 0; ++img; }
    }
  }
  partial[(size_t)blockIdx.x * 864 + threadIdx.x] = acc;
}

// out[j] = sum_b partial[b][j] (fixed order), fp64 -> fp32
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nblk, int width, float* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= width) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * width + j];
  out[j] = (float)s;
}

// ---- depthwise kxk stride s, raw; w tap-major [k*k][C].  K and S are compile-time: with run-time values the tap loops carried
// an integer division / modulo per tap (dgrad: 1 ms per layer at 128 images).
template <int K, int S>
__global__ void __launch_bounds__(256) dw_raw_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                         float* __restrict__ out, int n, int H, int Ho, int C, int pad) {
  const size_t total = (size_t)n * Ho * Ho * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t p = i / C;
    const int ox = (int)(p % Ho); p /= Ho;
    const int oy = (int)(p % Ho);
    const int img = (int)(p / Ho);
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int iy = oy * S + ky - pad;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int ix = ox * S + kx - pad;
        if (ix < 0 || ix >= H) continue;
        acc = fmaf(in[(((size_t)img * H + iy) * H + ix) * C + c], w[(ky * K + kx) * C + c], acc);
      }
    }
    out[i] = acc;
  }
}

// dx[n][iy][ix][c] = sum over taps of dy[n][oy][ox][c] * w[ky][kx][c] with iy = oy*S + ky - pad
template <int K, int S>
__global__ void __launch_bounds__(256) dw_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                       float* __restrict__ dx, int n, int H, int Ho, int C, int pad) {
  const size_t total = (size_t)n * H * H * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    size_t p = i / C;
    const int ix = (int)(p % H); p /= H;
    const int iy = (int)(p % H);
    const int img = (int)(p / H);
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int ty = iy + pad - ky;
      if (ty < 0 || (S == 2 && (ty & 1))) continue;
      const int oy = S == 2 ? ty >> 1 : ty;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int tx = ix + pad - kx;
        if (tx < 0 || (S == 2 && (tx & 1))) continue;
        const int ox = S == 2 ? tx >> 1 : tx;
        if (ox >= Ho) continue;
        acc = fmaf(dy[(((size_t)img * Ho + oy) * Ho + ox) * C + c], w[(ky * K + kx) * C + c], acc);
      }
    }
    dx[i] = acc;
  }
}

// dw[tap][c] partial per block of output positions; block = 32 channels x 8 position lanes, K*K taps in registers
template <int K>
__global__ void __launch_bounds__(256) dw_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dy,
                                                       double* __restrict__ partial, int n, int H, int Ho, int C, int S,
                                                       int pad, int pos_per_block) {
  __shared__ double red[kRedRows][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const long long npos = (long long)n * Ho * Ho;
  const long long p0 = (long long)blockIdx.y * pos_per_block, p1 = min(npos, p0 + pos_per_block);
  // fp32 inside a thread's share of the chunk (a few hundred positions), fp64 across threads and chunks: the double-precision
  // converts and FMAs per tap made this kernel 8 ms of the unfrozen step
  float acc[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = 0.f;
  if (c < C) {
    for (long long p = p0 + rl; p < p1; p += kRedRows) {
      const unsigned pp = (unsigned)p;                      // (npos < 2^31: 32-bit divisions)
      const int ox = (int)(pp % (unsigned)Ho), oy = (int)((pp / (unsigned)Ho) % (unsigned)Ho), img = (int)(pp / (unsigned)(Ho * Ho));
      const float g = dy[(size_t)p * C + c];
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int iy = oy * S + ky - pad;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int ix = ox * S + kx - pad;
          if (ix < 0 || ix >= H) continue;
          acc[ky * K + kx] = fmaf(g, in[(((size_t)img * H + iy) * H + ix) * C + c], acc[ky * K + kx]);
        }
      }
    }
  }
  for (int t = 0; t < K * K; ++t) {
    __syncthreads();
    red[rl][cl] = (double)acc[t];
    __syncthreads();
    if (rl == 0 && c < C) {
      double s = red[0][cl];
      for (int i = 1; i < kRedRows; ++i) s += red[i][cl];
      partial[((size_t)blockIdx.y * K * K + t) * C + c] = s;
    }
  }
}

// ---- squeeze-excite
// out[g][c] = mean over rows of x[g][rows][c]  (adaptive_avg_pool2d(x, 1), model.py:110); block = (32 channels, 8 lanes)
__global__ void __launch_bounds__(256) group_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int rows, int C) {
  __shared__ double red[kRedRows][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl, g = blockIdx.y;
  double a = 0.0;
  if (c < C)
    for (int r = rl; r < rows; r += kRedRows) a += (double)x[((size_t)g * rows + r) * C + c];
  red[rl][cl] = a;
  __syncthreads();
  if (rl == 0 && c < C) {
    for (int i = 1; i < kRedRows; ++i) a += red[i][cl];
    out[(size_t)g * C + c] = (float)(a / rows);
  }
}

// s_pre = Wr m + br; gate = sigmoid(We swish(s_pre) + be); one block per image.  wr [SQ][C], we [C][SQ] (reference layouts)
__global__ void __launch_bounds__(256) se_fc_fwd_kernel(const float* __restrict__ m, const float* __restrict__ wr,
                                                        const float* __restrict__ br, const float* __restrict__ we,
                                                        const float* __restrict__ be, float* __restrict__ gate,
                                                        float* __restrict__ s_pre, int C, int SQ) {
  extern __shared__ float sm[];          // [SQ] swish(s_pre)
  const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* mi = m + (size_t)img * C;
  for (int j = warp; j < SQ; j += 8) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(wr[(size_t)j * C + c], mi[c], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      a += br[j];
      s_pre[(size_t)img * SQ + j] = a;
      sm[j] = swish_f(a);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = be[c];
    for (int j = 0; j < SQ; ++j) a = fmaf(we[(size_t)c * SQ + j], sm[j], a);
    gate[(size_t)img * C + c] = 1.0f / (1.0f + expf(-a));
  }
}

// per image: de = dgate * g (1 - g); ds = We^T de; dspre = ds * swish'(s_pre); dm = Wr^T dspre.  Writes de [n][C],
// dspre [n][SQ], s = swish(s_pre) [n][SQ] (the weight gradients are outer-product sums over the images) and dm [n][C].
__global__ void __launch_bounds__(256) se_fc_bwd_kernel(const float* __restrict__ dgate, const float* __restrict__ gate,
                                                        const float* __restrict__ s_pre, const float* __restrict__ wr,
                                                        const float* __restrict__ we, float* __restrict__ de,
                                                        float* __restrict__ dspre, float* __restrict__ s_out,
                                                        float* __restrict__ dm, int C, int SQ) {
  extern __shared__ float sm[];          // [C] de, then [SQ] dspre
  float* de_s = sm;
  float* ds_s = sm + C;
  const int img = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float g = gate[(size_t)img * C + c];
    const float v = dgate[(size_t)img * C + c] * g * (1.0f - g);
    de_s[c] = v;
    de[(size_t)img * C + c] = v;
  }
  __syncthreads();
  for (int j = warp; j < SQ; j += 8) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(we[(size_t)c * SQ + j], de_s[c], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) {
      const float sp = s_pre[(size_t)img * SQ + j];
      const float d = a * swish_grad(sp);
      ds_s[j] = d;
      dspre[(size_t)img * SQ + j] = d;
      s_out[(size_t)img * SQ + j] = swish_f(sp);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.f;
    for (int j = 0; j < SQ; ++j) a = fmaf(wr[(size_t)j * C + c], ds_s[j], a);
    dm[(size_t)img * C + c] = a;
  }
}

// out[p][q] = sum_n a[n][p] * b[n][q]   (weight gradients of the SE layers: a few hundred images, fixed order)
__global__ void __launch_bounds__(256) outer_sum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        float* __restrict__ out, int n, int P, int Q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * Q) return;
  const int p = i / Q, q = i - p * Q;
  double s = 0.0;
  for (int k = 0; k < n; ++k) s += (double)a[(size_t)k * P + p] * (double)b[(size_t)k * Q + q];
  out[i] = (float)s;
}

// xg = x * gate[image]  (model.py:115) -- the A operand of the project conv's weight gradient
__global__ void __launch_bounds__(256) gate_mul_kernel(const float* __restrict__ x, const float* __restrict__ gate,
                                                       float* __restrict__ out, size_t total, int per_img, int C) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    out[i] = x[i] * gate[(i / per_img) * C + (i % C)];
}

// dgate[g][c] = sum over rows of dxg * x; block = (32 channels, 8 lanes) per image
__global__ void __launch_bounds__(256) gate_bwd_reduce_kernel(const float* __restrict__ dxg, const float* __restrict__ x,
                                                              float* __restrict__ dgate, int rows, int C) {
  __shared__ double red[kRedRows][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl, g = blockIdx.y;
  double a = 0.0;
  if (c < C)
    for (int r = rl; r < rows; r += kRedRows) {
      const size_t i = ((size_t)g * rows + r) * C + c;
      a += (double)dxg[i] * (double)x[i];
    }
  red[rl][cl] = a;
  __syncthreads();
  if (rl == 0 && c < C) {
    for (int i = 1; i < kRedRows; ++i) a += red[i][cl];
    dgate[(size_t)g * C + c] = (float)a;
  }
}

// dx = dxg * gate[image] + dmean[image] / rows   (gate multiply + average-pool backward)
__global__ void __launch_bounds__(256) gate_bwd_apply_kernel(const float* __restrict__ dxg, const float* __restrict__ gate,
                                                             const float* __restrict__ dmean, float* __restrict__ dx,
                                                             size_t total, int per_img, int C, float inv_rows) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t gc = (i / per_img) * C + (i % C);
    dx[i] = fmaf(dxg[i], gate[gc], dmean[gc] * inv_rows);
  }
}

// out = x * scale[image] (+ skip): drop-connect + residual add forward; with skip == null also the backward of the scale
__global__ void __launch_bounds__(256) scale_add_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                        const float* __restrict__ skip, float* __restrict__ out,
                                                        size_t total, int per_img) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float v = x[i];
    if (scale) v *= scale[i / per_img];
    if (skip) v += skip[i];
    out[i] = v;
  }
}

// ---- weight gradient of a 1x1 convolution: dw[co][ci] = sum_rows dy[r][co] * a[r][ci].  The extractor's channel counts
// (16 ... 1152) are not multiples of the 64-wide tiles mt_grad_prep / mt_linear_wgrad work on, so this is a plain
// shared-memory tile kernel: block = 32 x 32 outputs over one chunk of rows, fp32 inside a chunk, chunks summed in a
// fixed order in fp64 (reduce_partials_kernel).
__global__ void __launch_bounds__(256) conv1x1_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ a,
                                                            double* __restrict__ partial, long long rows, int CO, int CI,
                                                            long long rows_per_chunk) {
  __shared__ float sdy[32][33], sa[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // ty 0..7: output rows ty, ty+8, ty+16, ty+24
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  const long long r0 = (long long)blockIdx.z * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long r = r0; r < r1; r += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long rr = r + ty + i * 8;
      const bool ok = rr < r1;
      sdy[ty + i * 8][tx] = (ok && co0 + tx < CO) ? dy[(size_t)rr * CO + co0 + tx] : 0.f;
      sa[ty + i * 8][tx] = (ok && ci0 + tx < CI) ? a[(size_t)rr * CI + ci0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float av = sa[k][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(sdy[k][ty + i * 8], av, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty + i * 8, ci = ci0 + tx;
    if (co < CO && ci < CI) partial[(size_t)blockIdx.z * CO * CI + (size_t)co * CI + ci] = (double)acc[i];
  }
}

int wgrad_chunks(long long rows, int co, int ci) {
  const int tiles = ((co + 31) / 32) * ((ci + 31) / 32);
  long long chunks = std::max<long long>(1, std::min<long long>(64, (long long)current_sms() * 4 / tiles));
  chunks = std::min<long long>(chunks, (rows + 255) / 256);
  return (int)std::max<long long>(chunks, 1);
}

int ew_grid(size_t total) { return (int)std::min<size_t>((total + 255) / 256, (size_t)current_sms() * 16); }

struct RedGeom { int cblk, nblk, per_block; };
RedGeom red_geom(long long rows, int C) {
  RedGeom g;
  g.cblk = (C + 31) / 32;
  const int want = std::max(1, current_sms() * 4 / g.cblk);
  g.per_block = (int)std::max<long long>(kRedRows * 8, (rows + want - 1) / want);
  g.nblk = (int)((rows + g.per_block - 1) / g.per_block);
  return g;
}

}  // namespace
}  // namespace mt

using namespace mt;

// one workspace size that serves every reduction below on a [rows][c] tensor: BatchNorm statistics / backward partials,
// the 25-tap depthwise weight gradient, the stem weight gradient and the SE backward scratch
extern "C" size_t mt_extractor_train_workspace_bytes(long long rows, int c) {
  if (rows <= 0 || c <= 0) return 0;
  const RedGeom g = red_geom(rows, c);
  return (size_t)(g.nblk + 1) * (size_t)c * 25 * sizeof(double) + (size_t)c * 16 + (size_t)current_sms() * 2 * 864 * sizeof(double) +
         (size_t)65536 * 16;
}

extern "C" int mt_bn_stats(const float* x, float* mean, float* var, float* running_mean, float* running_var, float momentum,
                           long long rows, int c, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(x && mean && var && workspace && rows > 0 && c > 0, "bn_stats: bad argument");
  const RedGeom g = red_geom(rows, c);
  MT_REQUIRE(workspace_bytes >= (size_t)g.nblk * c * 2 * sizeof(double), "bn_stats: workspace too small");
  MT_REQUIRE(rows < (1LL << 31), "bn_stats: too many rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* part = reinterpret_cast<double*>(workspace);
  chan_sums_kernel<0><<<dim3(g.cblk, g.nblk), 256, 0, st>>>(x, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0.f, part,
                                                            (int)rows, c, g.per_block);
  MT_LAUNCH_CHECK("chan_sums_kernel");
  bn_stats_finish_kernel<<<(c + 127) / 128, 128, 0, st>>>(part, g.nblk, (int)rows, c, mean, var, running_mean, running_var,
                                                          momentum);
  MT_LAUNCH_CHECK("bn_stats_finish_kernel");
  return MT_OK;
}

extern "C" int mt_bn_act_fwd(const float* x, const float* mean, const float* var, const float* gamma, const float* beta,
                             int act, float eps, float* out, long long rows, int c, void* stream) {
  MT_REQUIRE(x && mean && var && gamma && beta && out && rows > 0 && c > 0 && (act == 0 || act == 1), "bn_act_fwd: bad argument");
  const size_t total = (size_t)rows * c;
  bn_act_fwd_kernel<<<ew_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, mean, var, gamma, beta, act, eps, out,
                                                                                        total, c);
  MT_LAUNCH_CHECK("bn_act_fwd_kernel");
  return MT_OK;
}

extern "C" int mt_bn_act_bwd(const float* dy, const float* x, const float* mean, const float* var, const float* gamma,
                             const float* beta, int act, float eps, float* dx, float* dgamma, float* dbeta, long long rows,
                             int c, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(dy && x && mean && var && gamma && beta && dx && dgamma && dbeta && workspace && rows > 0 && c > 0,
             "bn_act_bwd: bad argument");
  const RedGeom g = red_geom(rows, c);
  const size_t need = (size_t)g.nblk * c * 2 * sizeof(double) + (size_t)c * 2 * sizeof(float);
  MT_REQUIRE(workspace_bytes >= need && rows < (1LL << 31), "bn_act_bwd: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* part = reinterpret_cast<double*>(workspace);
  float* sums = reinterpret_cast<float*>(part + (size_t)g.nblk * c * 2);
  chan_sums_kernel<1><<<dim3(g.cblk, g.nblk), 256, 0, st>>>(x, dy, mean, var, gamma, beta, act, eps, part, (int)rows, c,
                                                            g.per_block);
  MT_LAUNCH_CHECK("chan_sums_kernel");
  bn_bwd_finish_kernel<<<(c + 127) / 128, 128, 0, st>>>(part, g.nblk, c, dgamma, dbeta, sums);
  MT_LAUNCH_CHECK("bn_bwd_finish_kernel");
  const size_t total = (size_t)rows * c;
  bn_act_bwd_kernel<<<ew_grid(total), 256, 0, st>>>(dy, x, mean, var, gamma, beta, sums, act, eps, dx, total, c,
                                                    1.0f / (float)rows);
  MT_LAUNCH_CHECK("bn_act_bwd_kernel");
  return MT_OK;
}

extern "C" int mt_stem_raw_fwd(const float* x, const float* w, float* out, int n_img, int h, int w_, void* stream) {
  MT_REQUIRE(x && w && out && n_img > 0 && h > 0 && w_ > 0, "stem_raw_fwd: bad argument");
  const int Ho = (h + 1) / 2, Wo = (w_ + 1) / 2;
  const size_t total = (size_t)n_img * Ho * Wo * 32;
  stem_raw_fwd_kernel<<<ew_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, w, out, n_img, h, w_, Ho, Wo,
                                                                                          same_pad_lo_t(h, 3, 2));
  MT_LAUNCH_CHECK("stem_raw_fwd_kernel");
  return MT_OK;
}

extern "C" int mt_stem_wgrad(const float* x, const float* dy, float* dw, int n_img, int h, int w_, void* workspace,
                             size_t workspace_bytes, void* stream) {
  MT_REQUIRE(x && dy && dw && workspace && n_img > 0, "stem_wgrad: bad argument");
  const int Ho = (h + 1) / 2, Wo = (w_ + 1) / 2;
  const long long rows = (long long)n_img * Ho * Wo;
  const int nblk = (int)std::min<long long>(rows, current_sms() * 2);
  const int per = (int)((rows + nblk - 1) / nblk);
  MT_REQUIRE(workspace_bytes >= (size_t)nblk * 864 * sizeof(double), "stem_wgrad: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* part = reinterpret_cast<double*>(workspace);
  stem_wgrad_kernel<<<nblk, 864, 0, st>>>(x, dy, part, n_img, h, w_, Ho, Wo, same_pad_lo_t(h, 3, 2), per);
  MT_LAUNCH_CHECK("stem_wgrad_kernel");
  reduce_partials_kernel<<<(864 + 127) / 128, 128, 0, st>>>(part, nblk, 864, dw);
  MT_LAUNCH_CHECK("reduce_partials_kernel");
  return MT_OK;
}

extern "C" int mt_dwconv_raw_fwd(const float* in, const float* w, float* out, int n_img, int h, int c, int k, int s,
                                 void* stream) {
  MT_REQUIRE(in && w && out && n_img > 0 && h > 0 && c > 0 && (k == 3 || k == 5) && (s == 1 || s == 2), "dwconv_raw_fwd: bad argument");
  const int Ho = (h + s - 1) / s;
  const size_t total = (size_t)n_img * Ho * Ho * c;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int pad = same_pad_lo_t(h, k, s), grid = ew_grid(total);
  if (k == 3 && s == 1) dw_raw_fwd_kernel<3, 1><<<grid, 256, 0, st>>>(in, w, out, n_img, h, Ho, c, pad);
  else if (k == 3) dw_raw_fwd_kernel<3, 2><<<grid, 256, 0, st>>>(in, w, out, n_img, h, Ho, c, pad);
  else if (s == 1) dw_raw_fwd_kernel<5, 1><<<grid, 256, 0, st>>>(in, w, out, n_img, h, Ho, c, pad);
  else dw_raw_fwd_kernel<5, 2><<<grid, 256, 0, st>>>(in, w, out, n_img, h, Ho, c, pad);
  MT_LAUNCH_CHECK("dw_raw_fwd_kernel");
  return MT_OK;
}

extern "C" int mt_dwconv_dgrad(const float* dy, const float* w, float* dx, int n_img, int h, int c, int k, int s, void* stream) {
  MT_REQUIRE(dy && w && dx && n_img > 0 && h > 0 && c > 0 && (k == 3 || k == 5) && (s == 1 || s == 2), "dwconv_dgrad: bad argument");
  const int Ho = (h + s - 1) / s;
  const size_t total = (size_t)n_img * h * h * c;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int pad = same_pad_lo_t(h, k, s), grid = ew_grid(total);
  if (k == 3 && s == 1) dw_dgrad_kernel<3, 1><<<grid, 256, 0, st>>>(dy, w, dx, n_img, h, Ho, c, pad);
  else if (k == 3) dw_dgrad_kernel<3, 2><<<grid, 256, 0, st>>>(dy, w, dx, n_img, h, Ho, c, pad);
  else if (s == 1) dw_dgrad_kernel<5, 1><<<grid, 256, 0, st>>>(dy, w, dx, n_img, h, Ho, c, pad);
  else dw_dgrad_kernel<5, 2><<<grid, 256, 0, st>>>(dy, w, dx, n_img, h, Ho, c, pad);
  MT_LAUNCH_CHECK("dw_dgrad_kernel");
  return MT_OK;
}

extern "C" int mt_dwconv_wgrad(const float* in, const float* dy, float* dw, int n_img, int h, int c, int k, int s,
                               void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(in && dy && dw && workspace && n_img > 0 && h > 0 && c > 0 && (k == 3 || k == 5) && (s == 1 || s == 2),
             "dwconv_wgrad: bad argument");
  const int Ho = (h + s - 1) / s;
  const long long npos = (long long)n_img * Ho * Ho;
  const RedGeom g = red_geom(npos, c);
  MT_REQUIRE(workspace_bytes >= (size_t)g.nblk * k * k * c * sizeof(double), "dwconv_wgrad: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* part = reinterpret_cast<double*>(workspace);
  const int pad = same_pad_lo_t(h, k, s);
  if (k == 3) dw_wgrad_kernel<3><<<dim3(g.cblk, g.nblk), 256, 0, st>>>(in, dy, part, n_img, h, Ho, c, s, pad, g.per_block);
  else dw_wgrad_kernel<5><<<dim3(g.cblk, g.nblk), 256, 0, st>>>(in, dy, part, n_img, h, Ho, c, s, pad, g.per_block);
  MT_LAUNCH_CHECK("dw_wgrad_kernel");
  reduce_partials_kernel<<<(k * k * c + 127) / 128, 128, 0, st>>>(part, g.nblk, k * k * c, dw);
  MT_LAUNCH_CHECK("reduce_partials_kernel");
  return MT_OK;
}

extern "C" size_t mt_conv1x1_wgrad_workspace_bytes(long long rows, int co, int ci) {
  if (rows <= 0 || co <= 0 || ci <= 0) return 0;
  return (size_t)wgrad_chunks(rows, co, ci) * co * ci * sizeof(double);
}

extern "C" int mt_conv1x1_wgrad(const float* dy, const float* a, float* dw, long long rows, int co, int ci, void* workspace,
                                size_t workspace_bytes, void* stream) {
  MT_REQUIRE(dy && a && dw && workspace && rows > 0 && co > 0 && ci > 0, "conv1x1_wgrad: bad argument");
  const int chunks = wgrad_chunks(rows, co, ci);
  MT_REQUIRE(workspace_bytes >= (size_t)chunks * co * ci * sizeof(double), "conv1x1_wgrad: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  double* part = reinterpret_cast<double*>(workspace);
  const long long per = ((rows + chunks - 1) / chunks + 31) / 32 * 32;
  conv1x1_wgrad_kernel<<<dim3((co + 31) / 32, (ci + 31) / 32, chunks), 256, 0, st>>>(dy, a, part, rows, co, ci, per);
  MT_LAUNCH_CHECK("conv1x1_wgrad_kernel");
  reduce_partials_kernel<<<(co * ci + 127) / 128, 128, 0, st>>>(part, chunks, co * ci, dw);
  MT_LAUNCH_CHECK("reduce_partials_kernel");
  return MT_OK;
}

extern "C" int mt_group_mean(const float* x, float* out, int groups, int rows, int c, void* stream) {
  MT_REQUIRE(x && out && groups > 0 && groups <= 65535 && rows > 0 && c > 0, "group_mean: bad argument");
  group_mean_kernel<<<dim3((c + 31) / 32, groups), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, out, rows, c);
  MT_LAUNCH_CHECK("group_mean_kernel");
  return MT_OK;
}

extern "C" int mt_se_fc_fwd(const float* mean, const float* wr, const float* br, const float* we, const float* be, float* gate,
                            float* s_pre, int n_img, int c, int sq, void* stream) {
  MT_REQUIRE(mean && wr && br && we && be && gate && s_pre && n_img > 0 && c > 0 && sq > 0 && sq <= 1024, "se_fc_fwd: bad argument");
  se_fc_fwd_kernel<<<n_img, 256, (size_t)sq * 4, reinterpret_cast<cudaStream_t>(stream)>>>(mean, wr, br, we, be, gate, s_pre, c, sq);
  MT_LAUNCH_CHECK("se_fc_fwd_kernel");
  return MT_OK;
}

extern "C" int mt_se_fc_bwd(const float* dgate, const float* gate, const float* s_pre, const float* mean, const float* wr,
                            const float* we, float* dmean, float* dwr, float* dbr, float* dwe, float* dbe, int n_img, int c,
                            int sq, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(dgate && gate && s_pre && mean && wr && we && dmean && dwr && dbr && dwe && dbe && workspace, "se_fc_bwd: null pointer");
  MT_REQUIRE(n_img > 0 && c > 0 && sq > 0 && (size_t)(c + sq) * 4 <= 48 * 1024, "se_fc_bwd: bad shape");
  const size_t need = (size_t)n_img * (c + 2 * sq) * sizeof(float);
  MT_REQUIRE(workspace_bytes >= need, "se_fc_bwd: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* de = reinterpret_cast<float*>(workspace);            // [n][C]
  float* dspre = de + (size_t)n_img * c;                      // [n][SQ]
  float* s = dspre + (size_t)n_img * sq;                      // [n][SQ]
  se_fc_bwd_kernel<<<n_img, 256, (size_t)(c + sq) * 4, st>>>(dgate, gate, s_pre, wr, we, de, dspre, s, dmean, c, sq);
  MT_LAUNCH_CHECK("se_fc_bwd_kernel");
  outer_sum_kernel<<<(c * sq + 255) / 256, 256, 0, st>>>(de, s, dwe, n_img, c, sq);          // dWe [C][SQ]
  MT_LAUNCH_CHECK("outer_sum_kernel");
  outer_sum_kernel<<<(c * sq + 255) / 256, 256, 0, st>>>(dspre, mean, dwr, n_img, sq, c);    // dWr [SQ][C]
  MT_LAUNCH_CHECK("outer_sum_kernel");
  int rc = mt_colsum_f32(de, dbe, n_img, c, 0, stream);
  if (rc) return rc;
  return mt_colsum_f32(dspre, dbr, n_img, sq, 0, stream);
}

extern "C" int mt_gate_mul(const float* x, const float* gate, float* out, int n_img, int rows, int c, void* stream) {
  MT_REQUIRE(x && gate && out && n_img > 0 && rows > 0 && c > 0, "gate_mul: bad argument");
  const size_t total = (size_t)n_img * rows * c;
  gate_mul_kernel<<<ew_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, gate, out, total, rows * c, c);
  MT_LAUNCH_CHECK("gate_mul_kernel");
  return MT_OK;
}

extern "C" int mt_gate_bwd(const float* dxg, const float* x, const float* gate, const float* dmean, float* dgate, float* dx,
                           int n_img, int rows, int c, int phase, void* stream) {
  MT_REQUIRE(dxg && n_img > 0 && n_img <= 65535 && rows > 0 && c > 0, "gate_bwd: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (phase == 0) {                       // dgate = sum_rows dxg * x
    MT_REQUIRE(x && dgate, "gate_bwd: null pointer");
    gate_bwd_reduce_kernel<<<dim3((c + 31) / 32, n_img), 256, 0, st>>>(dxg, x, dgate, rows, c);
    MT_LAUNCH_CHECK("gate_bwd_reduce_kernel");
  } else {                                // dx = dxg * gate + dmean / rows
    MT_REQUIRE(gate && dmean && dx, "gate_bwd: null pointer");
    const size_t total = (size_t)n_img * rows * c;
    gate_bwd_apply_kernel<<<ew_grid(total), 256, 0, st>>>(dxg, gate, dmean, dx, total, rows * c, c, 1.0f / (float)rows);
    MT_LAUNCH_CHECK("gate_bwd_apply_kernel");
  }
  return MT_OK;
}

extern "C" int mt_scale_add(const float* x, const float* scale, const float* skip, float* out, int n_img, long long per_img,
                            void* stream) {
  MT_REQUIRE(x && out && n_img > 0 && per_img > 0 && per_img < (1LL << 31), "scale_add: bad argument");
  const size_t total = (size_t)n_img * per_img;
  scale_add_kernel<<<ew_grid(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, scale, skip, out, total, (int)per_img);
  MT_LAUNCH_CHECK("scale_add_kernel");
  return MT_OK;
}
