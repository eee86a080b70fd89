// Backward pass of the Size-Invariant TimeSformer (the `loss.backward()` of the reference's train.py:376-378 for
// models/size_invariant_timesformer.py:224-276 with a frozen feature extractor, train.py:344-346).
//
// The contractions of the backward pass reuse the tcgen05 GEMMs of gemm.cu:
//   dgrad  dX[M][K]  = dY[M][N] * W[N][K]          -> mt_pointwise_fwd(a = dY, w = W^T [K][N])
//   wgrad  dW[N][K] += dY^T[N][M] * X[M][K]        -> mt_linear_residual_fwd(a = dY^T [N][Mp], w = X^T [K][Mp], x = dW)
// so this file holds what is not a GEMM: the operand preparation (cast / transpose / bias-gradient column sums in one
// pass), LayerNorm, GEGLU, the divided attention core with the identity mask (probabilities recomputed from the saved
// qkv, nothing of size groups x keys is stored by the forward), the embedding tables and the classification head.
// Everything is templated on the activation type T (float: exact path used for tight parity; bf16: training path).
// Reductions are fixed-order (partials + a column-sum kernel) except the embedding tables, which use fp32 atomics
// like torch's own embedding backward.
#include <float.h>

#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "attention_bwd_mma.cuh"
#include "common.cuh"

namespace mt {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

// ---------------------------------------------------------------------------------------------------
// out[c] = sum_r in[r][c]   (fixed order).  block = 32 columns x 8 row lanes, grid = ceil(C / 32)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C,
                                                     int accumulate) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < C)
    for (int r = ty; r < R; r += 8) s += in[(size_t)r * C + c];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    out[c] = accumulate ? out[c] + t : t;
  }
}

int launch_colsum(const float* in, float* out, int R, int C, int accumulate, cudaStream_t st) {
  colsum_kernel<<<(C + 31) / 32, 256, 0, st>>>(in, out, R, C, accumulate);
  MT_LAUNCH_CHECK("colsum_kernel");
  return MT_OK;
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward (PreNorm, size_invariant_timesformer.py:18-26).  One warp per row, persistent:
//   xhat = (x - mean) * rstd;  dyg = dy * gamma;  dx = rstd * (dyg - mean(dyg) - xhat * mean(dyg * xhat))
//   gx[row] += dx   (gx is the fp32 gradient of the residual stream: x_out = x + f(LN(x)))
//   part[block][0:dim] = sum_rows dy * xhat,  part[block][dim:2dim] = sum_rows dy
// ---------------------------------------------------------------------------------------------------
template <typename T, int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const T* __restrict__ dy, float* __restrict__ gx,
                                                            float* __restrict__ part, int rows, int dim) {
  __shared__ float red[8][NV * 128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[NV][4], dg[NV][4], db[NV][4];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    load4(gamma + (i * 32 + lane) * 4, g[i]);
#pragma unroll
    for (int k = 0; k < 4; ++k) { dg[i][k] = 0.f; db[i][k] = 0.f; }
  }
  const float inv_dim = 1.0f / (float)dim;
  // software pipeline: the next row's x / dy loads are in flight while this row is reduced and written
  const int stride = gridDim.x * 8;
  int row = blockIdx.x * 8 + warp;
  float nx[NV][4], nd[NV][4];
  if (row < rows) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      load4(x + (size_t)row * dim + (i * 32 + lane) * 4, nx[i]);
      load4(dy + (size_t)row * dim + (i * 32 + lane) * 4, nd[i]);
    }
  }
  for (; row < rows; row += stride) {
    float xv[NV][4], dv[NV][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { xv[i][k] = nx[i][k]; dv[i][k] = nd[i][k]; }
      s += (xv[i][0] + xv[i][1]) + (xv[i][2] + xv[i][3]);
    }
    if (row + stride < rows) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        load4(x + (size_t)(row + stride) * dim + (i * 32 + lane) * 4, nx[i]);
        load4(dy + (size_t)(row + stride) * dim + (i * 32 + lane) * 4, nd[i]);
      }
    }
    const float mean = warp_sum(s) * inv_dim;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) { xv[i][k] -= mean; q += xv[i][k] * xv[i][k]; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * inv_dim + 1e-5f);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        xv[i][k] *= rstd;                       // xhat
        const float dyg = dv[i][k] * g[i][k];
        c1 += dyg;
        c2 = fmaf(dyg, xv[i][k], c2);
        dg[i][k] = fmaf(dv[i][k], xv[i][k], dg[i][k]);
        db[i][k] += dv[i][k];
      }
    c1 = warp_sum(c1) * inv_dim;
    c2 = warp_sum(c2) * inv_dim;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float* gp = gx + (size_t)row * dim + (i * 32 + lane) * 4;
      float4 o = *reinterpret_cast<float4*>(gp);
      o.x += rstd * (dv[i][0] * g[i][0] - c1 - xv[i][0] * c2);
      o.y += rstd * (dv[i][1] * g[i][1] - c1 - xv[i][1] * c2);
      o.z += rstd * (dv[i][2] * g[i][2] - c1 - xv[i][2] * c2);
      o.w += rstd * (dv[i][3] * g[i][3] - c1 - xv[i][3] * c2);
      *reinterpret_cast<float4*>(gp) = o;
    }
  }
  // fixed-order reduction over the 8 warps of the block, dgamma then dbeta
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) red[warp][i * 128 + lane * 4 + k] = pass == 0 ? dg[i][k] : db[i][k];
    __syncthreads();
    for (int c = threadIdx.x; c < dim; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][c];
      part[(size_t)blockIdx.x * 2 * dim + pass * dim + c] = t;
    }
  }
}

int ln_bwd_grid(int rows) { return std::min((rows + 7) / 8, sm_count() * 2); }   // 2 resident blocks per SM (120 registers)

template <typename T>
int launch_ln_bwd(const float* x, const float* gamma, const void* dy, float* gx, float* part, int rows, int dim,
                  cudaStream_t st) {
  const int grid = ln_bwd_grid(rows);
  const T* d = reinterpret_cast<const T*>(dy);
  switch (dim >> 7) {
    case 1: layernorm_bwd_kernel<T, 1><<<grid, 256, 0, st>>>(x, gamma, d, gx, part, rows, dim); break;
    case 2: layernorm_bwd_kernel<T, 2><<<grid, 256, 0, st>>>(x, gamma, d, gx, part, rows, dim); break;
    case 3: layernorm_bwd_kernel<T, 3><<<grid, 256, 0, st>>>(x, gamma, d, gx, part, rows, dim); break;
    case 4: layernorm_bwd_kernel<T, 4><<<grid, 256, 0, st>>>(x, gamma, d, gx, part, rows, dim); break;
    case 5: layernorm_bwd_kernel<T, 5><<<grid, 256, 0, st>>>(x, gamma, d, gx, part, rows, dim); break;
    case 6: layernorm_bwd_kernel<T, 6><<<grid, 256, 0, st>>>(x, gamma, d, gx, part, rows, dim); break;
    case 7: layernorm_bwd_kernel<T, 7><<<grid, 256, 0, st>>>(x, gamma, d, gx, part, rows, dim); break;
    default: layernorm_bwd_kernel<T, 8><<<grid, 256, 0, st>>>(x, gamma, d, gx, part, rows, dim); break;
  }
  MT_LAUNCH_CHECK("layernorm_bwd_kernel");
  return MT_OK;
}

// ---------------------------------------------------------------------------------------------------
// Operand preparation for the backward GEMMs, one pass over src [M][C] (TIn = float for the fp32 residual-stream
// gradient, T for activations):  out_rm T [M][C] (cast), out_t T [C][Mp] (transpose, columns m >= M zero: the wgrad
// GEMM contracts over Mp), colpart f32 [Mp/64][C] (per-tile column sums -> bias gradient).  Each output is optional.
// rows_per_batch > 0 drops the CLS row of every video: source row = m + m / rows_per_batch + 1 (patch-embedding wgrad).
// 64 x 64 tiles, 256 threads.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 load2f(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 load2f(const bf16* p) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
}
__device__ __forceinline__ void store2f(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store2f(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// Each thread moves element pairs: 128-byte (bf16) / 256-byte (f32) row segments per warp on the way in, 128-byte
// segments of the transposed rows on the way out.
template <typename TIn, typename T>
__global__ void __launch_bounds__(256) grad_prep_kernel(const TIn* __restrict__ src, T* __restrict__ out_rm,
                                                        T* __restrict__ out_t, float* __restrict__ colpart, int M, int C,
                                                        int Mp, int rows_per_batch) {
  __shared__ float tile[64][65];
  const int c0 = blockIdx.x * 64, m0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll 4
  for (int i = 0; i < 8; ++i) {
    const int mm = i * 8 + ty, m = m0 + mm;
    float2 v = make_float2(0.f, 0.f);
    if (m < M) {
      const size_t srow = rows_per_batch ? (size_t)m + m / rows_per_batch + 1 : (size_t)m;
      v = load2f(src + srow * C + c0 + tx * 2);
      if (out_rm) store2f(out_rm + (size_t)m * C + c0 + tx * 2, v.x, v.y);
    }
    tile[mm][tx * 2] = v.x;
    tile[mm][tx * 2 + 1] = v.y;
  }
  __syncthreads();
  if (out_t) {
#pragma unroll 4
    for (int i = 0; i < 8; ++i) {
      const int cc = i * 8 + ty;
      store2f(out_t + (size_t)(c0 + cc) * Mp + m0 + tx * 2, tile[tx * 2][cc], tile[tx * 2 + 1][cc]);
    }
  }
  if (colpart && threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll 8
    for (int mm = 0; mm < 64; ++mm) s += tile[mm][threadIdx.x];
    colpart[(size_t)blockIdx.y * C + c0 + threadIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------------
// GEGLU (size_invariant_timesformer.py:60-63) on the interleaved layout of mt_ff_weights_t (blocks of 64 columns of h =
// 32 value columns followed by their 32 gate columns):  out = u * gelu_erf(g)
//   backward: du = dout * gelu(g);  dg = dout * u * (Phi(g) + g * phi(g))
// One thread per 8 output columns.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) geglu_fwd_kernel(const T* __restrict__ h, T* __restrict__ out, size_t total8,
                                                        int hd) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total8) return;
  const int per_row = hd >> 3;
  const size_t row = idx / per_row;
  const int oc = (int)(idx - row * per_row) * 8;
  const int hc = (oc >> 5) * 64 + (oc & 31);
  float u[8], g[8];
  load8(h + row * 2 * hd + hc, u);
  load8(h + row * 2 * hd + hc + 32, g);
  if (sizeof(T) == 4) {                       // exact path: erff
#pragma unroll
    for (int i = 0; i < 8; ++i) u[i] *= gelu_erf(g[i]);
  } else {                                    // bf16 path: the packed form of the fused GEMM epilogue (inference) -- one arithmetic
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 r = geglu2(make_float2(u[2 * i], u[2 * i + 1]), make_float2(g[2 * i], g[2 * i + 1]));
      u[2 * i] = r.x; u[2 * i + 1] = r.y;
    }
  }
  store8(out + row * hd + oc, u);
}

template <typename T>
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const T* __restrict__ h, const T* __restrict__ dout,
                                                        T* __restrict__ dh, size_t total8, int hd) {
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total8) return;
  const int per_row = hd >> 3;
  const size_t row = idx / per_row;
  const int oc = (int)(idx - row * per_row) * 8;
  const int hc = (oc >> 5) * 64 + (oc & 31);
  float u[8], g[8], d[8];
  load8(h + row * 2 * hd + hc, u);
  load8(h + row * 2 * hd + hc + 32, g);
  load8(dout + row * hd + oc, d);
  if (sizeof(T) == 4) {                       // exact path
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float cdf = 0.5f * (1.0f + erff(g[i] * 0.70710678118654752f));
      const float pdf = 0.39894228040143268f * expf(-0.5f * g[i] * g[i]);
      const float du = d[i] * g[i] * cdf;
      const float dg = d[i] * u[i] * fmaf(g[i], pdf, cdf);
      u[i] = du;
      g[i] = dg;
    }
  } else {                                    // bf16 path: packed math, one erf evaluation for cdf and pdf
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 g2 = make_float2(g[2 * i], g[2 * i + 1]), u2 = make_float2(u[2 * i], u[2 * i + 1]);
      const float2 d2 = make_float2(d[2 * i], d[2 * i + 1]);
      float2 cdf, pdf;
      gelu_cdf_pdf2(g2, cdf, pdf);
      const float2 du = __fmul2_rn(__fmul2_rn(d2, g2), cdf);
      const float2 dg = __fmul2_rn(__fmul2_rn(d2, u2), __ffma2_rn(g2, pdf, cdf));
      u[2 * i] = du.x; u[2 * i + 1] = du.y;
      g[2 * i] = dg.x; g[2 * i + 1] = dg.y;
    }
  }
  store8(dh + row * 2 * hd + hc, u);
  store8(dh + row * 2 * hd + hc + 32, g);
}

// The same with the bias gradient of net.0 folded in: a block walks rows blockIdx.x, + gridDim.x, ... with one thread
// per 8-column group and keeps the column sums of (du | dg) in registers; part[block][2*hd] are summed over the blocks
// in a fixed order by colsum_kernel (deterministic).  Saves the separate pass over dh (205 MB at B = 32).
template <typename T>
__global__ void __launch_bounds__(256) geglu_bwd_colsum_kernel(const T* __restrict__ h, const T* __restrict__ dout,
                                                               T* __restrict__ dh, float* __restrict__ part, int rows,
                                                               int hd) {
  const int cg = blockIdx.y * 256 + threadIdx.x;
  if (cg >= (hd >> 3)) return;
  const int oc = cg * 8;
  const int hc = (oc >> 5) * 64 + (oc & 31);
  float su[8], sg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { su[i] = 0.f; sg[i] = 0.f; }
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    float u[8], g[8], d[8];
    load8(h + (size_t)row * 2 * hd + hc, u);
    load8(h + (size_t)row * 2 * hd + hc + 32, g);
    load8(dout + (size_t)row * hd + oc, d);
    if (sizeof(T) == 4) {                     // exact path
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float cdf = 0.5f * (1.0f + erff(g[i] * 0.70710678118654752f));
        const float pdf = 0.39894228040143268f * expf(-0.5f * g[i] * g[i]);
        const float du = d[i] * g[i] * cdf;
        const float dg = d[i] * u[i] * fmaf(g[i], pdf, cdf);
        u[i] = du;
        g[i] = dg;
      }
    } else {                                  // bf16 path: packed math (same arithmetic as geglu_bwd_kernel<bf16>)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 g2 = make_float2(g[2 * i], g[2 * i + 1]), u2 = make_float2(u[2 * i], u[2 * i + 1]);
        const float2 d2 = make_float2(d[2 * i], d[2 * i + 1]);
        float2 cdf, pdf;
        gelu_cdf_pdf2(g2, cdf, pdf);
        const float2 du = __fmul2_rn(__fmul2_rn(d2, g2), cdf);
        const float2 dg = __fmul2_rn(__fmul2_rn(d2, u2), __ffma2_rn(g2, pdf, cdf));
        u[2 * i] = du.x; u[2 * i + 1] = du.y;
        g[2 * i] = dg.x; g[2 * i + 1] = dg.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { su[i] += u[i]; sg[i] += g[i]; }
    store8(dh + (size_t)row * 2 * hd + hc, u);
    store8(dh + (size_t)row * 2 * hd + hc + 32, g);
  }
  float* p = part + (size_t)blockIdx.x * 2 * hd + hc;
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = su[i]; p[32 + i] = sg[i]; }
}

// ---------------------------------------------------------------------------------------------------
// Attention core backward (Attention.forward :114-141, attn :80-87), probabilities recomputed from qkv.
// Workspace (fp32):  ws_kv [B*heads][N][2]    (dS_j, p_j) of the CLS query for every key j: its contribution to dK_j / dV_j is
//                                             the rank-1 pair dS_j * q_cls / p_j * dO_cls, rebuilt by the consumers (round 1
//                                             wrote the 128 products per key: 103 MB per launch at B = 32)
//                    ws_q  [B*heads][64]      dQ of the CLS query
//                    ws_cls[B*heads][G][128]  per-group contribution to dK / dV of the CLS key
// (1) attn_cls_bwd_kernel   : CLS query over all N keys            -> ws_kv, ws_q
// (2) attn_group_bwd_kernel : one block per (b, h, group)          -> dq, dk, dv of the group's tokens (+ ws_kv), ws_cls
// (3) attn_cls_finish_kernel: dq, dk, dv of token 0 = ws_q, ws_kv[0] + sum_g ws_cls
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) attn_cls_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                           const uint8_t* __restrict__ mask, float* __restrict__ ws_kv,
                                                           float* __restrict__ ws_q, int N, int f, int n, int heads) {
  extern __shared__ float sm[];
  float* sc = sm;              // [N] probabilities
  float* dp = sm + N;          // [N] dP, then dS
  float* q0 = dp + N;          // [64]
  float* d0 = q0 + 64;         // [64] dO of the CLS row
  float* red = d0 + 64;        // [32]
  float* part = red + 32;      // [32][64]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int inner = heads * 64, ld = 3 * inner;
  const T* base = qkv + (size_t)b * N * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 64) {
    q0[tid] = to_f(base[h * 64 + tid]);
    d0[tid] = to_f(dout[(size_t)b * N * inner + h * 64 + tid]);
  }
  __syncthreads();
  float lmax = -FLT_MAX;
  for (int j = tid; j < N; j += 256) {
    const T* kr = base + (size_t)j * ld + inner + h * 64;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float kv[8];
      load8(kr + c * 8, kv);
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(q0[c * 8 + i], kv[i], s);
    }
    if (j > 0 && !mask[b * f + (j - 1) / n]) s = -FLT_MAX;
    sc[j] = s;
    lmax = fmaxf(lmax, s);
  }
  lmax = warp_max(lmax);
  if (lane == 0) red[warp] = lmax;
  __syncthreads();
  float gmax = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) gmax = fmaxf(gmax, red[i]);
  __syncthreads();
  float lsum = 0.f;
  for (int j = tid; j < N; j += 256) {
    const float e = expf(sc[j] - gmax);
    sc[j] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  float gsum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) gsum += red[i];
  const float inv = 1.0f / gsum;
  __syncthreads();
  // dP_j = dO . V_j ;  D = sum_j p_j dP_j
  float ld_ = 0.f;
  for (int j = tid; j < N; j += 256) {
    const T* vr = base + (size_t)j * ld + 2 * inner + h * 64;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float vv[8];
      load8(vr + c * 8, vv);
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(d0[c * 8 + i], vv[i], s);
    }
    const float pj = sc[j] * inv;
    sc[j] = pj;
    dp[j] = s;
    ld_ = fmaf(pj, s, ld_);
  }
  ld_ = warp_sum(ld_);
  if (lane == 0) red[warp] = ld_;
  __syncthreads();
  float D = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) D += red[i];
  for (int j = tid; j < N; j += 256) dp[j] = sc[j] * (dp[j] - D);   // dS_j (each thread rewrites its own entries)
  __syncthreads();
  // (dS_j, p_j) per key for the consumers; thread = (key lane kg of 32, 8-dim chunk dc): dQ0 += dS_j K_j
  const int dc = tid & 7, kg = tid >> 3;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int j = tid; j < N; j += 256)
    *reinterpret_cast<float2*>(ws_kv + ((size_t)blockIdx.x * N + j) * 2) = make_float2(dp[j], sc[j]);
  for (int j = kg; j < N; j += 32) {
    float kv[8];
    load8(base + (size_t)j * ld + inner + h * 64 + dc * 8, kv);
    const float ds = dp[j];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(ds, kv[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[kg * 64 + dc * 8 + i] = acc[i];
  __syncthreads();
  if (tid < 64) {
    float r = 0.f;
#pragma unroll 8
    for (int g = 0; g < 32; ++g) r += part[g * 64 + tid];
    ws_q[(size_t)blockIdx.x * 64 + tid] = r;
  }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(128) attn_group_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                             const uint8_t* __restrict__ mask,
                                                             const uint8_t* __restrict__ idmask, T* __restrict__ dqkv,
                                                             const float* __restrict__ ws_kv, float* __restrict__ ws_cls,
                                                             int f, int n, int heads) {
  extern __shared__ float sm[];
  const int G = MODE == MT_ATTN_TIME ? n : f;
  const int Gq = MODE == MT_ATTN_TIME ? f : n;
  const int Gk = Gq + 1;
  float* Ks = sm;                  // [Gk][65]
  float* Vs = Ks + Gk * 65;        // [Gk][65]
  float* Qs = Vs + Gk * 65;        // [Gq][65]
  float* Os = Qs + Gq * 65;        // [Gq][65]  dO
  float* Ps = Os + Gq * 65;        // [Gq][64]  probabilities
  float* Ss = Ps + Gq * 64;        // [Gq][64]  dS
  const int g = blockIdx.x % G;
  const int h = (blockIdx.x / G) % heads;
  const int b = blockIdx.x / (G * heads);
  const int N = 1 + f * n, inner = heads * 64, ld = 3 * inner;
  const T* base = qkv + (size_t)b * N * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto token = [&](int j) -> int {
    if (j == 0) return 0;
    return MODE == MT_ATTN_TIME ? 1 + (j - 1) * n + g : 1 + g * n + (j - 1);
  };
  for (int e = tid; e < Gk * 8; e += 128) {
    const int j = e >> 3, c = e & 7;
    const T* row = base + (size_t)token(j) * ld + h * 64 + c * 8;
    float kv[8], vv[8];
    load8(row + inner, kv);
    load8(row + 2 * inner, vv);
#pragma unroll
    for (int i = 0; i < 8; ++i) { Ks[j * 65 + c * 8 + i] = kv[i]; Vs[j * 65 + c * 8 + i] = vv[i]; }
    if (j > 0) {
      float qv[8], ov[8];
      load8(row, qv);
      load8(dout + ((size_t)b * N + token(j)) * inner + h * 64 + c * 8, ov);
#pragma unroll
      for (int i = 0; i < 8; ++i) { Qs[(j - 1) * 65 + c * 8 + i] = qv[i]; Os[(j - 1) * 65 + c * 8 + i] = ov[i]; }
    }
  }
  __syncthreads();
  // phase 1: one warp per query: p, dP, dS, dQ
  for (int i = warp; i < Gq; i += 4) {
    const float* qi = Qs + i * 65;
    const float* oi = Os + i * 65;
    float s[2], dpv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int j = lane + r * 32;
      float a = -FLT_MAX, d = 0.f;
      if (j < Gk) {
        a = 0.f;
#pragma unroll 16
        for (int c = 0; c < 64; ++c) {
          a = fmaf(qi[c], Ks[j * 65 + c], a);
          d = fmaf(oi[c], Vs[j * 65 + c], d);
        }
        if (MODE == MT_ATTN_TIME && j > 0) {
          const bool ok = mask[b * f + (j - 1)] && idmask[((size_t)b * f + i) * f + (j - 1)];
          if (!ok) a = -FLT_MAX;
        }
      }
      s[r] = a;
      dpv[r] = d;
    }
    const float mx = warp_max(fmaxf(s[0], s[1]));
    const float e0 = (lane < Gk) ? expf(s[0] - mx) : 0.f;
    const float e1 = (lane + 32 < Gk) ? expf(s[1] - mx) : 0.f;
    const float inv = 1.0f / warp_sum(e0 + e1);
    const float p0 = e0 * inv, p1 = e1 * inv;
    const float D = warp_sum(fmaf(p0, dpv[0], p1 * dpv[1]));
    Ps[i * 64 + lane] = p0;
    Ps[i * 64 + lane + 32] = p1;
    Ss[i * 64 + lane] = p0 * (dpv[0] - D);
    Ss[i * 64 + lane + 32] = p1 * (dpv[1] - D);
    __syncwarp();
    float q0 = 0.f, q1 = 0.f;
    for (int j = 0; j < Gk; ++j) {
      const float ds = Ss[i * 64 + j];
      q0 = fmaf(ds, Ks[j * 65 + lane], q0);
      q1 = fmaf(ds, Ks[j * 65 + lane + 32], q1);
    }
    T* qrow = dqkv + ((size_t)b * N + token(i + 1)) * ld + h * 64;
    qrow[lane] = from_f<T>(q0);
    qrow[lane + 32] = from_f<T>(q1);
  }
  __syncthreads();
  // phase 2: dK_j = sum_i dS_ij Q_i, dV_j = sum_i P_ij dO_i  (+ the CLS-row contribution of ws_kv), fixed order
  const float* wkv = ws_kv + (size_t)(b * heads + h) * N * 2;
  const T* q_cls = qkv + (size_t)b * N * ld + h * 64;                  // the CLS query (pre-scaled) and its dO
  const T* d_cls = dout + (size_t)b * N * inner + h * 64;
  for (int e = tid; e < Gk * 64; e += 128) {
    const int j = e >> 6, d = e & 63;
    float dk = 0.f, dv = 0.f;
    for (int i = 0; i < Gq; ++i) {
      dk = fmaf(Ss[i * 64 + j], Qs[i * 65 + d], dk);
      dv = fmaf(Ps[i * 64 + j], Os[i * 65 + d], dv);
    }
    if (j == 0) {
      float* w = ws_cls + ((size_t)(b * heads + h) * G + g) * 128;
      w[d] = dk;
      w[64 + d] = dv;
    } else {
      const int tok = token(j);
      const float2 sp = *reinterpret_cast<const float2*>(wkv + (size_t)tok * 2);
      dk = fmaf(sp.x, to_f(q_cls[d]), dk);
      dv = fmaf(sp.y, to_f(d_cls[d]), dv);
      T* row = dqkv + ((size_t)b * N + tok) * ld + h * 64 + d;
      row[inner] = from_f<T>(dk);
      row[2 * inner] = from_f<T>(dv);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(64) attn_cls_finish_kernel(T* __restrict__ dqkv, const T* __restrict__ qkv,
                                                             const T* __restrict__ dout, const float* __restrict__ ws_kv,
                                                             const float* __restrict__ ws_q,
                                                             const float* __restrict__ ws_cls, int N, int G, int heads) {
  const int b = blockIdx.x / heads, h = blockIdx.x % heads, d = threadIdx.x;
  const int inner = heads * 64, ld = 3 * inner;
  const float2 sp = *reinterpret_cast<const float2*>(ws_kv + (size_t)blockIdx.x * N * 2);    // the CLS key seen by the CLS query
  float dk = sp.x * to_f(qkv[(size_t)b * N * ld + h * 64 + d]), dv = sp.y * to_f(dout[(size_t)b * N * inner + h * 64 + d]);
  for (int g = 0; g < G; ++g) {
    dk += ws_cls[((size_t)blockIdx.x * G + g) * 128 + d];
    dv += ws_cls[((size_t)blockIdx.x * G + g) * 128 + 64 + d];
  }
  T* row = dqkv + (size_t)b * N * ld + h * 64 + d;
  row[0] = from_f<T>(ws_q[(size_t)blockIdx.x * 64 + d]);
  row[inner] = from_f<T>(dk);
  row[2 * inner] = from_f<T>(dv);
}

struct AttnBwdWs { size_t kv, q, cls, total; };
AttnBwdWs attn_bwd_ws(int B, int f, int n, int heads) {
  const size_t N = 1 + (size_t)f * n, bh = (size_t)B * heads;
  AttnBwdWs l;
  l.kv = 0;
  l.q = l.kv + ((bh * N * 2 * 4 + 255) & ~(size_t)255);
  l.cls = l.q + bh * 64 * 4;
  l.total = l.cls + bh * (size_t)std::max(f, n) * 128 * 4;
  return l;
}

template <typename T>
int launch_attn_bwd(const void* qkv_, const void* dout_, const uint8_t* mask, const uint8_t* idmask, int mode,
                    void* dqkv_, int B, int f, int n, int heads, char* ws, cudaStream_t st) {
  const int N = 1 + f * n;
  const T* qkv = reinterpret_cast<const T*>(qkv_);
  const T* dout = reinterpret_cast<const T*>(dout_);
  T* dqkv = reinterpret_cast<T*>(dqkv_);
  const AttnBwdWs l = attn_bwd_ws(B, f, n, heads);
  float* ws_kv = reinterpret_cast<float*>(ws + l.kv);
  float* ws_q = reinterpret_cast<float*>(ws + l.q);
  float* ws_cls = reinterpret_cast<float*>(ws + l.cls);
  {
    // K of every token read twice (scores, dQ), V once; two floats per key written
    ProfScope prof(st, 10.0 * B * heads * 64.0 * N, (double)B * N * heads * (64.0 * 3 * sizeof(T) + 8.0), "attn_cls_bwd");
    const size_t smem = (size_t)(2 * N + 64 + 64 + 32 + 2048) * sizeof(float);
    auto kern = attn_cls_bwd_kernel<T>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(attn_cls_bwd)");
    }
    kern<<<B * heads, 256, smem, st>>>(qkv, dout, mask, ws_kv, ws_q, N, f, n, heads);
    MT_LAUNCH_CHECK("attn_cls_bwd_kernel");
  }
  const int G = mode == MT_ATTN_TIME ? n : f, Gq = mode == MT_ATTN_TIME ? f : n, Gk = Gq + 1;
  {
    ProfScope prof(st, 10.0 * B * heads * G * 64.0 * Gq * Gk, (double)B * N * heads * (64.0 * 7 * sizeof(T) + 512.0),   // q, k, v, dO in; dq, dk, dv out; workspace row in
                   mode == MT_ATTN_TIME ? "attn_time_bwd" : "attn_space_bwd");
    const size_t smem = (size_t)(2 * Gk * 65 + 2 * Gq * 65 + 2 * Gq * 64) * sizeof(float);
    int rc_mma = MT_ERR_UNSUPPORTED;
    if constexpr (std::is_same<T, bf16>::value) {
      // bf16 path: warp-level tensor-core kernel (attention_bwd_mma.cuh); MINTIME_B200_ATTN_BWD=simt keeps the FFMA one
      static int use_mma = -1;
      if (use_mma < 0) {
        const char* e = getenv("MINTIME_B200_ATTN_BWD");
        use_mma = (e && e[0] == 's') ? 0 : 1;
      }
      if (use_mma)
        rc_mma = attn::launch_attn_group_bwd_mma(mode, qkv, dout, mask, idmask, dqkv, ws_kv, ws_cls, B, f, n, heads, st);
      if (rc_mma != MT_OK && rc_mma != MT_ERR_UNSUPPORTED) return rc_mma;
    }
    if (rc_mma == MT_OK) {
      // (launched above)
    } else if (mode == MT_ATTN_TIME) {
      auto kern = attn_group_bwd_kernel<T, MT_ATTN_TIME>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024));
      if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(attn_time_bwd)");
      kern<<<B * heads * G, 128, smem, st>>>(qkv, dout, mask, idmask, dqkv, ws_kv, ws_cls, f, n, heads);
    } else {
      auto kern = attn_group_bwd_kernel<T, MT_ATTN_SPACE>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024));
      if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(attn_space_bwd)");
      kern<<<B * heads * G, 128, smem, st>>>(qkv, dout, mask, idmask, dqkv, ws_kv, ws_cls, f, n, heads);
    }
    if (rc_mma != MT_OK) MT_LAUNCH_CHECK("attn_group_bwd_kernel");
  }
  attn_cls_finish_kernel<T><<<B * heads, 64, 0, st>>>(dqkv, qkv, dout, ws_kv, ws_q, ws_cls, N, G, heads);
  MT_LAUNCH_CHECK("attn_cls_finish_kernel");
  return MT_OK;
}

// ---------------------------------------------------------------------------------------------------
// Token-build backward (size_invariant_timesformer.py:225-248): gradient of the embedding tables and the CLS token
// from g0 = dL/dx0 (fp32 [B][1+f*n][dim]).  One block per (video, frame) + one per video for the CLS row; the n patch
// rows of a frame share their size-embedding row, so that sum is taken in registers first.  fp32 atomics: rows of
// different videos collide (like torch's embedding_dense_backward the summation order is not fixed).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) embed_bwd_kernel(const float* __restrict__ g0,
                                                        const long long* __restrict__ positions,
                                                        const int* __restrict__ size_idx, float* __restrict__ dpos,
                                                        float* __restrict__ dsize, float* __restrict__ dcls, int f, int n,
                                                        int dim, int table_rows) {
  const int b = blockIdx.x / (f + 1), fr = blockIdx.x % (f + 1);
  const int N = 1 + f * n;
  const float* gb = g0 + (size_t)b * N * dim;
  for (int c = threadIdx.x; c < dim; c += 128) {
    if (fr == f) {
      const float v = gb[c];
      const long long p = positions ? positions[(size_t)b * N] : 0;
      if ((unsigned long long)p >= (unsigned long long)table_rows) __trap();      // (the forward already trapped on it)
      if (dpos) atomicAdd(dpos + (size_t)p * dim + c, v);
      if (dsize) atomicAdd(dsize + c, v);
      atomicAdd(dcls + c, v);
    } else {
      float acc = 0.f;
      for (int t = 0; t < n; ++t) {
        const int tok = 1 + fr * n + t;
        const float v = gb[(size_t)tok * dim + c];
        const long long p = positions ? positions[(size_t)b * N + tok] : (long long)tok;
        if ((unsigned long long)p >= (unsigned long long)table_rows) __trap();
        if (dpos) atomicAdd(dpos + (size_t)p * dim + c, v);
        acc += v;
      }
      if (dsize) {
        const int si = size_idx[b * f + fr];
        if ((unsigned)si >= (unsigned)table_rows) __trap();
        atomicAdd(dsize + (size_t)si * dim + c, acc);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Head backward (to_out = LayerNorm + Linear on x[:, 0], :195-198, :270): one block per video.
//   part[b] = [ dW (classes x dim) | db (classes) | dgamma (dim) | dbeta (dim) ];  gx[b][0][:] = dL/dx[b][0]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                       const float* __restrict__ be, const float* __restrict__ w,
                                                       const float* __restrict__ dlogits, float* __restrict__ gx,
                                                       float* __restrict__ part, int tokens, int dim, int classes) {
  extern __shared__ float sm[];
  float* xh = sm;            // [dim] xhat
  float* dxn = sm + dim;     // [dim] gradient of the normalised row
  __shared__ float red[4];
  const int bi = blockIdx.x;
  const float* xr = x + (size_t)bi * tokens * dim;
  const float* dl = dlogits + (size_t)bi * classes;
  float* pb = part + (size_t)bi * ((size_t)classes * dim + classes + 2 * dim);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    return (red[0] + red[1]) + (red[2] + red[3]);
  };
  float s = 0.f;
  for (int c = tid; c < dim; c += 128) s += xr[c];
  const float mean = block_sum(s) / (float)dim;
  float q = 0.f;
  for (int c = tid; c < dim; c += 128) { const float d = xr[c] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(block_sum(q) / (float)dim + 1e-5f);
  float c1 = 0.f, c2 = 0.f;
  for (int c = tid; c < dim; c += 128) {
    const float xhat = (xr[c] - mean) * rstd;
    float d = 0.f;
    for (int k = 0; k < classes; ++k) d = fmaf(dl[k], w[(size_t)k * dim + c], d);
    xh[c] = xhat;
    dxn[c] = d;
    const float xn = xhat * g[c] + be[c];
    for (int k = 0; k < classes; ++k) pb[(size_t)k * dim + c] = dl[k] * xn;
    pb[(size_t)classes * dim + classes + c] = d * xhat;
    pb[(size_t)classes * dim + classes + dim + c] = d;
    c1 += d * g[c];
    c2 = fmaf(d * g[c], xhat, c2);
  }
  for (int k = tid; k < classes; k += 128) pb[(size_t)classes * dim + k] = dl[k];
  c1 = block_sum(c1) / (float)dim;
  c2 = block_sum(c2) / (float)dim;
  for (int c = tid; c < dim; c += 128)
    gx[(size_t)bi * tokens * dim + c] = rstd * (dxn[c] * g[c] - c1 - xh[c] * c2);
}

}  // namespace
}  // namespace mt

using namespace mt;

extern "C" int mt_linear_wgrad(int precision, const void* dy_t, const void* x_t, float* dw, int n_out, int k_in, int mp,
                               void* stream) {
  MT_REQUIRE(dy_t && x_t && dw && n_out > 0 && k_in > 0 && mp > 0, "linear_wgrad: bad argument");
  GemmArgs g{};
  g.a = dy_t; g.w = x_t; g.M = n_out; g.N = k_in; g.K = mp;
  g.epi.kind = EPI_RESID_F32; g.epi.M = n_out; g.epi.N = k_in; g.epi.bias = nullptr; g.epi.out = dw; g.epi.ldo = k_in;
  // few output tiles, a very long contraction: spread the k-blocks of every tile over the SMs
  const int tiles = ((n_out + 127) / 128) * ((k_in + 255) / 256);
  g.splits = std::max(1, (sm_count() + tiles / 2) / tiles);
  return launch_gemm(precision, g, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mt_linear_wgrad_nt(int precision, const void* dy, const void* x, float* dw, int n_out, int k_in, int m,
                                  void* stream) {
  MT_REQUIRE(dy && x && dw && n_out > 0 && k_in > 0 && m > 0, "linear_wgrad_nt: bad argument");
  MT_REQUIRE(precision == MT_PREC_BF16 && n_out % 8 == 0 && k_in % 8 == 0,
             "linear_wgrad_nt: bf16 only, n_out (%d) and k_in (%d) multiples of 8", n_out, k_in);
  GemmArgs g{};
  g.a = dy; g.w = x; g.M = n_out; g.N = k_in; g.K = m;
  g.mn_major = 1;
  g.epi.kind = EPI_RESID_F32; g.epi.M = n_out; g.epi.N = k_in; g.epi.bias = nullptr; g.epi.out = dw; g.epi.ldo = k_in;
  const int tiles = ((n_out + 127) / 128) * ((k_in + 255) / 256);
  g.splits = std::max(1, (sm_count() + tiles / 2) / tiles);
  return launch_gemm(precision, g, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mt_colsum_f32(const float* in, float* out, int rows, int cols, int accumulate, void* stream) {
  MT_REQUIRE(in && out && rows > 0 && cols > 0, "colsum: bad argument");
  return launch_colsum(in, out, rows, cols, accumulate, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" size_t mt_layernorm_bwd_workspace_bytes(int rows, int dim) {
  return (size_t)ln_bwd_grid(rows > 0 ? rows : 1) * 2 * (size_t)dim * sizeof(float);
}

extern "C" int mt_layernorm_bwd(int precision, const float* x, const float* gamma, const void* dy, float* gx,
                                float* dgamma_dbeta, int rows, int dim, void* workspace, size_t workspace_bytes,
                                void* stream) {
  MT_REQUIRE(x && gamma && dy && gx && dgamma_dbeta && rows > 0, "layernorm_bwd: bad argument");
  MT_REQUIRE(dim % 128 == 0 && dim <= 1024, "layernorm_bwd: dim must be a multiple of 128 <= 1024 (got %d)", dim);
  MT_REQUIRE(precision == MT_PREC_FP32 || precision == MT_PREC_BF16, "layernorm_bwd: unknown precision %d", precision);
  if (!workspace || workspace_bytes < mt_layernorm_bwd_workspace_bytes(rows, dim)) {
    set_error("layernorm_bwd: workspace too small");
    return MT_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* part = reinterpret_cast<float*>(workspace);
  ProfScope prof(st, 20.0 * rows * dim, (double)rows * dim * (12 + (precision == MT_PREC_FP32 ? 4 : 2)), "layernorm_bwd");
  int rc = precision == MT_PREC_FP32 ? launch_ln_bwd<float>(x, gamma, dy, gx, part, rows, dim, st)
                                     : launch_ln_bwd<bf16>(x, gamma, dy, gx, part, rows, dim, st);
  if (rc) return rc;
  return launch_colsum(part, dgamma_dbeta, ln_bwd_grid(rows), 2 * dim, 0, st);
}

extern "C" size_t mt_grad_prep_workspace_bytes(int m, int c) {
  return (size_t)((m + 63) / 64) * (size_t)c * sizeof(float);
}

extern "C" int mt_grad_prep(int precision, const void* src, int src_is_f32, void* out_rm, void* out_t, float* colsum,
                            int m, int c, int mp, int rows_per_batch, void* workspace, size_t workspace_bytes,
                            void* stream) {
  MT_REQUIRE(src && m > 0 && c > 0, "grad_prep: bad argument");
  MT_REQUIRE(c % 64 == 0, "grad_prep: the column count must be a multiple of 64 (got %d)", c);
  MT_REQUIRE(precision == MT_PREC_FP32 || precision == MT_PREC_BF16, "grad_prep: unknown precision %d", precision);
  MT_REQUIRE(!out_t || (mp % 64 == 0 && mp >= m), "grad_prep: mp (%d) must be a multiple of 64 >= m (%d)", mp, m);
  MT_REQUIRE(rows_per_batch >= 0, "grad_prep: rows_per_batch < 0");
  if (colsum && (!workspace || workspace_bytes < mt_grad_prep_workspace_bytes(m, c))) {
    set_error("grad_prep: workspace too small");
    return MT_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t es = precision == MT_PREC_FP32 ? 4 : 2;
  const int tiles_m = out_t ? mp / 64 : (m + 63) / 64;
  float* part = colsum ? reinterpret_cast<float*>(workspace) : nullptr;
  if (colsum && tiles_m != (m + 63) / 64) {
    // the zero-padded tail tiles of out_t would write partial rows beyond the workspace
    MT_REQUIRE(workspace_bytes >= (size_t)tiles_m * c * sizeof(float), "grad_prep: workspace too small for mp");
  }
  dim3 grid(c / 64, tiles_m);
  {
    ProfScope prof(st, 0.0, (double)m * c * ((src_is_f32 ? 4 : es) + (out_rm ? es : 0) + (out_t ? es : 0)), "grad_prep C%d%s%s%s", c,
                   src_is_f32 ? " f32" : "", out_rm ? " cast" : "", out_t ? " T" : "");
    const bool f32 = precision == MT_PREC_FP32;
    if (src_is_f32 || f32) {
      if (f32)
        grad_prep_kernel<float, float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(src), reinterpret_cast<float*>(out_rm),
                                                           reinterpret_cast<float*>(out_t), part, m, c, mp, rows_per_batch);
      else
        grad_prep_kernel<float, bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(src), reinterpret_cast<bf16*>(out_rm),
                                                          reinterpret_cast<bf16*>(out_t), part, m, c, mp, rows_per_batch);
    } else {
      grad_prep_kernel<bf16, bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(src), reinterpret_cast<bf16*>(out_rm),
                                                       reinterpret_cast<bf16*>(out_t), part, m, c, mp, rows_per_batch);
    }
    MT_LAUNCH_CHECK("grad_prep_kernel");
  }
  if (colsum) return launch_colsum(part, colsum, tiles_m, c, 0, st);
  return MT_OK;
}

extern "C" int mt_geglu_fwd(int precision, const void* h, void* out, int m, int n_out, void* stream) {
  MT_REQUIRE(h && out && m > 0 && n_out > 0 && n_out % 32 == 0, "geglu_fwd: bad argument (n_out %% 32)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total8 = (size_t)m * (n_out / 8);
  const unsigned grid = (unsigned)((total8 + 255) / 256);
  const size_t es = precision == MT_PREC_FP32 ? 4 : 2;
  ProfScope prof(st, 12.0 * m * n_out, (double)m * n_out * 3 * es, "geglu_fwd");
  if (precision == MT_PREC_FP32)
    geglu_fwd_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(h), reinterpret_cast<float*>(out), total8, n_out);
  else if (precision == MT_PREC_BF16)
    geglu_fwd_kernel<bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(h), reinterpret_cast<bf16*>(out), total8, n_out);
  else { set_error("geglu_fwd: unknown precision %d", precision); return MT_ERR_ARG; }
  MT_LAUNCH_CHECK("geglu_fwd_kernel");
  return MT_OK;
}

extern "C" int mt_geglu_bwd(int precision, const void* h, const void* dout, void* dh, int m, int n_out, void* stream) {
  MT_REQUIRE(h && dout && dh && m > 0 && n_out > 0 && n_out % 32 == 0, "geglu_bwd: bad argument (n_out %% 32)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t total8 = (size_t)m * (n_out / 8);
  const unsigned grid = (unsigned)((total8 + 255) / 256);
  const size_t es = precision == MT_PREC_FP32 ? 4 : 2;
  ProfScope prof(st, 24.0 * m * n_out, (double)m * n_out * 5 * es, "geglu_bwd");
  if (precision == MT_PREC_FP32)
    geglu_bwd_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(h), reinterpret_cast<const float*>(dout),
                                                  reinterpret_cast<float*>(dh), total8, n_out);
  else if (precision == MT_PREC_BF16)
    geglu_bwd_kernel<bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(h), reinterpret_cast<const bf16*>(dout),
                                                 reinterpret_cast<bf16*>(dh), total8, n_out);
  else { set_error("geglu_bwd: unknown precision %d", precision); return MT_ERR_ARG; }
  MT_LAUNCH_CHECK("geglu_bwd_kernel");
  return MT_OK;
}

static int geglu_colsum_grid(int m) { return std::min(m, 4 * sm_count()); }

extern "C" size_t mt_geglu_bwd_colsum_workspace_bytes(int m, int n_out) {
  if (m <= 0 || n_out <= 0) return 0;
  return (size_t)geglu_colsum_grid(m) * 2 * (size_t)n_out * sizeof(float);
}

extern "C" int mt_geglu_bwd_colsum(int precision, const void* h, const void* dout, void* dh, float* colsum, int m, int n_out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(h && dout && dh && colsum && m > 0 && n_out > 0 && n_out % 32 == 0, "geglu_bwd_colsum: bad argument (n_out %% 32)");
  if (!workspace || workspace_bytes < mt_geglu_bwd_colsum_workspace_bytes(m, n_out)) {
    set_error("geglu_bwd_colsum: workspace too small");
    return MT_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dim3 grid(geglu_colsum_grid(m), (n_out / 8 + 255) / 256);
  float* part = reinterpret_cast<float*>(workspace);
  const size_t es = precision == MT_PREC_FP32 ? 4 : 2;
  {
    ProfScope prof(st, 24.0 * m * n_out, (double)m * n_out * 5 * es, "geglu_bwd");
    if (precision == MT_PREC_FP32)
      geglu_bwd_colsum_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(h), reinterpret_cast<const float*>(dout),
                                                           reinterpret_cast<float*>(dh), part, m, n_out);
    else if (precision == MT_PREC_BF16)
      geglu_bwd_colsum_kernel<bf16><<<grid, 256, 0, st>>>(reinterpret_cast<const bf16*>(h), reinterpret_cast<const bf16*>(dout),
                                                          reinterpret_cast<bf16*>(dh), part, m, n_out);
    else { set_error("geglu_bwd_colsum: unknown precision %d", precision); return MT_ERR_ARG; }
    MT_LAUNCH_CHECK("geglu_bwd_colsum_kernel");
  }
  return launch_colsum(part, colsum, (int)grid.x, 2 * n_out, 0, st);
}

extern "C" size_t mt_divided_attn_bwd_workspace_bytes(int batch, int f, int n, int heads) {
  if (batch <= 0 || f <= 0 || n <= 0 || heads <= 0) return 0;
  return attn_bwd_ws(batch, f, n, heads).total;
}

extern "C" int mt_divided_attn_bwd(int precision, const void* qkv, const void* dout, const uint8_t* mask,
                                   const uint8_t* identities_mask, int mode, void* dqkv, int batch, int f, int n,
                                   int heads, int dim_head, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(qkv && dout && dqkv && mask && batch > 0, "divided_attn_bwd: bad argument");
  MT_REQUIRE(dim_head == 64, "divided_attn_bwd: dim_head must be 64 (got %d)", dim_head);
  MT_REQUIRE(mode == MT_ATTN_TIME || mode == MT_ATTN_SPACE, "divided_attn_bwd: bad mode %d", mode);
  MT_REQUIRE(mode != MT_ATTN_TIME || identities_mask, "divided_attn_bwd: time attention needs identities_mask");
  MT_REQUIRE(f >= 1 && f <= 63 && n >= 1 && n <= 63 && heads >= 1, "divided_attn_bwd: f, n must be in 1..63");
  MT_REQUIRE(precision == MT_PREC_FP32 || precision == MT_PREC_BF16, "divided_attn_bwd: unknown precision %d", precision);
  if (!workspace || workspace_bytes < mt_divided_attn_bwd_workspace_bytes(batch, f, n, heads) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15)) {
    set_error("divided_attn_bwd: workspace missing, too small or not 16-byte aligned");
    return MT_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  return precision == MT_PREC_FP32
             ? launch_attn_bwd<float>(qkv, dout, mask, identities_mask, mode, dqkv, batch, f, n, heads, ws, st)
             : launch_attn_bwd<bf16>(qkv, dout, mask, identities_mask, mode, dqkv, batch, f, n, heads, ws, st);
}

extern "C" int mt_embed_bwd(const float* g0, const int64_t* positions, const int32_t* size_embedding, float* dpos,
                            float* dsize, float* dcls, int batch, int f, int n, int dim, int table_rows, void* stream) {
  MT_REQUIRE(g0 && dcls && batch > 0 && f > 0 && n > 0 && dim > 0 && table_rows > f * n, "embed_bwd: bad argument");
  MT_REQUIRE(!dsize || size_embedding, "embed_bwd: dsize needs size_embedding");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ProfScope prof(st, 0.0, (double)batch * (1 + f * n) * dim * 12.0, "embed_bwd");
  embed_bwd_kernel<<<batch * (f + 1), 128, 0, st>>>(g0, reinterpret_cast<const long long*>(positions), size_embedding, dpos,
                                                    dsize, dcls, f, n, dim, table_rows);
  MT_LAUNCH_CHECK("embed_bwd_kernel");
  return MT_OK;
}

extern "C" size_t mt_head_bwd_workspace_bytes(int batch, int dim, int num_classes) {
  return (size_t)batch * ((size_t)num_classes * dim + num_classes + 2 * (size_t)dim) * sizeof(float);
}

extern "C" int mt_head_bwd(const float* x, const float* ln_g, const float* ln_b, const float* w, const float* dlogits,
                           float* gx, float* grads, int batch, int tokens, int dim, int num_classes, void* workspace,
                           size_t workspace_bytes, void* stream) {
  MT_REQUIRE(x && ln_g && ln_b && w && dlogits && gx && grads && batch > 0 && dim > 0 && num_classes > 0,
             "head_bwd: bad argument");
  if (!workspace || workspace_bytes < mt_head_bwd_workspace_bytes(batch, dim, num_classes)) {
    set_error("head_bwd: workspace too small");
    return MT_ERR_WORKSPACE;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* part = reinterpret_cast<float*>(workspace);
  head_bwd_kernel<<<batch, 128, 2 * (size_t)dim * sizeof(float), st>>>(x, ln_g, ln_b, w, dlogits, gx, part, tokens, dim,
                                                                      num_classes);
  MT_LAUNCH_CHECK("head_bwd_kernel");
  return launch_colsum(part, grads, batch, num_classes * dim + num_classes + 2 * dim, 0, st);
}
