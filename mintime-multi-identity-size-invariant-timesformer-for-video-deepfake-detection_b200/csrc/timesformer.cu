// Size-Invariant TimeSformer forward (reference models/size_invariant_timesformer.py:224-276).
// Linear layers run in gemm.cu; this file holds LayerNorm, the divided (time / space) attention core
// with the identity mask, the CLS-row attention, token assembly for the CLS row, the head, and the
// layer schedule.  Residual stream x is fp32 [B][1+f*n][dim]; GEMM operands are T.
#include <float.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>

#include "attention_mma.cuh"
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

// ---------------------------------------------------------------------------------------------------
// LayerNorm(dim, eps 1e-5) f32 -> T, one warp per row  (PreNorm, :18-26)
// ---------------------------------------------------------------------------------------------------
// Persistent form: a warp walks rows with a grid stride and has the NEXT row's loads in flight while it reduces
// and writes the current one (the one-row-per-warp version ran at 43 % occupancy and 3 TB/s: blocks lived ~1 us).
template <typename T, int NV>   // NV = dim / 128 float4 per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ b, T* __restrict__ out,
                                                        float* __restrict__ x_copy, int rows, int dim) {
  const int lane = threadIdx.x & 31;
  const int wstride = gridDim.x * 8;
  int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4 gg[NV], bv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gg[i] = *reinterpret_cast<const float4*>(g + (i * 32 + lane) * 4);
    bv[i] = *reinterpret_cast<const float4*>(b + (i * 32 + lane) * 4);
  }
  float4 v[NV], nx[NV];
  {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * dim);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = xr[i * 32 + lane];
  }
  for (; row < rows; row += wstride) {
    const int nrow = row + wstride;
    if (nrow < rows) {
      const float4* xr = reinterpret_cast<const float4*>(x + (size_t)nrow * dim);
#pragma unroll
      for (int i = 0; i < NV; ++i) nx[i] = xr[i * 32 + lane];
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / (float)dim;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + bb * bb) + (c * c + d * d);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)dim + 1e-5f);
    T* orow = out + (size_t)row * dim;
    if (x_copy != nullptr) {          // training: the sub-block's input survives the in-place residual update that follows
      float4* cr = reinterpret_cast<float4*>(x_copy + (size_t)row * dim);
#pragma unroll
      for (int i = 0; i < NV; ++i) cr[i * 32 + lane] = v[i];
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c0 = (i * 32 + lane) * 4;
      const float o0 = (v[i].x - mean) * rstd * gg[i].x + bv[i].x, o1 = (v[i].y - mean) * rstd * gg[i].y + bv[i].y;
      const float o2 = (v[i].z - mean) * rstd * gg[i].z + bv[i].z, o3 = (v[i].w - mean) * rstd * gg[i].w + bv[i].w;
      if (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(orow) + c0) = make_float4(o0, o1, o2, o3);
      } else {
        uint2 u;
        *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(o0, o1);
        *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(o2, o3);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(orow) + c0) = u;
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = nx[i];
  }
}

// ---------------------------------------------------------------------------------------------------
// CLS row (:117-120): query 0 of each (b,h) attends all N keys; keys of padded frames are masked
// (cls_attn_mask :258-260).  Probabilities are the attention map the model returns (:271).
// grid = B*heads, block = 256, dim_head = 64.
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) attn_cls_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ mask,
                                                       T* __restrict__ out, float* __restrict__ cls_attn, int N, int f,
                                                       int n, int heads) {
  extern __shared__ float sm[];
  float* sc = sm;              // [N]
  float* q0 = sm + N;          // [64]
  float* red = q0 + 64;        // [32]
  float* part = red + 32;      // [32][64] partial outputs
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int inner = heads * 64, ld = 3 * inner;
  const T* base = qkv + (size_t)b * N * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 64) q0[tid] = to_f(base[h * 64 + tid]);
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
    if (j > 0 && !mask[b * f + (j - 1) / n]) s = -FLT_MAX;   // masked_fill(~mask, -finfo.max) (:83-84)
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
  for (int j = tid; j < N; j += 256) {
    const float pj = sc[j] * inv;
    sc[j] = pj;
    if (cls_attn) cls_attn[(size_t)blockIdx.x * N + j] = pj;
  }
  __syncthreads();
  // o = P V: thread = (key group kg of 32, 8-dim chunk dc); 16-byte (8-dim) loads of the V rows, ~25
  // independent loads per thread, then a fixed-order reduction over the 32 key groups in shared memory
  const int dc = tid & 7, kg = tid >> 3;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int j = kg; j < N; j += 32) {
    float vv[8];
    load8(base + (size_t)j * ld + 2 * inner + h * 64 + dc * 8, vv);
    const float pj = sc[j];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pj, vv[i], acc[i]);
  }
  __syncthreads();
  float* partial = part;                 // [32][64] floats
#pragma unroll
  for (int i = 0; i < 8; ++i) partial[kg * 64 + dc * 8 + i] = acc[i];
  __syncthreads();
  if (tid < 64) {
    float r = 0.f;
#pragma unroll 8
    for (int g = 0; g < 32; ++g) r += partial[g * 64 + tid];
    out[(size_t)b * N * inner + h * 64 + tid] = from_f<T>(r);
  }
}

// ---------------------------------------------------------------------------------------------------
// Patch rows (:122-135).  One block per group:
//   TIME  group (b,h,patch p): queries = the f frames at patch p, keys = CLS + those f tokens,
//         key k allowed iff mask[b][k] & identities_mask[b][q][k]  (frame_mask :252-255); CLS always.
//   SPACE group (b,h,frame fr): queries = the n patches of the frame, keys = CLS + those n tokens, no mask.
// K/V of the group sit in shared memory as fp32; each warp owns a query at a time.
// ---------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(128) attn_group_kernel(const T* __restrict__ qkv, const uint8_t* __restrict__ mask,
                                                         const uint8_t* __restrict__ idmask, T* __restrict__ out,
                                                         int f, int n, int heads) {
  __shared__ float Ks[64][65];
  __shared__ float Vs[64][65];
  __shared__ float Qs[4][64];
  __shared__ float Ps[4][64];
  const int G = MODE == MT_ATTN_TIME ? n : f;
  const int Gq = MODE == MT_ATTN_TIME ? f : n;
  const int Gk = Gq + 1;
  const int g = blockIdx.x % G;
  const int h = (blockIdx.x / G) % heads;
  const int b = blockIdx.x / (G * heads);
  const int N = 1 + f * n, inner = heads * 64, ld = 3 * inner;
  const T* base = qkv + (size_t)b * N * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto token = [&](int j) -> int {  // j = 0 is CLS, j >= 1 the (j-1)-th member of the group
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
    for (int i = 0; i < 8; ++i) { Ks[j][c * 8 + i] = kv[i]; Vs[j][c * 8 + i] = vv[i]; }
  }
  __syncthreads();
  for (int i = warp; i < Gq; i += 4) {
    const int tq = token(i + 1);
    const T* qr = base + (size_t)tq * ld + h * 64;
    Qs[warp][lane] = to_f(qr[lane]);
    Qs[warp][lane + 32] = to_f(qr[lane + 32]);
    __syncwarp();
    float s[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int j = lane + r * 32;
      float a = -FLT_MAX;
      if (j < Gk) {
        a = 0.f;
#pragma unroll 16
        for (int d = 0; d < 64; ++d) a = fmaf(Qs[warp][d], Ks[j][d], a);
        if (MODE == MT_ATTN_TIME && j > 0) {
          const bool ok = mask[b * f + (j - 1)] && idmask[((size_t)b * f + i) * f + (j - 1)];
          if (!ok) a = -FLT_MAX;
        }
      }
      s[r] = a;
    }
    const float mx = warp_max(fmaxf(s[0], s[1]));
    const float e0 = (lane < Gk) ? expf(s[0] - mx) : 0.f;
    const float e1 = (lane + 32 < Gk) ? expf(s[1] - mx) : 0.f;
    const float inv = 1.0f / warp_sum(e0 + e1);
    Ps[warp][lane] = e0 * inv;
    Ps[warp][lane + 32] = e1 * inv;
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Gk; ++j) {
      const float pj = Ps[warp][j];
      o0 = fmaf(pj, Vs[j][lane], o0);
      o1 = fmaf(pj, Vs[j][lane + 32], o1);
    }
    T* orow = out + ((size_t)b * N + tq) * inner + h * 64;
    orow[lane] = from_f<T>(o0);
    orow[lane + 32] = from_f<T>(o1);
    __syncwarp();
  }
}

// x[b][0] = cls_token + pos_emb[positions[b][0]] + size_emb[0]   (:231-248)
__global__ void cls_row_kernel(const float* __restrict__ cls, const float* __restrict__ pos_tab,
                               const float* __restrict__ size_tab, const long long* __restrict__ positions, float* x,
                               int tokens, int dim, int table_rows) {
  const int b = blockIdx.x;
  const long long p = positions ? positions[(size_t)b * tokens] : 0;
  if ((unsigned long long)p >= (unsigned long long)table_rows) __trap();     // nn.Embedding would raise IndexError
  for (int c = threadIdx.x; c < dim; c += blockDim.x) {
    float v = cls[c] + pos_tab[(size_t)p * dim + c];
    if (size_tab) v += size_tab[c];
    x[(size_t)b * tokens * dim + c] = v;
  }
}

// logits[b] = LayerNorm(x[b][0]) W^T + bias  (:195-198, :270-276); one block (128 threads) per video
__global__ void __launch_bounds__(128) head_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                   const float* __restrict__ be, const float* __restrict__ w,
                                                   const float* __restrict__ bias, float* __restrict__ logits,
                                                   int tokens, int dim, int classes) {
  extern __shared__ float sm[];
  float* xn = sm;          // [dim]
  __shared__ float red[4];
  const float* xr = x + (size_t)blockIdx.x * tokens * dim;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float s = 0.f;
  for (int c = tid; c < dim; c += 128) s += xr[c];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  const float mean = ((red[0] + red[1]) + (red[2] + red[3])) / (float)dim;
  __syncthreads();
  float q = 0.f;
  for (int c = tid; c < dim; c += 128) { const float d = xr[c] - mean; q += d * d; }
  q = warp_sum(q);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  const float rstd = 1.0f / sqrtf(((red[0] + red[1]) + (red[2] + red[3])) / (float)dim + 1e-5f);
  for (int c = tid; c < dim; c += 128) xn[c] = (xr[c] - mean) * rstd * g[c] + be[c];
  __syncthreads();
  for (int k = warp; k < classes; k += 4) {
    float a = 0.f;
    for (int c = lane; c < dim; c += 32) a = fmaf(xn[c], w[(size_t)k * dim + c], a);
    a = warp_sum(a);
    if (lane == 0) logits[(size_t)blockIdx.x * classes + k] = a + bias[k];
  }
}

template <typename T>
int launch_attn_t(const void* qkv, const uint8_t* mask, const uint8_t* idmask, int mode, void* out, float* cls_attn,
                  int B, int f, int n, int heads, float* cls_parts, cudaStream_t st) {
  const int N = 1 + f * n;
  const T* q = reinterpret_cast<const T*>(qkv);
  T* o = reinterpret_cast<T*>(out);
  const size_t smem = (size_t)(N + 64 + 32 + 2048) * sizeof(float);
  // bf16 path with a tensor-core grouped kernel: the CLS row is computed inside it (partials) + cls_combine_kernel
  const bool fused_cls = std::is_same<T, bf16>::value && cls_parts != nullptr && (mode == MT_ATTN_SPACE || f <= 32);
  if (!fused_cls) {
    ProfScope prof(st, 4.0 * B * heads * 64.0 * N, (double)B * N * heads * 64 * 2 * sizeof(T), "attn_cls");
    attn_cls_kernel<T><<<B * heads, 256, smem, st>>>(q, mask, o, cls_attn, N, f, n, heads);
    MT_LAUNCH_CHECK("attn_cls_kernel");
  }
  const double gq = mode == MT_ATTN_TIME ? f : n, groups = (double)B * heads * (mode == MT_ATTN_TIME ? n : f);
  ProfScope prof(st, 4.0 * groups * 64.0 * gq * (gq + 1), (double)B * N * heads * 64 * 4 * sizeof(T),
                 mode == MT_ATTN_TIME ? "attn_time" : "attn_space");
  if constexpr (std::is_same<T, bf16>::value) {
    // bf16 path: warp-level tensor-core kernels (attention_mma.cuh)
    auto combine = [&]() -> int {
      attn::cls_combine_kernel<<<B * heads, 64, 0, st>>>(q, N, cls_parts, o, cls_attn, N, mode == MT_ATTN_SPACE ? f : n, heads);
      MT_LAUNCH_CHECK("cls_combine_kernel");
      return MT_OK;
    };
    if (fused_cls && mode == MT_ATTN_SPACE) {
      attn::attn_space_mma_kernel<<<B * heads * f, 160, 0, st>>>(q, mask, o, cls_parts, cls_attn, f, n, heads);
      MT_LAUNCH_CHECK("attn_space_mma_kernel");
      return combine();
    }
    const int grid = B * heads * ((n + 3) / 4);
    auto launch_time = [&](auto kern, int nkt, int mt_) -> int {
      const size_t dyn = 4 * (size_t)(mt_ * 16 * 128 + 2 * nkt * 16 * 128);
      if (dyn > 40 * 1024) {                     // (the kernel also has ~3 KiB of static shared memory)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(attn_time)");
      }
      kern<<<grid, 256, dyn, st>>>(q, mask, idmask, o, cls_parts, cls_attn, f, n, heads);
      MT_LAUNCH_CHECK("attn_time_mma_kernel");
      return combine();
    };
    if (fused_cls && mode == MT_ATTN_TIME) {
      if (f <= 15) return launch_time(attn::attn_time_mma_kernel<1, 1>, 1, 1);
      if (f == 16) return launch_time(attn::attn_time_mma_kernel<2, 1>, 2, 1);
      if (f <= 31) return launch_time(attn::attn_time_mma_kernel<2, 2>, 2, 2);
      return launch_time(attn::attn_time_mma_kernel<3, 2>, 3, 2);
    }
    // (no workspace / other frame counts: generic kernel below, CLS row already done by attn_cls_kernel)
  }
  if (mode == MT_ATTN_TIME)
    attn_group_kernel<T, MT_ATTN_TIME><<<B * heads * n, 128, 0, st>>>(q, mask, idmask, o, f, n, heads);
  else
    attn_group_kernel<T, MT_ATTN_SPACE><<<B * heads * f, 128, 0, st>>>(q, mask, idmask, o, f, n, heads);
  MT_LAUNCH_CHECK("attn_group_kernel");
  return MT_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t attn_ws_bytes(int B, int f, int n, int heads) {
  return (size_t)B * heads * (size_t)std::max(f, n) * attn::kClsStride * sizeof(float);
}

struct TsfWs { size_t x, xn, qkv, o, h, cls, total; };   // cls: CLS-row partials (+ the fused kernel's CLS q/k/v copy)
TsfWs tsf_ws_layout(const mt_tsf_cfg_t& c, int B, int precision) {
  const size_t es = precision == MT_PREC_FP32 ? 4 : 2;
  const size_t rows = (size_t)B * (1 + (size_t)c.num_frames * c.num_patches);
  const size_t inner = (size_t)c.heads * c.dim_head;
  TsfWs l;
  size_t off = 0;
  l.x = off;   off += align_up(rows * c.dim * 4, 1024);
  l.xn = off;  off += align_up(rows * c.dim * es, 1024);
  l.qkv = off; off += align_up(rows * 3 * inner * es, 1024);
  l.o = off;   off += align_up(rows * inner * es, 1024);
  l.h = off;   off += align_up(rows * 4 * c.dim * es, 1024);
  l.cls = off; off += align_up(std::max(attn_ws_bytes(B, c.num_frames, c.num_patches, c.heads),
                                        mt_fused_attn_workspace_bytes(B, c.num_frames, c.num_patches, c.heads)), 1024);
  l.total = off;
  return l;
}

// MINTIME_B200_FUSED_ATTN=0 (read once) keeps the unfused projection + attention kernels: A/B measurements
bool fused_attn_enabled() {
  static const bool on = [] {
    const char* e = getenv("MINTIME_B200_FUSED_ATTN");
    return !(e && e[0] == '0');
  }();
  return on;
}

int check_cfg(const mt_tsf_cfg_t* c) {
  MT_REQUIRE(c, "tsf: null config");
  MT_REQUIRE(c->dim_head == 64, "tsf: dim_head must be 64 (got %d)", c->dim_head);
  MT_REQUIRE(c->dim % 128 == 0 && c->dim <= 1024, "tsf: dim must be a multiple of 128 <= 1024 (got %d)", c->dim);
  MT_REQUIRE(c->depth >= 1 && c->depth <= MT_TSF_MAX_DEPTH, "tsf: depth out of range (%d)", c->depth);
  MT_REQUIRE(c->num_frames >= 1 && c->num_frames <= 63 && c->num_patches >= 1 && c->num_patches <= 63,
             "tsf: frames/patches per group must be <= 63 (f=%d n=%d)", c->num_frames, c->num_patches);
  MT_REQUIRE(c->heads >= 1 && c->channels % 8 == 0 && c->num_classes >= 1, "tsf: bad heads/channels/classes");
  return MT_OK;
}

}  // namespace
}  // namespace mt

using namespace mt;

extern "C" int mt_layernorm_fwd(int precision, const float* x, const float* gamma, const float* beta, void* out,
                                int rows, int dim, void* stream) {
  return mt_layernorm_copy_fwd(precision, x, gamma, beta, out, nullptr, rows, dim, stream);
}

extern "C" int mt_layernorm_copy_fwd(int precision, const float* x, const float* gamma, const float* beta, void* out,
                                     float* x_copy, int rows, int dim, void* stream) {
  MT_REQUIRE(x && gamma && beta && out && rows > 0, "layernorm: bad argument");
  MT_REQUIRE(dim % 128 == 0 && dim <= 1024, "layernorm: dim must be a multiple of 128 <= 1024 (got %d)", dim);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min((rows + 7) / 8, sms * 6);      // persistent: <= 6 blocks (48 warps) per SM
  ProfScope prof(st, 8.0 * rows * dim, (double)rows * dim * (4 + (x_copy ? 4 : 0) + (precision == MT_PREC_FP32 ? 4 : 2)), "layernorm");
  const int nv = dim >> 7;
#define MT_LN_LAUNCH(TT, NVV) layernorm_kernel<TT, NVV><<<grid, 256, 0, st>>>(x, gamma, beta, reinterpret_cast<TT*>(out), x_copy, rows, dim)
  if (precision != MT_PREC_FP32 && precision != MT_PREC_BF16) {
    set_error("layernorm: unknown precision %d", precision);
    return MT_ERR_ARG;
  }
  const bool f32 = precision == MT_PREC_FP32;
  switch (nv) {
    case 1: if (f32) MT_LN_LAUNCH(float, 1); else MT_LN_LAUNCH(bf16, 1); break;
    case 2: if (f32) MT_LN_LAUNCH(float, 2); else MT_LN_LAUNCH(bf16, 2); break;
    case 3: if (f32) MT_LN_LAUNCH(float, 3); else MT_LN_LAUNCH(bf16, 3); break;
    case 4: if (f32) MT_LN_LAUNCH(float, 4); else MT_LN_LAUNCH(bf16, 4); break;
    case 5: if (f32) MT_LN_LAUNCH(float, 5); else MT_LN_LAUNCH(bf16, 5); break;
    case 6: if (f32) MT_LN_LAUNCH(float, 6); else MT_LN_LAUNCH(bf16, 6); break;
    case 7: if (f32) MT_LN_LAUNCH(float, 7); else MT_LN_LAUNCH(bf16, 7); break;
    default: if (f32) MT_LN_LAUNCH(float, 8); else MT_LN_LAUNCH(bf16, 8); break;
  }
#undef MT_LN_LAUNCH
  MT_LAUNCH_CHECK("layernorm_kernel");
  return MT_OK;
}

extern "C" size_t mt_divided_attn_workspace_bytes(int batch, int f, int n, int heads) {
  if (batch <= 0 || f <= 0 || n <= 0 || heads <= 0) return 0;
  return attn_ws_bytes(batch, f, n, heads);
}

extern "C" int mt_divided_attn_fwd(int precision, const void* qkv, const uint8_t* mask, const uint8_t* identities_mask,
                                   int mode, void* out, float* cls_attn, int batch, int f, int n, int heads,
                                   int dim_head, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(qkv && mask && out, "divided_attn: null pointer");
  MT_REQUIRE(mode == MT_ATTN_TIME || mode == MT_ATTN_SPACE, "divided_attn: unknown mode %d", mode);
  MT_REQUIRE(mode == MT_ATTN_SPACE || identities_mask, "divided_attn: time mode needs identities_mask");
  MT_REQUIRE(dim_head == 64, "divided_attn: dim_head must be 64 (got %d)", dim_head);
  MT_REQUIRE(batch > 0 && heads > 0 && f >= 1 && f <= 63 && n >= 1 && n <= 63, "divided_attn: bad shape B=%d f=%d n=%d",
             batch, f, n);
  MT_REQUIRE((size_t)(1 + f * n + 2144) * 4 <= 48 * 1024, "divided_attn: too many tokens (%d)", 1 + f * n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (workspace && workspace_bytes < attn_ws_bytes(batch, f, n, heads)) {
    set_error("divided_attn: workspace too small (%zu < %zu)", workspace_bytes, attn_ws_bytes(batch, f, n, heads));
    return MT_ERR_WORKSPACE;
  }
  MT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "divided_attn: workspace must be 16-byte aligned");
  float* parts = reinterpret_cast<float*>(workspace);   // null: CLS row by the stand-alone kernel
  if (precision == MT_PREC_FP32)
    return launch_attn_t<float>(qkv, mask, identities_mask, mode, out, cls_attn, batch, f, n, heads, parts, st);
  if (precision == MT_PREC_BF16)
    return launch_attn_t<bf16>(qkv, mask, identities_mask, mode, out, cls_attn, batch, f, n, heads, parts, st);
  set_error("divided_attn: unknown precision %d", precision);
  return MT_ERR_ARG;
}

extern "C" int mt_linear_residual_fwd(int precision, const void* a, const void* w, const float* bias, float* x, int m,
                                      int n, int k, void* stream) {
  GemmArgs g{};
  g.a = a; g.w = w; g.M = m; g.N = n; g.K = k;
  g.epi.kind = EPI_RESID_F32; g.epi.M = m; g.epi.N = n; g.epi.bias = bias; g.epi.out = x; g.epi.ldo = n;
  return launch_gemm(precision, g, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mt_linear_geglu_fwd(int precision, const void* a, const void* w, const float* bias, void* out, int m,
                                   int n, int k, void* stream) {
  GemmArgs g{};
  g.a = a; g.w = w; g.M = m; g.N = n; g.K = k;
  g.epi.kind = EPI_GEGLU; g.epi.M = m; g.epi.N = n; g.epi.bias = bias; g.epi.out = out; g.epi.ldo = n / 2;
  return launch_gemm(precision, g, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mt_patch_embed_fwd(int precision, const mt_tsf_weights_t* w, const mt_tsf_cfg_t* cfg, const void* feats,
                                  const int32_t* size_embedding, const int64_t* positions, float* x, int batch,
                                  void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  MT_REQUIRE(w && feats && x && batch > 0, "patch_embed: bad argument");
  MT_REQUIRE(!cfg->enable_pos_emb || positions, "patch_embed: positions required when enable-pos-emb is on");
  MT_REQUIRE(!cfg->enable_size_emb || (size_embedding && w->size_emb), "patch_embed: size embedding inputs missing");
  const int fn = cfg->num_frames * cfg->num_patches;
  const long long* pos = cfg->enable_pos_emb ? reinterpret_cast<const long long*>(positions) : nullptr;
  const float* size_tab = cfg->enable_size_emb ? w->size_emb : nullptr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int table_rows = cfg->num_frames * cfg->channels + 1;      // num_positions + 1 (:173-180)
  MT_REQUIRE(fn + 1 <= table_rows, "patch_embed: %d tokens exceed the %d rows of the embedding tables", fn + 1, table_rows);
  cls_row_kernel<<<batch, 128, 0, st>>>(w->cls_token, w->pos_emb, size_tab, pos, x, fn + 1, cfg->dim, table_rows);
  MT_LAUNCH_CHECK("cls_row_kernel");
  GemmArgs g{};
  g.a = feats; g.w = w->w_patch; g.M = batch * fn; g.N = cfg->dim; g.K = cfg->channels;
  g.epi.kind = EPI_PATCH_EMBED; g.epi.M = g.M; g.epi.N = g.N; g.epi.bias = w->b_patch; g.epi.out = x;
  g.epi.ldo = cfg->dim; g.epi.rows_per_batch = fn; g.epi.n_patches = cfg->num_patches; g.epi.frames = cfg->num_frames;
  g.epi.pos_tab = w->pos_emb; g.epi.size_tab = size_tab; g.epi.positions = pos; g.epi.size_idx = size_embedding;
  g.epi.table_rows = table_rows;
  return launch_gemm(precision, g, st);
}

extern "C" int mt_head_fwd(const float* x, const float* ln_g, const float* ln_b, const float* w, const float* bias,
                           float* logits, int batch, int tokens, int dim, int num_classes, void* stream) {
  MT_REQUIRE(x && ln_g && ln_b && w && bias && logits && batch > 0 && tokens > 0 && dim > 0 && num_classes > 0,
             "head: bad argument");
  head_kernel<<<batch, 128, (size_t)dim * 4, reinterpret_cast<cudaStream_t>(stream)>>>(x, ln_g, ln_b, w, bias, logits,
                                                                                       tokens, dim, num_classes);
  MT_LAUNCH_CHECK("head_kernel");
  return MT_OK;
}

extern "C" size_t mt_tsf_workspace_bytes(const mt_tsf_cfg_t* cfg, int batch, int precision) {
  if (!cfg || batch <= 0) return 0;
  return tsf_ws_layout(*cfg, batch, precision).total;
}

extern "C" int mt_tsf_fwd(const mt_tsf_weights_t* w, const mt_tsf_cfg_t* cfg, const void* feats, const uint8_t* mask,
                          const uint8_t* identities_mask, const int32_t* size_embedding, const int64_t* positions,
                          float* logits, float* space_attn, float* time_attn, int batch, int precision, void* workspace,
                          size_t workspace_bytes, void* stream) {
  int rc = check_cfg(cfg);
  if (rc) return rc;
  MT_REQUIRE(w && feats && mask && identities_mask && logits && workspace && batch > 0, "tsf_fwd: bad argument");
  MT_REQUIRE(precision == MT_PREC_FP32 || precision == MT_PREC_BF16, "tsf_fwd: unknown precision %d", precision);
  const TsfWs l = tsf_ws_layout(*cfg, batch, precision);
  if (workspace_bytes < l.total) {
    set_error("tsf_fwd: workspace too small (%zu < %zu)", workspace_bytes, l.total);
    return MT_ERR_WORKSPACE;
  }
  MT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "tsf_fwd: workspace must be 1024-byte aligned");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* x = reinterpret_cast<float*>(ws + l.x);
  void* xn = ws + l.xn;
  void* qkv = ws + l.qkv;
  void* o = ws + l.o;
  void* hbuf = ws + l.h;
  const int f = cfg->num_frames, n = cfg->num_patches, D = cfg->dim, H = cfg->heads;
  const int N = 1 + f * n, rows = batch * N, inner = H * cfg->dim_head;

  rc = mt_patch_embed_fwd(precision, w, cfg, feats, size_embedding, positions, x, batch, stream);
  if (rc) return rc;
  for (int l_ = 0; l_ < cfg->depth; ++l_) {
    const bool last = l_ == cfg->depth - 1;
    for (int mode = 0; mode < 2; ++mode) {   // time, then space (:264-267)
      const mt_attn_weights_t& aw = mode == 0 ? w->time_attn[l_] : w->space_attn[l_];
      float* amap = !last ? nullptr : (mode == 0 ? time_attn : space_attn);
      rc = mt_layernorm_fwd(precision, x, aw.ln_g, aw.ln_b, xn, rows, D, stream);
      if (rc) return rc;
      if (precision == MT_PREC_BF16 && aw.w_qkv_heads && fused_attn_enabled() &&
          mt_fused_attn_supported(f, n, H, cfg->dim_head, D)) {
        // projection + attention core in one kernel: qkv stays on chip (attention_fused.cuh)
        rc = mt_fused_attn_fwd(xn, aw.w_qkv_heads, mask, identities_mask, mode == 0 ? MT_ATTN_TIME : MT_ATTN_SPACE, o, amap,
                               batch, f, n, H, cfg->dim_head, D, ws + l.cls, l.total - l.cls, stream);
        if (rc) return rc;
      } else {
        rc = mt_pointwise_fwd(precision, xn, aw.w_qkv, nullptr, nullptr, 0, nullptr, 0, qkv, rows, 3 * inner, D, stream);
        if (rc) return rc;
        rc = mt_divided_attn_fwd(precision, qkv, mask, identities_mask, mode == 0 ? MT_ATTN_TIME : MT_ATTN_SPACE, o, amap,
                                 batch, f, n, H, cfg->dim_head, ws + l.cls, l.total - l.cls, stream);
        if (rc) return rc;
      }
      rc = mt_linear_residual_fwd(precision, o, aw.w_out, aw.b_out, x, rows, D, inner, stream);
      if (rc) return rc;
    }
    const mt_ff_weights_t& fw = w->ff[l_];
    rc = mt_layernorm_fwd(precision, x, fw.ln_g, fw.ln_b, xn, rows, D, stream);
    if (rc) return rc;
    rc = mt_linear_geglu_fwd(precision, xn, fw.w1, fw.b1, hbuf, rows, 8 * D, D, stream);
    if (rc) return rc;
    rc = mt_linear_residual_fwd(precision, hbuf, fw.w2, fw.b2, x, rows, D, 4 * D, stream);
    if (rc) return rc;
  }
  return mt_head_fwd(x, w->out_ln_g, w->out_ln_b, w->out_w, w->out_b, logits, batch, N, D, cfg->num_classes, stream);
}

// ---------------------------------------------------------------------------------------------------
// Attention aggregation of predict.py / test.py (reference utils.py:68-96), per video on the device:
//   tok[i]  = max over heads of the CLS attention of token i            (utils.py:75, written for b = 1)
//   chunks  = np.array_split(tok, f): the first N % f chunks hold N/f + 1 tokens (chunk 0 = CLS + frame 0)
//   out[k]  = softmax_k( mean(chunk_k) * scale )                         (utils.py:84-86)
// for space, time and their elementwise sum -> out [B][3][f].  One block per video.
// ---------------------------------------------------------------------------------------------------
namespace mt {
namespace {
__global__ void __launch_bounds__(256) aggregate_attn_kernel(const float* __restrict__ space, const float* __restrict__ time,
                                                             float* __restrict__ out, int heads, int f, int N, float scale) {
  extern __shared__ float sm[];
  float* tok = sm;              // [3][N]
  float* fm = sm + 3 * N;       // [3][f]
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < N; i += 256) {
    float ms = -FLT_MAX, mt_ = -FLT_MAX;
    for (int h = 0; h < heads; ++h) {
      ms = fmaxf(ms, space[((size_t)b * heads + h) * N + i]);
      mt_ = fmaxf(mt_, time[((size_t)b * heads + h) * N + i]);
    }
    tok[i] = ms; tok[N + i] = mt_; tok[2 * N + i] = ms + mt_;
  }
  __syncthreads();
  const int base = N / f, rem = N % f;
  for (int e = tid; e < 3 * f; e += 256) {
    const int which = e / f, k = e % f;
    const int start = k * base + min(k, rem), len = base + (k < rem ? 1 : 0);
    float s = 0.f;
    for (int i = 0; i < len; ++i) s += tok[which * N + start + i];
    fm[e] = s / (float)len * scale;
  }
  __syncthreads();
  if (tid < 3) {
    float mx = -FLT_MAX;
    for (int k = 0; k < f; ++k) mx = fmaxf(mx, fm[tid * f + k]);
    float den = 0.f;
    for (int k = 0; k < f; ++k) den += expf(fm[tid * f + k] - mx);
    for (int k = 0; k < f; ++k) out[((size_t)b * 3 + tid) * f + k] = expf(fm[tid * f + k] - mx) / den;
  }
}
}  // namespace
}  // namespace mt

extern "C" int mt_aggregate_attn_fwd(const float* space_attn, const float* time_attn, float* out, int batch, int heads,
                                     int num_frames, int tokens, float scale, void* stream) {
  MT_REQUIRE(space_attn && time_attn && out && batch > 0 && heads > 0 && num_frames > 0 && tokens >= num_frames,
             "aggregate_attn: bad argument");
  const size_t smem = (size_t)(3 * tokens + 3 * num_frames) * sizeof(float);
  MT_REQUIRE(smem <= 48 * 1024, "aggregate_attn: too many tokens (%d)", tokens);
  mt::aggregate_attn_kernel<<<batch, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(space_attn, time_attn, out, heads,
                                                                                         num_frames, tokens, scale);
  MT_LAUNCH_CHECK("aggregate_attn_kernel");
  return MT_OK;
}

// ---------------------------------------------------------------------------------------------------
// Clip metadata on the device: the mask / identities_mask / size-embedding / temporal
// position tensors that DeepFakesDataset.__getitem__ (deepfakes_dataset.py:259-330) and predict.py's
// generate_masks (:254-352) assemble on the host for every clip, from the per-identity slot table:
//   identity i owns slots[i] consecutive face slots, the first n_real[i] hold faces, the rest are padding
//   size_embedding = bucket(ratio) in 1..20 for faces (SIZE_EMB_DICT, :30-31 / :259-263: (0..5)->1, (6..10)->2, ...),
//                    0 for padding (:273-274)
//   mask           = mask_padding ? (1 for faces, 0 for padding: predict.py:303) : all ones.  As EXECUTED,
//                    DeepFakesDataset never masks: its second `len(identity_images) < max_faces` test (:283) sees the
//                    list it has just padded to max_faces, so it always takes the all-ones branch (:286) -- pinned by
//                    tests/golden/clip_meta_ref.json, generated by running the reference class.
//   padded slots repeat the largest frame number appended SO FAR, over all identities up to this one
//                    (`max(images_frames)` on the clip-wide list, :277 / predict.py:304); 0 when there is none yet
//   identities_mask[q][k] = q and k belong to the same identity (:314-321)
//   positions      = [0] + for each slot the 49 token positions (p-1)*n+1 .. p*n of p = 1-based rank of the slot's
//                    frame number among the clip's distinct frame numbers (:323-329)
// One block per clip, one thread per slot.
// ---------------------------------------------------------------------------------------------------
namespace mt {
namespace {
__global__ void __launch_bounds__(64) clip_meta_kernel(const int* __restrict__ slots, const int* __restrict__ n_real,
                                                       const int* __restrict__ frame_no, const int* __restrict__ ratio,
                                                       int max_ids, int mask_padding, uint8_t* __restrict__ mask,
                                                       uint8_t* __restrict__ idmask, int* __restrict__ size_emb,
                                                       long long* __restrict__ positions, int f, int n) {
  __shared__ int ident[64], frames[64], first[64];
  const int b = blockIdx.x, j = threadIdx.x;
  int my_id = -1, start = 0, real = 0, fr = 0;
  if (j < f) {
    int s0 = 0;
    for (int i = 0; i < max_ids; ++i) {
      const int ns = slots[b * max_ids + i];
      if (j >= s0 && j < s0 + ns) { my_id = i; start = s0; }
      s0 += ns;
    }
    if (my_id >= 0) {
      real = j - start < n_real[b * max_ids + my_id];
      if (real) {
        fr = frame_no[b * f + j];
      } else {                                   // padding repeats the clip's largest frame number so far (0 if none)
        int s1 = 0;
        for (int i = 0; i <= my_id; ++i) {
          const int nr = n_real[b * max_ids + i];
          for (int k = 0; k < nr; ++k) fr = max(fr, frame_no[b * f + s1 + k]);
          s1 += slots[b * max_ids + i];
        }
      }
    }
    ident[j] = my_id;
    frames[j] = fr;
  }
  __syncthreads();
  if (j >= f) return;
  int is_first = 1;
  for (int k = 0; k < j; ++k) is_first &= frames[k] != fr;
  first[j] = is_first;
  __syncthreads();
  int rank = 1;                                  // 1-based rank among the distinct frame numbers
  for (int k = 0; k < f; ++k) rank += first[k] && frames[k] < fr;
  int bucket = 0;
  if (real) {
    const int r = ratio[b * f + j];
    bucket = r <= 5 ? 1 : min(20, (r + 4) / 5);  // (0..5)->1, (6..10)->2, ..., (96..100)->20; larger ratios clamp to 20
  }
  size_emb[b * f + j] = bucket;
  mask[b * f + j] = (real || !mask_padding) ? 1 : 0;
  for (int k = 0; k < f; ++k) idmask[((size_t)b * f + j) * f + k] = my_id >= 0 && ident[k] == my_id;
  long long* pos = positions + (size_t)b * (1 + f * n);
  if (j == 0) pos[0] = 0;
  for (int p = 0; p < n; ++p) pos[1 + j * n + p] = (long long)(rank - 1) * n + 1 + p;
}
}  // namespace
}  // namespace mt

extern "C" int mt_clip_meta_fwd(const int32_t* slots, const int32_t* n_real, const int32_t* frame_no, const int32_t* ratio,
                                int max_identities, int mask_padding, uint8_t* mask, uint8_t* identities_mask,
                                int32_t* size_embedding, int64_t* positions, int batch, int f, int n_patches, void* stream) {
  MT_REQUIRE(slots && n_real && frame_no && ratio && mask && identities_mask && size_embedding && positions,
             "clip_meta: null pointer");
  MT_REQUIRE(batch > 0 && f >= 1 && f <= 64 && n_patches >= 1 && max_identities >= 1, "clip_meta: bad shape B=%d f=%d n=%d ids=%d",
             batch, f, n_patches, max_identities);
  mt::clip_meta_kernel<<<batch, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      slots, n_real, frame_no, ratio, max_identities, mask_padding, mask, identities_mask, size_embedding,
      reinterpret_cast<long long*>(positions), f, n_patches);
  MT_LAUNCH_CHECK("clip_meta_kernel");
  return MT_OK;
}
