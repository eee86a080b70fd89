// EfficientNet-B0 feature extractor, eval mode (reference models/efficientnet/efficientnet_pytorch/
// model.py:267-288): stem -> 16 MBConv blocks -> head, all activations NHWC.
// The 1x1 convolutions run in gemm.cu; this file holds the stencil / reduction kernels
// (stem 3x3, depthwise kxk + BN + swish + SE squeeze, SE excitation) and the layer schedule.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "attention_mma.cuh"   // mma.sync / pack2 wrappers (stem_tc_kernel)
#include "dwconv_simt.cuh"
#include "mbconv_fused.cuh"

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
// one thread = one output pixel x all 32 output channels; filters are broadcast from shared memory
// (one LDS.128 feeds 4 FMAs), the 64-byte (bf16) output row of a pixel is written by its own thread so a
// warp writes 2 KiB contiguous.
// ---------------------------------------------------------------------------------------------------
template <typename T, typename TIN>
__global__ void __launch_bounds__(128) stem_kernel(const TIN* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ shift, T* __restrict__ out, int n_img,
                                                   int H, int W, int Ho, int Wo, int pad_lo) {
  __shared__ __align__(16) float ws[27 * 32];
  __shared__ __align__(16) float sh[32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < 32) sh[threadIdx.x] = shift[threadIdx.x];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)n_img * Ho * Wo) return;
  const int ox = (int)(pix % Wo);
  const int oy = (int)((pix / Wo) % Ho);
  const int img = (int)(pix / ((long long)Wo * Ho));
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
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
        const float4* wr = reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * 3 + ci) * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 ww = wr[q];
          acc[q * 4 + 0] = fmaf(v, ww.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(v, ww.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(v, ww.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(v, ww.w, acc[q * 4 + 3]);
        }
      }
    }
  }
  constexpr bool kExact = sizeof(T) == 4;
  T* orow = out + (size_t)pix * 32;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = silu<kExact>(acc[o * 8 + i] + sh[o * 8 + i]);
    store8(orow + o * 8, v);
  }
}

// ---------------------------------------------------------------------------------------------------
// stem on tensor cores (bf16 path): the 3x3x3 stencil is a K = 27 contraction -- per output pixel 864 FMAs
// on the CUDA cores (the FFMA kernel above is issue bound at 0.47 ms / 512 images) or 2 k-steps x 4 n-tiles of
// mma.sync m16n8k16 per 16 pixels.  A block stages the 2*TH+1 input rows of TH output rows in shared memory
// as bf16 (pixel values 0..255 are exact in bf16); with K ordered (ky, j = kx*3+ci padded to 10) the A
// fragment of a pixel is three contiguous 9-element runs of those rows, so it is read with plain 32-bit
// shared loads -- no im2col buffer.  Weights are bf16-rounded (B fragments in registers), accumulation
// fp32, BN shift + swish in the epilogue, output staged per warp for 512-byte coalesced stores.
// Requires W % 16 == 0 and an even H (TF-SAME pad_lo = 0); other shapes take the FFMA kernel.
// ---------------------------------------------------------------------------------------------------
constexpr int kStemTH = 8;        // output rows per block
constexpr int kStemRS = 688;      // shared-memory row pitch in elements (>= 224*3 + 10, multiple of 8)

__device__ __forceinline__ void stem_cvt_store(bf16* dst, const float* src) {     // 8 values
  const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
  *reinterpret_cast<uint4*>(dst) = make_uint4(attn::pack2(a.x, a.y), attn::pack2(a.z, a.w), attn::pack2(b.x, b.y),
                                              attn::pack2(b.z, b.w));
}
__device__ __forceinline__ void stem_cvt_store(bf16* dst, const uint8_t* src) {   // 8 values
  const uint2 u = *reinterpret_cast<const uint2*>(src);
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t w = i < 2 ? u.x : u.y;
    const uint32_t lo = (w >> ((i & 1) * 16)) & 0xffu, hi = (w >> ((i & 1) * 16 + 8)) & 0xffu;
    o[i] = attn::pack2((float)lo, (float)hi);
  }
  *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
}

template <typename TIN>
__global__ void __launch_bounds__(256) stem_tc_kernel(const TIN* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ shift, bf16* __restrict__ out, int H,
                                                      int W, int Ho, int Wo) {
  __shared__ __align__(16) bf16 rows[(2 * kStemTH + 1) * kStemRS];
  __shared__ __align__(16) uint8_t ostage[8][16 * 64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int img = blockIdx.y, oy0 = blockIdx.x * kStemTH;
  // ---- B fragments: k = ky*10 + j, j = kx*3 + ci (j = 9 and k >= 30 are zero rows); n = nt*8 + g
  uint32_t bfr[2][4][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        float wv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kk = ks * 16 + hi * 8 + 2 * t + e;
          const int ky = kk / 10, j = kk - ky * 10;
          wv[e] = (ky < 3 && j < 9) ? w[(ky * 9 + j) * 32 + nt * 8 + g] : 0.f;
        }
        bfr[ks][nt][hi] = attn::pack2(wv[0], wv[1]);
      }
  // per-lane element offsets of the A fragment pairs (k even -> j even -> 4-byte aligned)
  int aoff[2][2];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      int kk = ks * 16 + hi * 8 + 2 * t;
      if (kk >= 30) kk = 28;                       // zero weights: any finite value will do
      const int ky = kk / 10;
      aoff[ks][hi] = ky * kStemRS + (kk - ky * 10);
    }
  float2 hsh[4];                                   // BN shift / 2 for columns nt*8 + 2t, +1
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) hsh[nt] = make_float2(0.5f * shift[nt * 8 + 2 * t], 0.5f * shift[nt * 8 + 2 * t + 1]);

  // ---- stage the input rows 2*oy0 .. 2*oy0 + 2*TH (zero beyond the image: TF-SAME pads bottom / right)
  const int row_elems = W * 3;
  for (int i = tid; i < (2 * kStemTH + 1) * (kStemRS / 8); i += 256) {
    const int r = i / (kStemRS / 8), e = (i - r * (kStemRS / 8)) * 8;
    const int iy = 2 * oy0 + r;
    bf16* dst = rows + r * kStemRS + e;
    if (iy < H && e < row_elems) stem_cvt_store(dst, x + ((size_t)img * H + iy) * row_elems + e);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();

  const int mt_per_row = (Wo + 15) / 16;
  uint8_t* stg = ostage[warp];
  for (int mt = warp; mt < kStemTH * mt_per_row; mt += 8) {
    const int oyl = mt / mt_per_row, ox0 = (mt - oyl * mt_per_row) * 16;
    const int oy = oy0 + oyl;
    if (oy >= Ho) break;
    const bf16* base = rows + (2 * oyl) * kStemRS + 6 * (ox0 + g);
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(base + aoff[ks][0]);
      a[1] = *reinterpret_cast<const uint32_t*>(base + 48 + aoff[ks][0]);      // pixel + 8
      a[2] = *reinterpret_cast<const uint32_t*>(base + aoff[ks][1]);
      a[3] = *reinterpret_cast<const uint32_t*>(base + 48 + aoff[ks][1]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) attn::mma_bf16(acc[nt], a, bfr[ks][nt][0], bfr[ks][nt][1]);
    }
    // epilogue: rows g / g+8, columns nt*8 + 2t, +1 -> per-warp staging (16 px x 64 B, 16-byte chunks swizzled)
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int px = g + hr * 8;
        const float2 v = silu2(make_float2(acc[nt][hr * 2], acc[nt][hr * 2 + 1]), hsh[nt]);
        *reinterpret_cast<uint32_t*>(stg + px * 64 + ((nt ^ ((px >> 1) & 3)) << 4) + 4 * t) = attn::pack2(v.x, v.y);
      }
    }
    __syncwarp();
    bf16* orow = out + (((size_t)img * Ho + oy) * Wo + ox0) * 32;
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int px = (lane >> 2) + hr * 8, c = lane & 3;
      if (ox0 + px < Wo)
        *reinterpret_cast<uint4*>(orow + px * 32 + c * 8) =
            *reinterpret_cast<const uint4*>(stg + px * 64 + ((c ^ ((px >> 1) & 3)) << 4));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// depthwise kxk stride s + BN + swish, and per-(image, chunk, channel) partial sums for the
// squeeze-excite average pool (model.py:105-107,110).
//
// A thread owns 8 channels (one 16-byte NHWC vector) of an R x SX patch of output pixels and walks the
// (R-1)*S+K input rows it needs once, holding one input row segment in registers: every input vector
// is loaded once per thread and feeds up to K*R taps.  Lanes run over channel octets first, so a warp
// reads whole contiguous pixel rows (C*2 bytes each).  Filters come from L1 (16 B per tap per thread).
// blockDim = n_oct * SPB (SPB patches side by side), a block makes kDwPasses passes; pool partials are
// reduced per block in a fixed order (deterministic, no atomics).
// ---------------------------------------------------------------------------------------------------
template <typename T> struct Raw8;
template <> struct Raw8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) { load8(p, v); }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
  __device__ __forceinline__ float get(int i) const { return v[i]; }
};
template <> struct Raw8<bf16> {
  uint4 u;
  __device__ __forceinline__ void load(const bf16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { u = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ float get(int i) const {
    const uint32_t w = i < 2 ? u.x : (i < 4 ? u.y : (i < 6 ? u.z : u.w));
    return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
  }
};

constexpr int kDwPasses = 4;
__host__ __device__ constexpr int dw_sx(int s) { return s == 1 ? 4 : 2; }   // output columns per thread
// output rows per thread: 2 for 3x3 (4 input rows feed 2 output rows); 1 for 5x5, whose 2-row variant
// needs 168 registers + spills and leaves 9 warps per SM -- too few to hide the load latency
__host__ __device__ constexpr int dw_rows(int k) { return k == 3 ? 2 : 1; }

struct DwGeom {
  int n_oct, spb, threads, patches_x, patches, per_block, chunks;
};
inline DwGeom dw_geom(int H, int W, int C, int k, int s) {
  DwGeom g;
  const int Ho = (H + s - 1) / s, Wo = (W + s - 1) / s;
  g.n_oct = C / 8;
  g.spb = std::max(1, (256 + g.n_oct / 2) / g.n_oct);
  if (g.n_oct * g.spb > 320) g.spb = std::max(1, 320 / g.n_oct);
  g.threads = g.n_oct * g.spb;
  g.patches_x = (Wo + dw_sx(s) - 1) / dw_sx(s);
  g.patches = g.patches_x * ((Ho + dw_rows(k) - 1) / dw_rows(k));
  g.per_block = g.spb * kDwPasses;
  g.chunks = (g.patches + g.per_block - 1) / g.per_block;
  return g;
}

template <typename T, int K, int S>
__global__ void __launch_bounds__(320) dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w,
                                                      const float* __restrict__ shift, T* __restrict__ out,
                                                      float* __restrict__ pool_part, int H, int W, int Ho, int Wo, int C,
                                                      int pad_lo, int n_oct, int spb, int patches_x, int patches) {
  constexpr int SX = dw_sx(S), R = dw_rows(K);
  constexpr int NIN = (SX - 1) * S + K;          // input columns feeding SX outputs
  constexpr int NROW = (R - 1) * S + K;          // input rows feeding R output rows
  extern __shared__ float part[];                // [blockDim][8]
  const int tid = threadIdx.x;
  const int oct = tid % n_oct, slot = tid / n_oct;
  const int c0 = oct * 8;
  const int img = blockIdx.y;
  const T* in_img = in + (size_t)img * H * W * C + c0;
  T* out_img = out + (size_t)img * Ho * Wo * C + c0;
  float sh[8];
  load8(shift + c0, sh);
  float psum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) psum[i] = 0.f;
  constexpr bool kExact = sizeof(T) == 4;

  for (int pass = 0; pass < kDwPasses; ++pass) {
    const int pid = (blockIdx.x * kDwPasses + pass) * spb + slot;
    if (pid >= patches) break;
    const int oy0 = (pid / patches_x) * R, ox0 = (pid % patches_x) * SX;
    const int iy0 = oy0 * S - pad_lo, ix0 = ox0 * S - pad_lo;
    // accumulators and the input row window as packed float2 pairs: Blackwell's FFMA2 (fma.rn.f32x2) issues
    // two fp32 FMAs per instruction; inputs are converted bf16 -> fp32 ONCE when loaded (each vector then feeds
    // up to K*R taps), which halves the instruction count of the lazy-conversion version
    float2 acc[R][SX][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < SX; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][j][q] = make_float2(0.f, 0.f);

#pragma unroll
    for (int ir = 0; ir < NROW; ++ir) {
      const int iy = iy0 + ir;
      if (iy < 0 || iy >= H) continue;          // zero padding rows contribute nothing
      float2 row[NIN][4];
      const T* rp = in_img + (size_t)iy * W * C;
#pragma unroll
      for (int jj = 0; jj < NIN; ++jj) {
        const int ix = ix0 + jj;
        if (ix >= 0 && ix < W) {
          float v[8];
          load8(rp + (size_t)ix * C, v);
#pragma unroll
          for (int q = 0; q < 4; ++q) row[jj][q] = make_float2(v[2 * q], v[2 * q + 1]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) row[jj][q] = make_float2(0.f, 0.f);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int ky = ir - r * S;               // compile-time after unrolling
        if (ky < 0 || ky >= K) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          float wv[8];
          load8(w + (size_t)(ky * K + kx) * C + c0, wv);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 w2 = make_float2(wv[2 * q], wv[2 * q + 1]);
#pragma unroll
            for (int j = 0; j < SX; ++j) acc[r][j][q] = __ffma2_rn(row[j * S + kx][q], w2, acc[r][j][q]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int oy = oy0 + r;
      if (oy >= Ho) continue;
#pragma unroll
      for (int j = 0; j < SX; ++j) {
        const int ox = ox0 + j;
        if (ox >= Wo) continue;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[i] = silu<kExact>(((i & 1) ? acc[r][j][i >> 1].y : acc[r][j][i >> 1].x) + sh[i]);
          psum[i] += v[i];
        }
        store8(out_img + ((size_t)oy * Wo + ox) * C, v);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[tid * 8 + i] = psum[i];
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    const int o = c >> 3, ch = c & 7;
    float s = 0.f;
    for (int sl = 0; sl < spb; ++sl) s += part[(sl * n_oct + o) * 8 + ch];
    pool_part[((size_t)img * gridDim.x + blockIdx.x) * C + c] = s;   // one writer per entry
  }
}

// ---------------------------------------------------------------------------------------------------
// SE excitation (model.py:110-115): mean = sum_chunks(pool_part)/hw; gate = sigmoid(We*swish(Wr*mean+br)+be).
// A separate launch on purpose.  Folding it into the depthwise kernels as a "last block of the image (group) computes the
// gate" tail was built and measured in round 2: with the kernels' static work striding the block that computes a tail
// falls behind, so it is the last arriver of its next image as well and ends up computing ALL of its images' gates
// serially (k5 C1152 H7: 75 -> 1143 us; k3 C32 H112: 192 -> 657 us); and even with dynamic work claims the last groups'
// tails land after the depthwise work, where one 128..256-thread block needs about as long for the ~10 dependent L2
// round trips as this 512-thread kernel does including its launch.
// The arithmetic is tiny (2*SQ*C MACs per image); what costs is streaming the two weight matrices
// (up to 2 x 48 x 1152 floats) from L2, so one block handles kSeImgs images and every weight it loads feeds
// kSeImgs FMAs.  `we_t` is the expand weight transposed to [SQ][C] (lanes read consecutive channels).
// Fixed reduction orders: deterministic.
// ---------------------------------------------------------------------------------------------------
constexpr int kSeImgs = 4;
constexpr int kSeThreads = 512;
constexpr int kSeRowBatch = 12;          // rows of we_t per phase-3 work item
__global__ void __launch_bounds__(kSeThreads) se_gate_kernel(const float* __restrict__ pool_part, int n_chunks, float inv_hw,
                                                             const float* __restrict__ wr, const float* __restrict__ br,
                                                             const float* __restrict__ we_t, const float* __restrict__ be,
                                                             float* __restrict__ gate, int n_img, int C, int SQ) {
  // One block = kSeImgs images, one block per SM at most: the kernel is a chain of L2 round trips, so every
  // phase issues all the loads a thread needs as one batch of independent requests.
  extern __shared__ __align__(16) float sm[];
  float* mean = sm;                      // [kSeImgs][C]
  float* sq = sm + kSeImgs * C;          // [kSeImgs][SQ]  (SQ rounded up to 4)
  float* part3 = sq + kSeImgs * ((SQ + 3) & ~3);   // [JG][kSeImgs][C] phase-3 partial sums
  const int img0 = blockIdx.x * kSeImgs;
  const int n_here = min(kSeImgs, n_img - img0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kSeThreads / 32;
  // ---- phase 1: mean over the output pixels = sum of the per-tile partial sums / hw
  for (int i0 = tid; i0 < kSeImgs * C; i0 += 4 * kSeThreads) {
    float v[4][8];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // fixed order, double: deterministic and as accurate as the reference's mean
    for (int j0 = 0; j0 < n_chunks; j0 += 8) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kSeThreads, g = i / C, c = i - g * C;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          v[u][j] = (i < kSeImgs * C && g < n_here && j0 + j < n_chunks)
                        ? pool_part[((size_t)(img0 + g) * n_chunks + j0 + j) * C + c] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[u] += (double)v[u][j];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * kSeThreads < kSeImgs * C) mean[i0 + u * kSeThreads] = (float)(acc[u] * (double)inv_hw);
  }
  __syncthreads();
  // ---- phase 2: s[g][j] = swish(br[j] + sum_c wr[j][c] * mean[g][c]); one warp per squeeze row j, the row
  // fetched as up to 9 independent 16-byte loads per lane (C % 4 == 0): one L2 round trip per row
  const int nq = C >> 2;                 // float4 per row
  for (int j = warp; j < SQ; j += kWarps) {
    const float4* wrow = reinterpret_cast<const float4*>(wr + (size_t)j * C);
    float a[kSeImgs];
#pragma unroll
    for (int g = 0; g < kSeImgs; ++g) a[g] = 0.f;
    constexpr int kB2 = 9;
    for (int q0 = lane; q0 < nq; q0 += 32 * kB2) {
      float4 w4[kB2];
#pragma unroll
      for (int u = 0; u < kB2; ++u) w4[u] = (q0 + 32 * u < nq) ? wrow[q0 + 32 * u] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < kB2; ++u) {
        if (q0 + 32 * u < nq) {
#pragma unroll
          for (int g = 0; g < kSeImgs; ++g) {
            const float4 m = *reinterpret_cast<const float4*>(mean + g * C + 4 * (q0 + 32 * u));
            a[g] = fmaf(w4[u].x, m.x, fmaf(w4[u].y, m.y, fmaf(w4[u].z, m.z, fmaf(w4[u].w, m.w, a[g]))));
          }
        }
      }
    }
#pragma unroll
    for (int g = 0; g < kSeImgs; ++g) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[g] += __shfl_xor_sync(0xffffffffu, a[g], o);
    }
    if (lane == 0) {
      const float b = br[j];
#pragma unroll
      for (int g = 0; g < kSeImgs; ++g) sq[g * SQ + j] = silu<true>(a[g] + b);
    }
  }
  __syncthreads();
  // ---- phase 3: gate[g][c] = sigmoid(be[c] + sum_j we_t[j][c] * s[g][j]).  Work item = (4 channels, 12 rows):
  // 12 independent 16-byte loads, partial sums to shared memory, then a fixed-order sum over the row groups.
  const int JG = (SQ + kSeRowBatch - 1) / kSeRowBatch;
  for (int item = tid; item < nq * JG; item += kSeThreads) {
    const int jg = item / nq, q = item - jg * nq, j0 = jg * kSeRowBatch;
    float4 w4[kSeRowBatch];
#pragma unroll
    for (int u = 0; u < kSeRowBatch; ++u)
      w4[u] = (j0 + u < SQ) ? *reinterpret_cast<const float4*>(we_t + (size_t)(j0 + u) * C + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a[kSeImgs];
#pragma unroll
    for (int g = 0; g < kSeImgs; ++g) a[g] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < kSeRowBatch; ++u) {
      if (j0 + u < SQ) {
#pragma unroll
        for (int g = 0; g < kSeImgs; ++g) {
          const float sv = sq[g * SQ + j0 + u];
          a[g].x = fmaf(w4[u].x, sv, a[g].x); a[g].y = fmaf(w4[u].y, sv, a[g].y);
          a[g].z = fmaf(w4[u].z, sv, a[g].z); a[g].w = fmaf(w4[u].w, sv, a[g].w);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < kSeImgs; ++g) *reinterpret_cast<float4*>(part3 + ((size_t)jg * kSeImgs + g) * C + 4 * q) = a[g];
  }
  __syncthreads();
  for (int i = tid; i < n_here * nq; i += kSeThreads) {
    const int g = i / nq, q = i - g * nq;
    float4 a = *reinterpret_cast<const float4*>(be + 4 * q);
    for (int jg = 0; jg < JG; ++jg) {
      const float4 p = *reinterpret_cast<const float4*>(part3 + ((size_t)jg * kSeImgs + g) * C + 4 * q);
      a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    *reinterpret_cast<float4*>(gate + (size_t)(img0 + g) * C + 4 * q) =
        make_float4(sigmoidf_<true>(a.x), sigmoidf_<true>(a.y), sigmoidf_<true>(a.z), sigmoidf_<true>(a.w));
  }
}

template <typename T, typename TIN>
int launch_stem_t(const void* x, const float* w, const float* shift, void* out, int n_img, int H, int W,
                  cudaStream_t st) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = (long long)n_img * Ho * Wo;
  const int grid = (int)((total + 127) / 128);
  ProfScope prof(st, 2.0 * 27 * 32 * (double)n_img * Ho * Wo,
                 (double)n_img * ((double)H * W * 3 * sizeof(TIN) + (double)Ho * Wo * 32 * sizeof(T)), "stem");
  if constexpr (sizeof(T) == 2) {
    if (W % 16 == 0 && H % 2 == 0 && W * 3 + 10 <= kStemRS && n_img <= 65535) {
      dim3 grid((Ho + kStemTH - 1) / kStemTH, n_img);
      stem_tc_kernel<TIN><<<grid, 256, 0, st>>>(reinterpret_cast<const TIN*>(x), w, shift, reinterpret_cast<bf16*>(out), H, W,
                                                 Ho, Wo);
      MT_LAUNCH_CHECK("stem_tc_kernel");
      return MT_OK;
    }
  }
  stem_kernel<T, TIN><<<grid, 128, 0, st>>>(reinterpret_cast<const TIN*>(x), w, shift, reinterpret_cast<T*>(out),
                                            n_img, H, W, Ho, Wo, same_pad_lo(H, 3, 2));
  MT_LAUNCH_CHECK("stem_kernel");
  return MT_OK;
}

template <typename T, int K, int S>
int launch_dw_ks(const T* i, const float* w, const float* shift, T* o, float* pool, int n_img, int H, int W, int C,
                 cudaStream_t st) {
  const int Ho = (H + S - 1) / S, Wo = (W + S - 1) / S;
  const DwGeom g = dw_geom(H, W, C, K, S);
  dim3 grid(g.chunks, n_img);
  const size_t smem = sizeof(float) * (size_t)g.threads * 8;
  dwconv_kernel<T, K, S><<<grid, g.threads, smem, st>>>(i, w, shift, o, pool, H, W, Ho, Wo, C, same_pad_lo(H, K, S),
                                                        g.n_oct, g.spb, g.patches_x, g.patches);
  MT_LAUNCH_CHECK("dwconv_kernel");
  return MT_OK;
}

template <typename T>
int launch_dw_t(const void* in, const float* w, const float* shift, void* out, float* pool, int n_img, int H, int W,
                int C, int k, int s, cudaStream_t st) {
  const int Ho = (H + s - 1) / s, Wo = (W + s - 1) / s;
  const T* i = reinterpret_cast<const T*>(in);
  T* o = reinterpret_cast<T*>(out);
  ProfScope prof(st, 2.0 * k * k * (double)n_img * Ho * Wo * C,
                 (double)n_img * C * ((double)H * W + (double)Ho * Wo) * sizeof(T), "dwconv k%d s%d C%d H%d", k, s, C, H);
  if (k == 3 && s == 1) return launch_dw_ks<T, 3, 1>(i, w, shift, o, pool, n_img, H, W, C, st);
  if (k == 3 && s == 2) return launch_dw_ks<T, 3, 2>(i, w, shift, o, pool, n_img, H, W, C, st);
  if (k == 5 && s == 1) return launch_dw_ks<T, 5, 1>(i, w, shift, o, pool, n_img, H, W, C, st);
  if (k == 5 && s == 2) return launch_dw_ks<T, 5, 2>(i, w, shift, o, pool, n_img, H, W, C, st);
  set_error("dwconv: unsupported kernel %d / stride %d", k, s);
  return MT_ERR_UNSUPPORTED;
}

int device_sms() { return current_sms(); }

template <int K, int S, int MODE = 0>
int launch_dw_simt_ks(const CUtensorMap& tm, const float* w, const float* shift, bf16* o, float* pool, int n_img, int H,
                      int C, const DwSimtGeom& g, cudaStream_t st, int relu_in = 0) {
  const int Ho = (H + S - 1) / S;
  using Kern = void (*)(const CUtensorMap, const float*, const float*, bf16*, float*, int, int, int, int, int, DwSimtGeom, int);
  const int slot = g.CW == 32 ? 1 : (g.CW == 48 ? 2 : (g.CW == 64 ? 3 : 0));
  static const Kern kerns[4] = {dwconv_simt_kernel<K, S, 0, MODE>, dwconv_simt_kernel<K, S, 32, MODE>,
                                dwconv_simt_kernel<K, S, 48, MODE>, dwconv_simt_kernel<K, S, 64, MODE>};
  Kern kern = kerns[slot];
  if (first_use_on_device(reinterpret_cast<const void*>(kern))) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(dwconv_simt)");
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  // persistent blocks, statically scheduled: launch exactly as many as are co-resident (registers included),
  // so no block waits for a second wave
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, g.threads, g.smem) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const long long work = (long long)n_img * g.tiles;
  const int workers = (int)std::max(1LL, std::min(work, (long long)(device_sms() * per_sm) / g.n_cchunks));
  dim3 grid(workers, g.n_cchunks);
  kern<<<grid, g.threads, g.smem, st>>>(tm, w, shift, o, pool, n_img, Ho, Ho, C, same_pad_lo(H, K, S), g, relu_in);
  MT_LAUNCH_CHECK("dwconv_simt_kernel");
  return MT_OK;
}

// smem-staged FFMA2 kernel (square feature maps, as everywhere in B0)
int launch_dw_simt(const void* in, const float* w, const float* shift, void* out, float* pool, int n_img, int H, int C,
                   int k, int s, cudaStream_t st) {
  DwSimtGeom g;
  if (!dw_simt_geom(&g, H, H, C, k, s, n_img, device_sms())) {
    set_error("dwconv: no tile fits for h=%d c=%d k=%d s=%d", H, C, k, s);
    return MT_ERR_UNSUPPORTED;
  }
  CUtensorMap tm;
  int rc = make_tmap_nhwc_bf16_plain(&tm, in, n_img, H, H, C, g.CW, g.IW, g.IH);
  if (rc) return rc;
  const int Ho = (H + s - 1) / s;
  ProfScope prof(st, 2.0 * k * k * (double)n_img * Ho * Ho * C, (double)n_img * C * ((double)H * H + (double)Ho * Ho) * 2,
                 "dwconv_simt k%d s%d C%d H%d", k, s, C, H);
  bf16* o = reinterpret_cast<bf16*>(out);
  if (k == 3 && s == 1) return launch_dw_simt_ks<3, 1>(tm, w, shift, o, pool, n_img, H, C, g, st);
  if (k == 3 && s == 2) return launch_dw_simt_ks<3, 2>(tm, w, shift, o, pool, n_img, H, C, g, st);
  if (k == 5 && s == 1) return launch_dw_simt_ks<5, 1>(tm, w, shift, o, pool, n_img, H, C, g, st);
  return launch_dw_simt_ks<5, 2>(tm, w, shift, o, pool, n_img, H, C, g, st);
}

}  // namespace

// Plain 3x3 / pad 1 / stride 1 depthwise convolution on the same kernel (MODE 1: no shift, no activation, no pool sums; optional
// ReLU on the input): SeparableConv2d.conv1 of the Xception extractor, bf16 NHWC.  w: f32 [9][C] tap-major.
int launch_dw_plain_bf16(const void* in, const float* w, void* out, int n_img, int H, int C, int relu_in, cudaStream_t st) {
  DwSimtGeom g;
  if (!dw_simt_geom(&g, H, H, C, 3, 1, n_img, device_sms())) {
    set_error("dwconv(plain): no tile fits for h=%d c=%d", H, C);
    return MT_ERR_UNSUPPORTED;
  }
  CUtensorMap tm;
  int rc = make_tmap_nhwc_bf16_plain(&tm, in, n_img, H, H, C, g.CW, g.IW, g.IH);
  if (rc) return rc;
  ProfScope prof(st, 18.0 * (double)n_img * H * H * C, (double)n_img * C * (double)H * H * 4, "xc_dw3x3 C%d H%d", C, H);
  return launch_dw_simt_ks<3, 1, 1>(tm, w, nullptr, reinterpret_cast<bf16*>(out), nullptr, n_img, H, C, g, st, relu_in);
}

namespace {

bool dw_simt_ok(int h, int w_, int c, int k, int s) { return h == w_ && (k == 3 || k == 5) && (s == 1 || s == 2); }

// ---- fused expand + depthwise (mbconv_fused.cuh).  Measured on B200 at 512 images (us, fused vs expand GEMM +
// depthwise kernel): block 1 (16->96, 112^2) 503 vs 616; block 2 (24->144, 56^2) 395 vs 383; block 3 328 vs 304;
// block 4 (40->240, 28^2) 316 vs 248; block 5 140 vs 118 -- the fused kernel wins where the expand GEMM is at its
// worst (K = 16: 32-byte operand rows) and loses where the stencil dominates (7 stencil warps per SM against 12-16
// in the stand-alone kernel).  mt_mbconv_fwd therefore fuses blocks with cin <= 16 by default;
// MINTIME_B200_FUSE=all fuses every block that has a schedule, MINTIME_B200_FUSE=0 none.
int fuse_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("MINTIME_B200_FUSE");
    mode = 1;
    if (e && e[0] == '0') mode = 0;
    if (e && !strcmp(e, "all")) mode = 2;
  }
  return mode;
}

bool fused_front_geom(FusedGeom* g, int precision, int h, int cin, int cexp, int k, int s, int n_img) {
  if (precision != MT_PREC_BF16 || cexp == cin) return false;
  if ((k != 3 && k != 5) || (s != 1 && s != 2)) return false;
  return fused_geom(g, h, cin, cexp, k, s, n_img, 148);
}

// the choice mt_mbconv_fwd makes
bool fuse_block(FusedGeom* g, int precision, int h, int cin, int cexp, int k, int s, int n_img) {
  const int mode = fuse_mode();
  if (mode == 0 || (mode == 1 && cin > 16)) return false;
  return fused_front_geom(g, precision, h, cin, cexp, k, s, n_img);
}

template <int K, int S>
int launch_front_ks(const CUtensorMap& tin, const CUtensorMap& tw, const float* exp_shift, const float* w_dw,
                    const float* dw_shift, bf16* o, float* pool, int n_img, int H, int C, const FusedGeom& g, cudaStream_t st) {
  const int Ho = (H + S - 1) / S;
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const float*, const float*, const float*, bf16*, float*, int,
                        int, int, int, int, int, int, FusedGeom);
  const int slot = g.d.CW == 32 ? 0 : (g.d.CW == 48 ? 1 : 2);
  static const Kern kerns[3] = {mbconv_front_kernel<K, S, 32>, mbconv_front_kernel<K, S, 48>, mbconv_front_kernel<K, S, 64>};
  Kern kern = kerns[slot];
  if (first_use_on_device(reinterpret_cast<const void*>(kern))) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(mbconv_front)");
    // two ~100 KiB blocks per SM need the maximum shared-memory carve-out (the occupancy query honours it)
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  const int per_sm = 1;                          // one warp-specialised block per SM (it owns all 512 TMEM columns)
  const int threads = 32 + 32 * g.nd + ((g.d.threads + 31) & ~31);
  const long long work = (long long)n_img * g.d.tiles;
  const int workers = (int)std::max(1LL, std::min(work, (long long)(device_sms() * per_sm) / g.d.n_cchunks));
  dim3 grid(workers, g.d.n_cchunks);
  kern<<<grid, threads, g.smem, st>>>(tin, tw, exp_shift, w_dw, dw_shift, o, pool, n_img, H, H, Ho, Ho, C,
                                          same_pad_lo(H, K, S), g);
  MT_LAUNCH_CHECK("mbconv_front_kernel");
  return MT_OK;
}

int launch_front(const void* in, const void* w_exp, const float* exp_shift, const float* w_dw, const float* dw_shift,
                 void* out, float* pool, int n_img, int H, int cin, int cexp, int k, int s, const FusedGeom& g,
                 cudaStream_t st) {
  CUtensorMap tin, tw;
  int rc = make_tmap_nhwc_bf16_kmajor(&tin, in, n_img, H, H, cin, g.kbox, g.d.IW, g.d.IH);
  if (rc) return rc;
  rc = make_tmap_weights_kmajor(&tw, w_exp, cexp, cin, g.d.CW, g.kbox);
  if (rc) return rc;
  const int Ho = (H + s - 1) / s;
  ProfScope prof(st, 2.0 * (double)n_img * H * H * cin * cexp + 2.0 * k * k * (double)n_img * Ho * Ho * cexp,
                 (double)n_img * ((double)H * H * cin + (double)Ho * Ho * cexp) * 2, "mbconv_front k%d s%d C%d->%d H%d", k, s,
                 cin, cexp, H);
  bf16* o = reinterpret_cast<bf16*>(out);
  if (k == 3 && s == 1) return launch_front_ks<3, 1>(tin, tw, exp_shift, w_dw, dw_shift, o, pool, n_img, H, cexp, g, st);
  if (k == 3 && s == 2) return launch_front_ks<3, 2>(tin, tw, exp_shift, w_dw, dw_shift, o, pool, n_img, H, cexp, g, st);
  if (k == 5 && s == 1) return launch_front_ks<5, 1>(tin, tw, exp_shift, w_dw, dw_shift, o, pool, n_img, H, cexp, g, st);
  return launch_front_ks<5, 2>(tin, tw, exp_shift, w_dw, dw_shift, o, pool, n_img, H, cexp, g, st);
}

int dw_chunks(int precision, int h, int w_, int c, int k, int s) {
  // tensor-core kernels: a block sees every tile of an image -> one sum per (image, channel)
  if (precision == MT_PREC_BF16 && dw_simt_ok(h, w_, c, k, s)) {
    DwSimtGeom g;
    if (dw_simt_geom(&g, h, w_, c, k, s, 1, 148)) return g.tiles;
  }
  return dw_geom(h, w_, c, k, s).chunks;
}

int dwconv_dispatch(int precision, const void* in, const float* w, const float* shift, void* out, float* pool_part,
                    int n_img, int h, int w_, int c, int k, int s, cudaStream_t st) {
  MT_REQUIRE(in && w && shift && out && pool_part, "dwconv: null pointer");
  MT_REQUIRE(n_img > 0 && h > 0 && w_ > 0 && c >= 8 && c % 8 == 0 && c <= 2560,
             "dwconv: bad shape n=%d h=%d w=%d c=%d (c %% 8 == 0)", n_img, h, w_, c);
  MT_REQUIRE(n_img <= 65535, "dwconv: at most 65535 images per call (got %d)", n_img);
  if (precision == MT_PREC_FP32) return launch_dw_t<float>(in, w, shift, out, pool_part, n_img, h, w_, c, k, s, st);
  if (precision == MT_PREC_BF16) {
    if (dw_simt_ok(h, w_, c, k, s)) return launch_dw_simt(in, w, shift, out, pool_part, n_img, h, c, k, s, st);
    // non-square maps: register-strip kernel on global loads
    return launch_dw_t<bf16>(in, w, shift, out, pool_part, n_img, h, w_, c, k, s, st);
  }
  if (precision == 2) return launch_dw_t<bf16>(in, w, shift, out, pool_part, n_img, h, w_, c, k, s, st);  // debug: CUDA-core bf16
  set_error("dwconv: unknown precision %d", precision);
  return MT_ERR_ARG;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }


struct BlockWs { size_t exp, dw, pool, gate, counters, total; };
BlockWs block_ws_layout(const mt_mbconv_spec_t& b, int n_img, int precision) {
  const size_t es = precision == MT_PREC_FP32 ? 4 : 2;
  const size_t ho = (b.hw_in + b.stride - 1) / b.stride, cexp = (size_t)b.cin * b.expand;
  BlockWs l;
  size_t off = 0;
  l.exp = off;  off += b.expand != 1 ? align_up((size_t)b.hw_in * b.hw_in * cexp * n_img * es, 1024) : 0;
  l.dw = off;   off += align_up(ho * ho * cexp * n_img * es, 1024);
  size_t chunks = (size_t)dw_chunks(precision, b.hw_in, b.hw_in, (int)cexp, b.kernel, b.stride);
  FusedGeom fg;
  if (fuse_block(&fg, precision, b.hw_in, b.cin, (int)cexp, b.kernel, b.stride, n_img)) chunks = std::max(chunks, (size_t)fg.d.tiles);
  l.pool = off; off += align_up(chunks * cexp * n_img * 4, 1024);
  l.gate = off; off += align_up(cexp * n_img * 4, 1024);
  l.counters = off; off += align_up((size_t)n_img * 4, 1024);
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

extern "C" int mt_dwconv_chunks(int precision, int h, int w_, int c, int k, int s) {
  if (h <= 0 || w_ <= 0 || c < 8 || c % 8 != 0 || (s != 1 && s != 2) || (k != 3 && k != 5)) return 0;
  return dw_chunks(precision, h, w_, c, k, s);
}

extern "C" int mt_dwconv_fwd(int precision, const void* in, const float* w, const float* shift, void* out,
                             float* pool_part, int n_img, int h, int w_, int c, int k, int s, void* stream) {
  return dwconv_dispatch(precision, in, w, shift, out, pool_part, n_img, h, w_, c, k, s,
                         reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mt_expand_dwconv_chunks(int h, int cin, int cexp, int k, int s) {
  FusedGeom g;
  if (h <= 0 || !fused_front_geom(&g, MT_PREC_BF16, h, cin, cexp, k, s, 1)) return 0;
  return g.d.tiles;
}

extern "C" int mt_expand_dwconv_fwd(const void* in, const void* w_exp, const float* exp_shift, const float* w_dw,
                                    const float* dw_shift, void* out, float* pool_part, int n_img, int h, int cin,
                                    int cexp, int k, int s, void* stream) {
  MT_REQUIRE(in && w_exp && exp_shift && w_dw && dw_shift && out && pool_part, "expand_dwconv: null pointer");
  MT_REQUIRE(n_img > 0 && n_img <= 65535 && h > 0, "expand_dwconv: bad shape n=%d h=%d", n_img, h);
  FusedGeom g;
  if (!fused_front_geom(&g, MT_PREC_BF16, h, cin, cexp, k, s, n_img)) {
    set_error("expand_dwconv: no fused schedule for h=%d cin=%d cexp=%d k=%d s=%d", h, cin, cexp, k, s);
    return MT_ERR_UNSUPPORTED;
  }
  return launch_front(in, w_exp, exp_shift, w_dw, dw_shift, out, pool_part, n_img, h, cin, cexp, k, s, g,
                      reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int mt_se_gate_fwd(const float* pool_part, int n_chunks, int hw, const float* wr, const float* br,
                              const float* we, const float* be, float* gate, int n_img, int c, int sq, void* stream) {
  MT_REQUIRE(pool_part && wr && br && we && be && gate, "se_gate: null pointer");
  const int se_jg = (sq + kSeRowBatch - 1) / kSeRowBatch;
  const size_t se_smem = (size_t)kSeImgs * ((size_t)c + ((sq + 3) & ~3) + (size_t)se_jg * c) * 4;
  MT_REQUIRE(n_img > 0 && c > 0 && c % 4 == 0 && sq > 0 && sq <= 256 && hw > 0 && n_chunks > 0 && se_smem <= 200 * 1024,
             "se_gate: bad shape");
  if (first_use_on_device(reinterpret_cast<const void*>(se_gate_kernel))) {
    cudaError_t e = cudaFuncSetAttribute(se_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(se_gate)");
  }
  ProfScope prof(reinterpret_cast<cudaStream_t>(stream), 4.0 * n_img * (double)c * sq,
                 (double)n_img * c * 4 * (n_chunks + 1), "se_gate");
  se_gate_kernel<<<(n_img + kSeImgs - 1) / kSeImgs, kSeThreads, se_smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      pool_part, n_chunks, 1.0f / (float)hw, wr, br, we, be, gate, n_img, c, sq);
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
  int chunks = dw_chunks(precision, b.hw_in, b.hw_in, cexp, b.kernel, b.stride);
  FusedGeom fg;
  if (b.expand != 1 && fuse_block(&fg, precision, b.hw_in, b.cin, cexp, b.kernel, b.stride, n_img)) {
    // expand 1x1 + BN + swish + depthwise + BN + swish in one kernel: the expanded tensor stays on chip
    MT_REQUIRE(w->expand.w, "mbconv: expand weights missing");
    rc = launch_front(in, w->expand.w, w->expand.shift, w->dw_w, w->dw_shift, bdw, pool, n_img, b.hw_in, b.cin, cexp,
                      b.kernel, b.stride, fg, reinterpret_cast<cudaStream_t>(stream));
    if (rc) return rc;
    chunks = fg.d.tiles;
  } else {
    if (b.expand != 1) {
      MT_REQUIRE(w->expand.w, "mbconv: expand weights missing");
      rc = mt_pointwise_fwd(precision, in, w->expand.w, w->expand.shift, nullptr, 0, nullptr, 1, bexp,
                            n_img * b.hw_in * b.hw_in, cexp, b.cin, stream);
      if (rc) return rc;
      dw_in = bexp;
    }
    // depthwise (+ pool sums) and the SE excitation as two launches: the gate needs the whole image's pool
    rc = mt_dwconv_fwd(precision, dw_in, w->dw_w, w->dw_shift, bdw, pool, n_img, b.hw_in, b.hw_in, cexp, b.kernel,
                       b.stride, stream);
    if (rc) return rc;
  }
  rc = mt_se_gate_fwd(pool, chunks, ho * ho, w->se_reduce_w, w->se_reduce_b, w->se_expand_w, w->se_expand_b, gate, n_img,
                      cexp, sq, stream);
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
  // Strict algorithmic bytes of SURVEY.md 8(d): every layer's input read once and its output written once, activations
  // in T, the network input as fp32 -- 4.654 MB per image on the bf16 path (stem 1.405 + 16 MBConv blocks 3.091 + head
  // 0.157).  Booked on two nested scopes: the whole extractor and the MBConv blocks alone.
  const double es = precision == MT_PREC_FP32 ? 4.0 : 2.0;
  double blocks_bytes = 0.0, blocks_flops = 0.0;
  for (int i = 0; i < 16; ++i) {
    const BlockSpec& b = kBlocks[i];
    const double ho = (b.hw + b.s - 1) / b.s, cexp = (double)b.cin * b.e;
    blocks_bytes += es * ((double)b.hw * b.hw * b.cin + ho * ho * b.cout);
    blocks_flops += 2.0 * ((b.e != 1 ? (double)b.hw * b.hw * b.cin * cexp : 0.0) + (double)b.k * b.k * ho * ho * cexp +
                           ho * ho * cexp * b.cout);
  }
  const double stem_bytes = 224.0 * 224 * 3 * 4 + es * 112 * 112 * 32, head_bytes = es * 49 * (320 + 1280);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ProfScope whole(st, (double)n_img * (blocks_flops + 2.0 * 27 * 32 * 112 * 112 + 2.0 * 49 * 320 * 1280),
                  (double)n_img * (stem_bytes + blocks_bytes + head_bytes), "extractor (strict bytes)");
  int rc = mt_stem_fwd(precision, x, x_dtype, w->stem_w, w->stem_shift, cur, n_img, 224, 224, stream);
  if (rc) return rc;
  {
    ProfScope stack(st, (double)n_img * blocks_flops, (double)n_img * blocks_bytes, "mbconv_stack (strict bytes)");
    for (int i = 0; i < 16; ++i) {
      const mt_mbconv_spec_t b = spec_of(i);
      rc = mt_mbconv_fwd(precision, &b, &w->blocks[i], cur, nxt, n_img, ws + l.block, l.total - l.block, stream);
      if (rc) return rc;
      std::swap(cur, nxt);
    }
  }
  // head 1x1 (320 -> 1280) + BN + swish, written straight into the token layout the patch embedding reads
  return mt_pointwise_fwd(precision, cur, w->head.w, w->head.shift, nullptr, 0, nullptr, 1, feats, n_img * 49, 1280,
                          320, stream);
}
