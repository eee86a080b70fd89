// Depthwise kxk convolution + BN + swish + squeeze-excite pool partial sums, bf16 path
// (reference model.py:105-107,110).
//
// A depthwise stencil has 2*k*k flops per output element and no reuse across channels, so on B200 it is
// a CUDA-core job: tensor-core formulations spend 7/8 of their MACs on the zeros of a block-diagonal
// filter matrix and legacy mma.sync issues too slowly to make up for it (dwconv_tc.cuh, kept for
// comparison).  What makes the CUDA-core version fast is
//   * the input tile is staged in shared memory by ONE 4-D TMA load per tile (hardware zero fill = the
//     TF-"SAME" padding of utils.py:254-269), double buffered, so no thread ever waits on global memory;
//   * a thread owns one channel PAIR (one 32-bit bf16x2 word) of an R x 7 patch of output pixels; it
//     unpacks every input word once (2 ALU ops) and feeds it to up to k*R packed FFMA2 (fma.rn.f32x2:
//     two fp32 FMAs per issue slot), with the k*k filter taps of its channel pair held in registers
//     for the whole life of the persistent block;
//   * lanes run over channel pairs, so a warp reads / writes whole contiguous channel rows of a pixel.
// Pool sums are reduced per tile in a fixed order (deterministic, no atomics) and written as
// pool_part[image][tile][channel]; se_gate_kernel adds the tiles.
#pragma once
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "ptx.cuh"

namespace mt {

constexpr int kDwSX = 7;   // output columns per thread: divides every B0 feature-map width (112 ... 7)
__host__ __device__ constexpr int dw_simt_rows(int k, int s) { return (k == 3 && s == 1) ? 4 : 2; }

struct DwSimtGeom {
  int CW, CP, n_cchunks;     // channels per chunk, channel pairs per chunk, chunks
  int NS, threads;           // patches processed side by side; threads = CP * NS
  int TH, TW, IH, IW;        // output tile, input tile (pixels)
  int tiles_x, tiles_y, tiles;
  int strips_x, n_strips;    // R x 7 patches per tile
  int tile_bytes, tile_stride;
  int workers;               // persistent blocks per channel chunk
  size_t smem;
};

// Tile search: channel chunk CW (a divisor of C), output tile TH x TW and NS, minimising
// (wasted lanes) + (halo re-reads), under a 46 KiB input tile (2-deep ring, 2 blocks per SM).
inline bool dw_simt_geom(DwSimtGeom* out, int H, int W, int C, int k, int s, int n_img, int num_sms) {
  const int R = dw_simt_rows(k, s), SX = kDwSX;
  const int Ho = (H + s - 1) / s, Wo = (W + s - 1) / s;
  const int pad_h = std::max((Ho - 1) * s + k - H, 0) / 2, pad_w = std::max((Wo - 1) * s + k - W, 0) / 2;
  const int cap = 46 * 1024;
  double best = 1e30;
  DwSimtGeom g{};
  bool found = false;
  auto cdiv = [](int a, int b) { return (a + b - 1) / b; };
  // chunk widths with a compile-time kernel instantiation (immediate shared-memory offsets); any other
  // divisor of C (multiple of 8) runs the generic instantiation
  const bool fixed_cw = (C % 32 == 0) || (C % 48 == 0) || (C % 64 == 0);
  for (int CW = 8; CW <= 128 && CW <= C; CW += 8) {
    if (C % CW) continue;
    if (fixed_cw && CW != 32 && CW != 48 && CW != 64) continue;
    const int CP = CW / 2;
    for (int nx = 1; nx <= 8; ++nx) {
      int TW = cdiv(Wo, nx);
      if (TW >= SX) TW = cdiv(TW, SX) * SX;
      if (nx > 1 && cdiv(Wo, TW) != nx) continue;
      for (int ny = 1; ny <= Ho; ++ny) {
        int TH = cdiv(cdiv(Ho, ny), R) * R;
        if (ny > 1 && cdiv(Ho, TH) != ny) continue;
        const int IH = (TH - 1) * s + k, IW = (TW - 1) * s + k;
        const long long tb = (long long)IH * IW * CW * 2;
        if (tb > cap || IW > 256 || IH > 256) continue;
        const int tiles_x = cdiv(Wo, TW), tiles_y = cdiv(Ho, TH);
        const int n_strips = cdiv(TH, R) * cdiv(TW, SX);
        double halo = 0;
        for (int ty = 0; ty < tiles_y; ++ty)
          for (int tx = 0; tx < tiles_x; ++tx) {
            const int y0 = ty * TH * s - pad_h, x0 = tx * TW * s - pad_w;
            halo += (double)(std::min(H, y0 + IH) - std::max(0, y0)) * (std::min(W, x0 + IW) - std::max(0, x0));
          }
        halo /= (double)H * W;
        for (int NS = 1; NS <= 256 / CP; ++NS) {
          const int thr = CP * NS;
          if (thr < 96) continue;
          const int passes = cdiv(n_strips, NS);
          const double util = (double)Ho * Wo / ((double)passes * NS * R * SX * tiles_x * tiles_y);
          const double score = 1.0 / util + 0.5 * (halo - 1.0) + (thr < 192 ? 0.15 : 0.0) +
                               (tiles_x * tiles_y > 1 ? 0.02 : 0.0) - 1e-4 * CW;
          if (score < best) {
            best = score;
            found = true;
            g.CW = CW; g.CP = CP; g.n_cchunks = C / CW; g.NS = NS; g.threads = thr;
            g.TH = TH; g.TW = TW; g.IH = IH; g.IW = IW;
            g.tiles_x = tiles_x; g.tiles_y = tiles_y; g.tiles = tiles_x * tiles_y;
            g.strips_x = cdiv(TW, SX); g.n_strips = n_strips;
            g.tile_bytes = (int)tb; g.tile_stride = ((int)tb + 127) & ~127;
          }
        }
      }
    }
  }
  if (!found) return false;
  g.smem = 2 * (size_t)g.tile_stride + 2 * (size_t)g.threads * sizeof(float2) + 128;
  const int per_sm = std::max(1, std::min({(int)((220 * 1024) / (g.smem + 1024)), 2048 / g.threads, 4}));
  const long long work = (long long)n_img * g.tiles;
  g.workers = (int)std::max(1LL, std::min(work, (long long)(num_sms * per_sm + g.n_cchunks - 1) / g.n_cchunks));
  *out = g;
  return true;
}

// 4-D TMA descriptor over a bf16 NHWC tensor, box = CW channels x box_w x box_h x 1 image, no swizzle
int make_tmap_nhwc_bf16_plain(CUtensorMap_st* m, const void* base, int n, int h, int w, int c, int box_c, int box_w,
                              int box_h);

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
  return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// swish of (x + shift) for a channel pair: with h = (x + shift)/2, x*sigmoid(x) = h + h*tanh(h)
// (common.cuh silu<false>), as two packed FFMA2 around the two MUFU.TANH; `hsh` = shift/2.
__device__ __forceinline__ float2 silu2(float2 x, float2 hsh) {
  const float2 h = __ffma2_rn(x, make_float2(0.5f, 0.5f), hsh);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
  return __ffma2_rn(h, t, h);
}

// One R x 7 patch of outputs for one channel pair.  `tp` points at the channel pair's word of the patch's
// top-left INPUT pixel inside the shared-memory tile ([IH][IW][CW] bf16); pstride / rstride are the
// pixel / row pitches in bytes.
// MODE 0: the MBConv depthwise (BN shift + swish + squeeze-excite pool sums).  MODE 1: a plain depthwise convolution (no shift,
// no activation, no pool) with an optional ReLU folded into the loads -- SeparableConv2d.conv1 of the Xception extractor
// (reference models/xception.py:20 behind the ReLUs of :43-58).
template <int K, int S, int R, int MODE = 0>
__device__ __forceinline__ void dw_patch(const uint8_t* tp, const int pstride, const int rstride,
                                         const float2 (&wv)[K * K], float2 (&acc)[R][kDwSX], const bool relu_in = false) {
  constexpr int SX = kDwSX, NIN = (SX - 1) * S + K, NROW = (R - 1) * S + K;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < SX; ++j) acc[r][j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int ir = 0; ir < NROW; ++ir) {
    float2 row[NIN];
    const uint8_t* rp = tp + ir * rstride;
#pragma unroll
    for (int jj = 0; jj < NIN; ++jj) {
      row[jj] = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(rp + jj * pstride));
      if (MODE == 1 && relu_in) { row[jj].x = fmaxf(row[jj].x, 0.f); row[jj].y = fmaxf(row[jj].y, 0.f); }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int ky = ir - r * S;                 // compile-time after unrolling
      if (ky < 0 || ky >= K) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx)
#pragma unroll
        for (int j = 0; j < SX; ++j) acc[r][j] = __ffma2_rn(row[j * S + kx], wv[ky * K + kx], acc[r][j]);
    }
  }
}

// All R x 7 patches of one tile that belong to this thread's (channel pair, slot): stencil, BN shift + swish,
// bf16 store, and the thread's share of the squeeze-excite pool sum.  `tile` points at the channel pair's word
// of the tile's first input pixel.
template <int K, int S, int MODE = 0>
__device__ __forceinline__ float2 dw_tile(const uint8_t* tile, const int pstride, const int rstride,
                                          const float2 (&wv)[K * K], const float2 hsh, bf16* __restrict__ out, int img,
                                          int oy_t, int ox_t, int oy_end, int ox_end, int Ho, int Wo, int C, int c,
                                          int slot, int n_strips, int strips_x, int NS, const bool relu_in = false) {
  constexpr int R = dw_simt_rows(K, S), SX = kDwSX;
  float2 psum = make_float2(0.f, 0.f);
  for (int strip = slot; strip < n_strips; strip += NS) {
    const int sy = strip / strips_x, sx = strip - sy * strips_x;
    const int oy0 = sy * R, ox0 = sx * SX;
    float2 acc[R][SX];
    dw_patch<K, S, R, MODE>(tile + (oy0 * S) * rstride + (ox0 * S) * pstride, pstride, rstride, wv, acc, relu_in);
    // outputs: one bf16x2 word per pixel; whole patches (the common case) take the branch-free path
    uint8_t* op = reinterpret_cast<uint8_t*>(out + (((size_t)img * Ho + oy_t + oy0) * Wo + ox_t + ox0) * C + c);
    const size_t cbytes = (size_t)C * 2, rbytes = (size_t)Wo * cbytes;
    if (oy0 + R <= oy_end && ox0 + SX <= ox_end) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        uint8_t* p = op + r * rbytes;
#pragma unroll
        for (int j = 0; j < SX; ++j) {
          const float2 v = MODE == 1 ? acc[r][j] : silu2(acc[r][j], hsh);
          if (MODE == 0) psum = __fadd2_rn(psum, v);
          *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(v.x, v.y);
          p += cbytes;
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int j = 0; j < SX; ++j) {
          if (oy0 + r < oy_end && ox0 + j < ox_end) {
            const float2 v = MODE == 1 ? acc[r][j] : silu2(acc[r][j], hsh);
            if (MODE == 0) psum = __fadd2_rn(psum, v);
            *reinterpret_cast<uint32_t*>(op + r * rbytes + j * cbytes) = pack_bf16x2(v.x, v.y);
          }
        }
      }
    }
  }
  return psum;
}

template <int K, int S, int CWT, int MODE = 0>   // CWT: channels per chunk at compile time (0 = read g.CW)
__global__ void __launch_bounds__(256, 2)
dwconv_simt_kernel(const __grid_constant__ CUtensorMap tmap_in, const float* __restrict__ w,
                   const float* __restrict__ shift, bf16* __restrict__ out, float* __restrict__ pool_part, int n_img,
                   int Ho, int Wo, int C, int pad_lo, DwSimtGeom g, int relu_in) {
  extern __shared__ __align__(128) uint8_t dsm_raw[];
  uint8_t* ring = dsm_raw + ((128u - (ptx::smem_u32(dsm_raw) & 127u)) & 127u);   // stays a shared-space pointer
  float2* red = reinterpret_cast<float2*>(ring + 2 * g.tile_stride);   // [2][threads]
  __shared__ uint64_t bars[2];
  const int tid = threadIdx.x;
  const int CW = CWT ? CWT : g.CW, CP = CW / 2;
  const int cp = tid % CP, slot = tid / CP;
  const int cbase = blockIdx.y * CW;
  const int c = cbase + 2 * cp;
  const int n_work = n_img * g.tiles;
  const int my_steps = (n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  auto issue_load = [&](int step) {              // thread 0 only
    const int work = blockIdx.x + step * gridDim.x;
    const int img = work / g.tiles, t = work - img * g.tiles;
    const int tx = t % g.tiles_x, ty = t / g.tiles_x;
    uint64_t* bar = &bars[step & 1];
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)g.tile_bytes);
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(ptx::smem_u32(ring + (step & 1) * g.tile_stride)), "l"(reinterpret_cast<uint64_t>(&tmap_in)),
          "r"(ptx::smem_u32(bar)), "r"(cbase), "r"(tx * g.TW * S - pad_lo), "r"(ty * g.TH * S - pad_lo), "r"(img)
        : "memory");
  };
  if (tid == 0) {
    ptx::prefetch_tmap(&tmap_in);
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_mbar_init();
    if (my_steps > 0) issue_load(0);
  }
  float2 wv[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) wv[t] = *reinterpret_cast<const float2*>(w + (size_t)t * C + c);
  float2 hsh = make_float2(0.f, 0.f);
  if (MODE == 0) {
    hsh = *reinterpret_cast<const float2*>(shift + c);        // BN shift, pre-halved for silu2
    hsh.x *= 0.5f; hsh.y *= 0.5f;
  }
  const int pstride = CW * 2, rstride = g.IW * pstride;
  __syncthreads();                                // barriers initialised

  for (int step = 0; step < my_steps; ++step) {
    const int work = blockIdx.x + step * gridDim.x;
    const int img = work / g.tiles, t = work - img * g.tiles;
    const int tx = t % g.tiles_x, ty = t / g.tiles_x;
    // the other ring slot was released by the __syncthreads that ended step-1
    if (tid == 0 && step + 1 < my_steps) issue_load(step + 1);
    ptx::mbar_wait(&bars[step & 1], (step >> 1) & 1);
    const uint8_t* tile = ring + (step & 1) * g.tile_stride + cp * 4;
    const int oy_t = ty * g.TH, ox_t = tx * g.TW;
    const int oy_end = min(g.TH, Ho - oy_t), ox_end = min(g.TW, Wo - ox_t);   // valid outputs of this tile
    const float2 psum = dw_tile<K, S, MODE>(tile, pstride, rstride, wv, hsh, out, img, oy_t, ox_t, oy_end, ox_end, Ho, Wo, C, c,
                                            slot, g.n_strips, g.strips_x, g.NS, relu_in != 0);
    float2* rb = red + (step & 1) * g.threads;
    if (MODE == 0) rb[tid] = psum;
    __syncthreads();                              // tile consumed (ring slot free) + partial sums visible
    if (MODE == 0 && slot == 0) {
      float2 s = rb[cp];
      for (int sl = 1; sl < g.NS; ++sl) { const float2 v = rb[sl * CP + cp]; s.x += v.x; s.y += v.y; }
      *reinterpret_cast<float2*>(pool_part + ((size_t)img * g.tiles + t) * C + c) = s;   // one writer per entry
    }
  }
}

}  // namespace mt
