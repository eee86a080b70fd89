// Depthwise kxk convolution + BN + swish on tensor cores (bf16 path) -- reference model.py:105-107.
//
// A depthwise filter is a block-diagonal matrix: for one octet of channels and two filter taps (t0, t1)
//     out[p][c'] += sum_{(t,c)} in[p + t][c] * ( w[t][c] * [c == c'] ),     K index = (tap, channel) = 16
// which is exactly one mma.sync m16n8k16 (16 output pixels x 8 channels x (2 taps * 8 channels)).
// The A operand needs no im2col: ldmatrix takes one 16-byte row address per lane, and a row is the
// 8-channel vector of input pixel (p + t) in the NHWC tile, so the stencil gather (incl. stride 2) is
// just address arithmetic.  Only 1/8 of the MACs are useful, but the CUDA-core version needs ~25 FMA +
// ~25 bf16->fp32 conversions + loads per output and is issue-bound; here it is 13 (5x5) or 5 (3x3)
// ldmatrix+mma pairs per 128 outputs.
//
// Persistent blocks: a block owns one 64-channel chunk (its filter fragments stay in registers) and
// walks work items (image) x (spatial tiles of TH x TW outputs).  The (TH-1)*S+K by (TW-1)*S+K input
// tile of the NEXT step is fetched by a 4-D TMA load into the other half of a 2-deep shared-memory ring
// while the current one is computed (hardware zero fill implements the TF-"SAME" padding, the 128-byte
// swizzle makes the ldmatrix rows bank-conflict free).  Because a block sees every tile of an image for
// its channels, the squeeze-excite pool sum is accumulated in registers and written once per image.
#pragma once
#include <cuda.h>

#include "attention_mma.cuh"   // ldmatrix / mma.sync wrappers
#include "common.cuh"
#include "ptx.cuh"

namespace mt {

struct DwTcGeom {
  int TW, TH, IW, IH, tiles_x, tiles_y, n_cchunks, mtiles;
  int tile_bytes;
  int workers;      // persistent blocks per channel chunk
};

inline DwTcGeom dw_tc_geom(int H, int W, int C, int k, int s, int n_img = 1, int num_sms = 148) {
  DwTcGeom g;
  const int Ho = (H + s - 1) / s, Wo = (W + s - 1) / s;
  g.tiles_x = (Wo + 15) / 16;
  g.TW = (Wo + g.tiles_x - 1) / g.tiles_x;
  int th = std::max(1, 128 / g.TW);
  g.tiles_y = (Ho + th - 1) / th;
  g.TH = (Ho + g.tiles_y - 1) / g.tiles_y;
  auto in_dim = [&](int t) { return (t - 1) * s + k; };
  while (in_dim(g.TH) * in_dim(g.TW) * 128 > 44 * 1024 && g.TH > 1) {
    g.tiles_y = (Ho + (g.TH + 1) / 2 - 1) / ((g.TH + 1) / 2);
    g.TH = (Ho + g.tiles_y - 1) / g.tiles_y;
  }
  g.IW = in_dim(g.TW);
  g.IH = in_dim(g.TH);
  g.n_cchunks = (C + 63) / 64;
  g.mtiles = (g.TH * g.TW + 15) / 16;
  g.tile_bytes = g.IH * g.IW * 128;
  // resident blocks per SM: limited by the 2-deep tile ring (and 8 x 256 threads)
  const int ring = 2 * g.tile_bytes + 8 * 1024;
  const int per_sm = std::max(1, std::min(6, (200 * 1024) / ring));
  g.workers = std::max(1, std::min(n_img, (num_sms * per_sm + g.n_cchunks - 1) / g.n_cchunks));
  return g;
}

struct DwSeArgs {          // fused squeeze-excite tail (see effnet.cu)
  const float* wr; const float* br; const float* we_t; const float* be;
  float* gate; int* counters; int sq; float inv_hw;
};

template <int K, int S>
__global__ void __launch_bounds__(256) dwconv_tc_kernel(const __grid_constant__ CUtensorMap tmap_in,
                                                        const float* __restrict__ w, const float* __restrict__ shift,
                                                        bf16* __restrict__ out, float* __restrict__ pool, int n_img,
                                                        int Ho, int Wo, int C, int pad_lo, DwTcGeom g, DwSeArgs se) {
  constexpr int KK = K * K;
  constexpr int KS = (KK + 1) / 2;               // k-steps: two taps each
  extern __shared__ __align__(1024) uint8_t dsm_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~uintptr_t(1023));
  const int tile_stride = (g.tile_bytes + 1023) & ~1023;
  float* se_scratch = reinterpret_cast<float*>(ring + 2 * tile_stride);   // [C + SQ] floats (fused SE only)
  __shared__ uint64_t bars[2];
  __shared__ float red[8][8];                    // per warp: pool sums of its 8 channels
  __shared__ int is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cbase = blockIdx.y * 64;
  const int n_oc = min(8, (C - cbase) / 8);      // channel octets in this chunk
  const int tiles = g.tiles_x * g.tiles_y;
  // this block's sequence of steps: items (images) blockIdx.x, +gridDim.x, ... ; each item = `tiles` steps
  const int my_items = (n_img - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_steps = my_items * tiles;

  auto issue_load = [&](int step) {              // thread 0 only
    const int item = step / tiles, t = step - item * tiles;
    const int img = blockIdx.x + item * gridDim.x;
    const int tx = t % g.tiles_x, ty = t / g.tiles_x;
    uint64_t* bar = &bars[step & 1];
    ptx::mbar_arrive_expect_tx(bar, (uint32_t)g.tile_bytes);
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(ptx::smem_u32(ring + (step & 1) * tile_stride)), "l"(reinterpret_cast<uint64_t>(&tmap_in)),
          "r"(ptx::smem_u32(bar)), "r"(cbase), "r"(tx * g.TW * S - pad_lo), "r"(ty * g.TH * S - pad_lo), "r"(img)
        : "memory");
  };
  if (tid == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_mbar_init();
    if (total_steps > 0) issue_load(0);
  }
  // warp -> channel octet (and, when the chunk has fewer than 8 octets, a share of the m-tiles)
  const int wpo = n_oc >= 8 ? 1 : 8 / n_oc;       // warps per octet
  const int oct = warp % n_oc, msub = warp / n_oc;
  const bool warp_active = msub < wpo;
  const int gq = lane >> 2, tq = lane & 3;
  const int c0 = cbase + oct * 8;
  // block-diagonal B fragments: b0 <- tap 2ks, b1 <- tap 2ks+1; lane holds k = (2tq, 2tq+1), n = gq
  uint32_t wb[KS][2];
  float sh0 = 0.f, sh1 = 0.f;
  if (warp_active) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = 2 * ks + h;
        float w0 = 0.f, w1 = 0.f;
        if (t < KK) {
          if (gq == 2 * tq) w0 = w[(size_t)t * C + c0 + gq];
          if (gq == 2 * tq + 1) w1 = w[(size_t)t * C + c0 + gq];
        }
        wb[ks][h] = attn::pack2(w0, w1);
      }
    }
    sh0 = shift[c0 + 2 * tq];
    sh1 = shift[c0 + 2 * tq + 1];
  }
  // per-lane constants of the ldmatrix gather, hoisted out of all loops
  const int prow = (lane & 7) + ((lane >> 3) & 1) * 8;   // row of the m-tile this lane addresses
  const int thalf = lane >> 4;                           // which of the two taps of a k-step
  int toff[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int ta = 2 * ks, tb = (2 * ks + 1 < KK) ? 2 * ks + 1 : KK - 1;   // odd tap count: partner B rows are zero
    toff[ks] = thalf ? (tb / K) * g.IW + (tb % K) : (ta / K) * g.IW + (ta % K);
  }
  const uint32_t inv_tw = (65536u + g.TW - 1) / g.TW;    // q / TW == (q * inv_tw) >> 16 for q < 4096
  const int n_out = g.TH * g.TW;
  __syncthreads();                                       // barriers initialised

  float ps0 = 0.f, ps1 = 0.f;                            // pool sums of channels (2tq, 2tq+1) over the current image
  for (int step = 0; step < total_steps; ++step) {
    const int item = step / tiles, t = step - item * tiles;
    const int img = blockIdx.x + item * gridDim.x;
    const int tx = t % g.tiles_x, ty = t / g.tiles_x;
    // prefetch the next step's tile into the other ring slot (its previous contents were consumed before
    // the __syncthreads that ended step-1)
    if (tid == 0 && step + 1 < total_steps) issue_load(step + 1);
    ptx::mbar_wait(&bars[step & 1], (step >> 1) & 1);
    if (warp_active) {
      const uint32_t tile_addr = ptx::smem_u32(ring + (step & 1) * tile_stride);
      const int oy_base = ty * g.TH, ox_base = tx * g.TW;
      for (int mt = msub; mt < g.mtiles; mt += wpo) {
        int p = mt * 16 + prow;
        if (p >= n_out) p = n_out - 1;                   // clamp (result discarded)
        const int py = (int)(((uint32_t)p * inv_tw) >> 16), px = p - py * g.TW;
        const int ip0 = py * S * g.IW + px * S;          // input-tile pixel of tap (0,0)
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const int ip = ip0 + toff[ks];
          uint32_t a[4];
          attn::ldmatrix_x4(a, tile_addr + (ip << 7) + (((ip & 7) ^ oct) << 4));
          attn::mma_bf16(acc, a, wb[ks][0], wb[ks][1]);
        }
        // epilogue: rows gq and gq+8 of the m-tile, channels c0 + 2tq, +1
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int q = mt * 16 + gq + r * 8;
          if (q < n_out) {
            const int qy = (int)(((uint32_t)q * inv_tw) >> 16), qx = q - qy * g.TW;
            const int oy = oy_base + qy, ox = ox_base + qx;
            if (oy < Ho && ox < Wo) {
              const float v0 = silu<false>(acc[r * 2] + sh0), v1 = silu<false>(acc[r * 2 + 1] + sh1);
              ps0 += v0; ps1 += v1;
              *reinterpret_cast<uint32_t*>(out + (((size_t)img * Ho + oy) * Wo + ox) * C + c0 + 2 * tq) =
                  attn::pack2(v0, v1);
            }
          }
        }
      }
    }
    if (t + 1 < tiles) {
      __syncthreads();                                   // ring slot free for the load issued next step
      continue;
    }
    // ---- last tile of this image: pool sums (fixed reduction order) and the squeeze-excite tail
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      ps0 += __shfl_xor_sync(0xffffffffu, ps0, o);
      ps1 += __shfl_xor_sync(0xffffffffu, ps1, o);
    }
    if (lane < 4) { red[warp][2 * lane] = ps0; red[warp][2 * lane + 1] = ps1; }
    ps0 = ps1 = 0.f;
    __syncthreads();                                     // also frees the ring slot
    if (tid < n_oc * 8) {
      const int o = tid >> 3, ch = tid & 7;
      float s = 0.f;
      for (int ws = 0; ws < wpo; ++ws) s += red[ws * n_oc + o][ch];
      pool[(size_t)img * C + cbase + tid] = s;           // one writer per (image, channel)
    }
    if (se.wr == nullptr) continue;                      // (block-uniform)

    // fused squeeze-excite (model.py:110-115): the last channel chunk to finish this image computes its gate
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(se.counters + img, 1) == (int)gridDim.y - 1;
    __syncthreads();
    if (is_last) {
      __threadfence();
      if (tid == 0) se.counters[img] = 0;
      float* mean = se_scratch;
      float* sqv = se_scratch + C;
      for (int c = tid; c < C; c += 256) mean[c] = __ldcg(pool + (size_t)img * C + c) * se.inv_hw;
      __syncthreads();
      for (int j = warp; j < se.sq; j += 8) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(se.wr[(size_t)j * C + c], mean[c], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sqv[j] = silu<true>(s + se.br[j]);
      }
      __syncthreads();
      for (int c = tid; c < C; c += 256) {
        float s = se.be[c];
        for (int j = 0; j < se.sq; ++j) s = fmaf(se.we_t[(size_t)j * C + c], sqv[j], s);
        se.gate[(size_t)img * C + c] = sigmoidf_<true>(s);
      }
      __syncthreads();                                   // se_scratch / is_last reusable
    }
  }
}

}  // namespace mt
