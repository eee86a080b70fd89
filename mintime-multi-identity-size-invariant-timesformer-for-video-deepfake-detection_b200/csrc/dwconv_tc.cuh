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
// One block = (image, spatial tile of TH x TW outputs, 64-channel chunk): the (TH-1)*S+K by (TW-1)*S+K
// input tile is fetched with ONE 4-D TMA load (hardware zero fill implements the TF-"SAME" padding,
// 128-byte swizzle makes the ldmatrix rows bank-conflict free), 8 warps = 8 channel octets.
#pragma once
#include <cuda.h>

#include "attention_mma.cuh"   // ldmatrix / mma.sync wrappers
#include "common.cuh"
#include "ptx.cuh"

namespace mt {

struct DwTcGeom {
  int TW, TH, IW, IH, tiles_x, tiles_y, n_cchunks, mtiles;
  int tile_bytes;
};

inline DwTcGeom dw_tc_geom(int H, int W, int C, int k, int s) {
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
  return g;
}

struct DwSeArgs {          // fused squeeze-excite tail (see effnet.cu)
  const float* wr; const float* br; const float* we_t; const float* be;
  float* gate; int* counters; int sq; float inv_hw;
};

template <int K, int S>
__global__ void __launch_bounds__(256) dwconv_tc_kernel(const __grid_constant__ CUtensorMap tmap_in,
                                                        const float* __restrict__ w, const float* __restrict__ shift,
                                                        bf16* __restrict__ out, float* __restrict__ pool_part, int Ho,
                                                        int Wo, int C, int pad_lo, DwTcGeom g, DwSeArgs se) {
  constexpr int KK = K * K;
  constexpr int KS = (KK + 1) / 2;               // k-steps: two taps each
  extern __shared__ __align__(1024) uint8_t dsm_raw[];
  uint8_t* tile = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ float red[8][8];                    // per warp: pool sums of its 8 channels
  __shared__ int is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = blockIdx.x % g.tiles_x, ty = blockIdx.x / g.tiles_x;
  const int cbase = blockIdx.y * 64;
  const int img = blockIdx.z;
  const int n_oc = min(8, (C - cbase) / 8);      // channel octets in this chunk
  if (tid == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_mbar_init();
    ptx::mbar_arrive_expect_tx(&bar, (uint32_t)g.tile_bytes);
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(ptx::smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tmap_in)), "r"(ptx::smem_u32(&bar)), "r"(cbase),
          "r"(tx * g.TW * S - pad_lo), "r"(ty * g.TH * S - pad_lo), "r"(img)
        : "memory");
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
  __syncthreads();                               // mbarrier init visible
  ptx::mbar_wait(&bar, 0);

  float ps0 = 0.f, ps1 = 0.f;                    // pool partials of channels (2tq, 2tq+1)
  const int oy_base = ty * g.TH, ox_base = tx * g.TW;
  if (warp_active) {
    const uint32_t tile_addr = ptx::smem_u32(tile);
    // ldmatrix row of this lane: output pixel (lane & 7) + 8 * ((lane >> 3) & 1) of the m-tile, tap parity lane >> 4
    const int prow = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int thalf = lane >> 4;
    // per-lane tap offsets inside the input tile, hoisted out of the m-tile loop: the inner loop is then
    // add + swizzle + ldmatrix + mma (integer address math was 80 % of the issued instructions before)
    int toff[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int ta = 2 * ks, tb = (2 * ks + 1 < KK) ? 2 * ks + 1 : KK - 1;   // odd tap count: partner B rows are zero
      toff[ks] = thalf ? (tb / K) * g.IW + (tb % K) : (ta / K) * g.IW + (ta % K);
    }
    const uint32_t inv_tw = (65536u + g.TW - 1) / g.TW;   // q / TW == (q * inv_tw) >> 16 for q < 4096
    const int n_out = g.TH * g.TW;
    for (int mt = msub; mt < g.mtiles; mt += wpo) {
      int p = mt * 16 + prow;
      if (p >= n_out) p = n_out - 1;               // clamp (result discarded)
      const int py = (int)(((uint32_t)p * inv_tw) >> 16), px = p - py * g.TW;
      const int ip0 = py * S * g.IW + px * S;      // input-tile pixel of tap (0,0)
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
  // ---- pool partial of this block: reduce over the 8 pixel rows of the fragment, then over warps
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    ps0 += __shfl_xor_sync(0xffffffffu, ps0, o);
    ps1 += __shfl_xor_sync(0xffffffffu, ps1, o);
  }
  if (lane < 4) { red[warp][2 * lane] = ps0; red[warp][2 * lane + 1] = ps1; }
  __syncthreads();
  const int n_tiles = gridDim.x;
  if (tid < n_oc * 8) {
    const int o = tid >> 3, ch = tid & 7;
    float s = 0.f;
    for (int ws = 0; ws < wpo; ++ws) s += red[ws * n_oc + o][ch];     // fixed order
    pool_part[((size_t)img * n_tiles + blockIdx.x) * C + cbase + tid] = s;
  }
  if (se.wr == nullptr) return;

  // ---- fused squeeze-excite (model.py:110-115) by the last block of this image
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = atomicAdd(se.counters + img, 1) == (int)(gridDim.x * gridDim.y) - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (tid == 0) se.counters[img] = 0;
  float* mean = reinterpret_cast<float*>(tile);    // the input tile is dead: reuse it ([C] + [SQ] floats)
  float* sqv = mean + C;
  for (int c = tid; c < C; c += 256) {
    double a = 0.0;
    for (int j = 0; j < n_tiles; ++j) a += (double)__ldcg(pool_part + ((size_t)img * n_tiles + j) * C + c);
    mean[c] = (float)(a * (double)se.inv_hw);
  }
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
}

}  // namespace mt
