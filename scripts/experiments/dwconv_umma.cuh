// Depthwise kxk (stride 1) + BN + swish + SE pool sums on the 5th-gen tensor cores (tcgen05) -- bf16 path.
//
// Legacy mma.sync runs at ~1 HMMA.16816 / 32 cycles / SMSP on B200 (dwconv_tc.cuh is bound by it), so the
// depthwise stencil is mapped onto tcgen05.mma instead.  For one filter tap t and one group of 16 channels
//     D[128 positions][16 ch] += A_t[128 positions][16 ch] * diag(w[t][16 ch])          (M=128, N=16, K=16)
// where A_t is simply the NHWC input tile in shared memory (one 128-byte swizzled row per pixel, written by a
// 4-D TMA load whose zero fill implements the TF-"SAME" padding) addressed from row (position + tap offset):
// consecutive output positions of a full-width tile are consecutive tile pixels, so every tap is the same
// K-major SWIZZLE_128B operand with a shifted start address (no descriptor base_offset: measured).
// The diagonal B tiles (2 KiB per tap, all 4 channel groups) are built once per persistent block.
// 15/16 of the MACs multiply zeros; the tensor pipe has 30x the FMA pipe's rate and, unlike the CUDA-core
// formulation, needs no bf16->fp32 conversions or address math per tap.  Positions in the K-1 halo columns of
// each row are computed and dropped.  SMEM operand reads (4.5 KiB per MMA) bound the kernel.
//
//   warp 0   TMA producer (2-deep tile ring)         warp 1   MMA issuer, TMEM owner (2 x 64 columns)
//   warps 2-5 epilogue: tcgen05.ld -> +shift, swish -> 128-byte row stores, pool sums
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mt {

struct DwUmmaGeom {
  int TH, IW, IH, tiles_y, P, mtiles, n_cchunks, workers;
  int slot_bytes;      // one ring slot (tile + over-read slack), multiple of 1024
  int b_bytes;         // diagonal filter tiles
  int use_base_offset; // debug switch: descriptor base_offset for row-shifted A starts
};

inline DwUmmaGeom dw_umma_geom(int H, int W, int C, int k, int n_img, int num_sms) {
  DwUmmaGeom g;
  g.IW = W + k - 1;
  // rows per tile: as many as fit ~64 KiB, balanced over the image
  int th = std::max(1, (64 * 1024 / 128) / g.IW - (k - 1));
  th = std::min(th, H);
  g.tiles_y = (H + th - 1) / th;
  g.TH = (H + g.tiles_y - 1) / g.tiles_y;
  g.IH = g.TH + k - 1;
  g.P = (g.TH - 1) * g.IW + W;                     // linear positions holding valid outputs
  g.mtiles = (g.P + 127) / 128;
  const int max_row = g.mtiles * 128 + (k - 1) * g.IW + (k - 1);   // highest tile row an MMA may touch (+1)
  g.slot_bytes = ((std::max(max_row, g.IH * g.IW) * 128) + 1023) & ~1023;
  g.b_bytes = k * k * 2048;
  g.n_cchunks = (C + 63) / 64;
  g.workers = std::max(1, std::min(n_img, (num_sms + g.n_cchunks - 1) / g.n_cchunks));
  g.use_base_offset = 0;   // measured on B200: the 128B swizzle is applied on absolute smem address bits, so a
                           // row-shifted start needs NO descriptor base_offset (setting it scrambles the rows)
  return g;
}

// SWIZZLE_128B K-major descriptor whose start may sit on any 128-byte row of a 1024-byte-aligned tile
__device__ __forceinline__ uint64_t umma_desc_sw128_row(uint32_t smem_addr, int use_bo) {
  uint64_t d = ptx::umma_smem_desc_sw128(smem_addr);
  if (use_bo) d |= static_cast<uint64_t>((smem_addr >> 7) & 7) << 49;      // matrix base offset
  return d;
}

template <int K>
__global__ void __launch_bounds__(192, 1)
dwconv_umma_kernel(const __grid_constant__ CUtensorMap tmap_in, const float* __restrict__ w,
                   const float* __restrict__ shift, bf16* __restrict__ out, float* __restrict__ pool, int n_img, int H,
                   int W, int C, int pad_lo, DwUmmaGeom g) {
  constexpr int KK = K * K;
  extern __shared__ __align__(1024) uint8_t dsm_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = base;                                   // 2 slots
  uint8_t* bmat = base + 2 * g.slot_bytes;                // KK x 2 KiB
  __shared__ uint64_t full_bar[2], empty_bar[2], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_ptr_smem;
  __shared__ float red[4][64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cbase = blockIdx.y * 64;
  const int n_ch = min(64, C - cbase);                    // channels in this chunk (multiple of 8)
  const int n_grp = (n_ch + 15) / 16;                     // 16-channel MMA groups
  const int steps_per_img = g.tiles_y;
  const int my_items = (n_img - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_steps = my_items * steps_per_img;

  if (tid == 0) {
    ptx::prefetch_tmap(&tmap_in);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_ptr_smem, 128);
    ptx::tmem_relinquish();
  }
  // diagonal filter tiles: B_t[n][kk] = w[t][cbase + kk] if kk % 16 == n else 0  (n < 16, kk < 64),
  // stored K-major with the 128-byte swizzle (row n: 16-byte chunk c at (c ^ (n & 7)))
  for (int i = tid; i < g.b_bytes / 16; i += 192) reinterpret_cast<uint4*>(bmat)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = tid; i < KK * 64; i += 192) {
    const int t = i >> 6, kk = i & 63, n = kk & 15;
    if (kk < n_ch) {
      const int c = kk >> 3;
      uint8_t* p = bmat + t * 2048 + (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4) + (kk & 7) * 2;
      *reinterpret_cast<bf16*>(p) = __float2bfloat16_rn(w[(size_t)t * C + cbase + kk]);
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_ptr_smem;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      for (int step = 0; step < total_steps; ++step) {
        const int slot = step & 1;
        const int item = step / steps_per_img, ty = step - item * steps_per_img;
        const int img = blockIdx.x + item * gridDim.x;
        ptx::mbar_wait(&empty_bar[slot], ((step >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(g.IH * g.IW * 128));
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(ptx::smem_u32(ring + slot * g.slot_bytes)), "l"(reinterpret_cast<uint64_t>(&tmap_in)),
              "r"(ptx::smem_u32(&full_bar[slot])), "r"(cbase), "r"(-pad_lo), "r"(ty * g.TH - pad_lo), "r"(img)
            : "memory");
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(128, 16);
      const uint32_t b_addr = ptx::smem_u32(bmat);
      int it = 0;                                         // running M-tile counter (TMEM double buffer)
      for (int step = 0; step < total_steps; ++step) {
        const int slot = step & 1;
        ptx::mbar_wait(&full_bar[slot], (step >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t a_base = ptx::smem_u32(ring + slot * g.slot_bytes);
        for (int mt = 0; mt < g.mtiles; ++mt, ++it) {
          const int as = it & 1;
          ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(as * 64);
#pragma unroll 1
          for (int t = 0; t < KK; ++t) {
            const uint32_t a_row = a_base + (uint32_t)(mt * 128 + (t / K) * g.IW + (t % K)) * 128u;
            for (int gi = 0; gi < n_grp; ++gi)
              ptx::umma_bf16_ss(tmem_d + gi * 16, umma_desc_sw128_row(a_row + gi * 32, g.use_base_offset),
                                ptx::umma_smem_desc_sw128(b_addr + t * 2048 + gi * 32), idesc, t != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&tmem_full[as]);
        }
        ptx::umma_commit(&empty_bar[slot]);               // tile consumed: the producer may refill the slot
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue (4 warps)
    const int quad = warp & 3;
    const int trow = quad * 32 + lane;
    const uint32_t inv_iw = (65536u + g.IW - 1) / g.IW;   // p / IW == (p * inv_iw) >> 16 for p < 65536 / IW ...
    float psum[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) psum[i] = 0.f;
    float sh[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) sh[i] = (i < n_ch) ? shift[cbase + i] : 0.f;
    int it = 0;
    for (int step = 0; step < total_steps; ++step) {
      const int item = step / steps_per_img, ty = step - item * steps_per_img;
      const int img = blockIdx.x + item * gridDim.x;
      for (int mt = 0; mt < g.mtiles; ++mt, ++it) {
        const int as = it & 1;
        ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 64);
        const int p = mt * 128 + trow;
        const int y = (int)(((uint32_t)p * inv_iw) >> 16), x = p - y * g.IW;
        const int oy = ty * g.TH + y;
        const bool valid = p < g.P && x < W && y < g.TH && oy < H;
        bf16* orow = out + (((size_t)img * H + oy) * W + x) * C + cbase;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q * 16 < n_ch) {                            // (block-uniform)
            uint32_t r[16];
            ptx::tmem_ld_32x32b_x16(taddr + q * 16, r);
            ptx::tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                if (q * 16 + hh * 8 < n_ch) {
                  float v[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    v[i] = silu<false>(__uint_as_float(r[hh * 8 + i]) + sh[q * 16 + hh * 8 + i]);
                    psum[q * 16 + hh * 8 + i] += v[i];
                  }
                  store8(orow + q * 16 + hh * 8, v);
                }
              }
            }
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tmem_empty[as]);
      }
      if (ty + 1 < steps_per_img) continue;
      // ---- image finished: pool sums of this block's 64 channels (fixed order: lanes, then warps)
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float s = psum[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[quad][i] = s;
        psum[i] = 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (trow < n_ch) pool[(size_t)img * C + cbase + trow] = (red[0][trow] + red[1][trow]) + (red[2][trow] + red[3][trow]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace mt
