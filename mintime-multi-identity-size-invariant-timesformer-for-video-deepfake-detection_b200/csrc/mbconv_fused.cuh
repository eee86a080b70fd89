// Front half of an MBConv block in ONE kernel (reference model.py:98-107):
//     1x1 expand conv + BN + swish  ->  depthwise kxk + BN + swish  (+ squeeze-excite pool partial sums)
// The expanded activation (6x the block input, the largest tensor of the network) never goes to HBM:
//   * a persistent block owns one chunk of CW expanded channels; its slice of the expand weights sits in
//     shared memory and its depthwise taps in registers for the block's whole life;
//   * per tile, ONE 4-D TMA load brings the (TH-1)*S+K by (TW-1)*S+K input pixels x Cin channels (hardware zero
//     fill outside the image; 32/64/128-byte swizzled rows = a K-major UMMA operand, no im2col).  (Staging the
//     tile with cp.async instead was measured 25 % slower: ~50 address instructions per 16-byte piece.)
//   * tcgen05.mma contracts them with the weight slice: 128 pixels x CW channels per instruction, fp32
//     accumulators for the whole tile in TMEM (ceil(pixels/128) * CW columns);
//   * all warps drain TMEM (tcgen05.ld), add the BN shift, apply the swish, ZERO the pixels that lie in the
//     TF-"SAME" padding of the depthwise conv (utils.py:254-269 pads the *expanded* tensor) and write the bf16
//     tile to shared memory with a 16-byte padded pixel pitch (conflict-free for both the row-per-lane writes
//     here and the channel-per-lane reads of the stencil);
//   * the depthwise stencil then runs out of shared memory exactly as in dwconv_simt.cuh (dw_tile);
//   * the three stages run concurrently on different tiles (warp-specialised, see the kernel).
#pragma once
#include <cuda.h>

#include "dwconv_simt.cuh"

namespace mt {

struct FusedGeom {
  DwSimtGeom d;            // tile / thread geometry of the depthwise part (tile_bytes unused)
  int kbox, row_bytes;     // K extent (elements) and bytes of one operand row in smem: 16/32, 32/64, 64/128
  int ksteps;              // K = 16 MMAs per 128-pixel block
  int n_pix, MB;           // input-tile pixels, 128-pixel blocks
  int a_stride;            // bytes per input-tile buffer (MB * 128 * row_bytes)
  int w_stride;            // bytes of the weight slice (CW * row_bytes, 1024-aligned)
  int pitch, exp_bytes;    // expanded-tile pixel pitch (CW*2 + 16) and size
  int tmem_cols;
  int nd;                  // drain warps (4 or 8); the stencil gets the remaining warps of the 512-thread block
  size_t smem;
};

inline bool fused_geom(FusedGeom* out, int H, int cin, int C, int k, int s, int n_img, int num_sms) {
  if (cin > 64 || cin % 8 != 0) return false;
  const int R = dw_simt_rows(k, s), SX = kDwSX;
  const int W = H, Ho = (H + s - 1) / s, Wo = Ho;
  const int pad = std::max((Ho - 1) * s + k - H, 0) / 2;
  const int kbox = cin <= 16 ? 16 : (cin <= 32 ? 32 : 64), row_bytes = kbox * 2;
  auto cdiv = [](int a, int b) { return (a + b - 1) / b; };
  double best = 1e30;
  FusedGeom g{};
  bool found = false;
  for (int CW = 32; CW <= 64; CW += 16) {
    if (C % CW) continue;
    const int CP = CW / 2, pitch = CW * 2 + 16;
    const int w_stride = (CW * row_bytes + 1023) & ~1023;
    for (int nx = 1; nx <= 8; ++nx) {
      int TW = cdiv(Wo, nx);
      if (TW >= SX) TW = cdiv(TW, SX) * SX;
      if (nx > 1 && cdiv(Wo, TW) != nx) continue;
      for (int ny = 1; ny <= Ho; ++ny) {
        int TH = cdiv(cdiv(Ho, ny), R) * R;
        if (ny > 1 && cdiv(Ho, TH) != ny) continue;
        const int IH = (TH - 1) * s + k, IW = (TW - 1) * s + k;
        if (IW > 256 || IH > 256) continue;
        const int n_pix = IH * IW, MB = cdiv(n_pix, 128);
        if (MB * CW > 256) continue;               // two accumulator buffers in the SM's 512 TMEM columns
        const int a_stride = MB * 128 * row_bytes, exp_bytes = (n_pix * pitch + 127) & ~127;
        const int tiles_x = cdiv(Wo, TW), tiles_y = cdiv(Ho, TH);
        const int n_strips = cdiv(TH, R) * cdiv(TW, SX);
        double halo = 0;
        for (int ty = 0; ty < tiles_y; ++ty)
          for (int tx = 0; tx < tiles_x; ++tx) {
            const int y0 = ty * TH * s - pad, x0 = tx * TW * s - pad;
            halo += (double)(std::min(H, y0 + IH) - std::max(0, y0)) * (std::min(W, x0 + IW) - std::max(0, x0));
          }
        halo /= (double)H * W;
        // Warp split: drain-heavy shapes (K = 16: the expanded tile is 6x the input and every element costs a tanh)
        // get 8 drain warps -- one per SM sub-partition leaves the MUFU pipe half idle (measured 671 -> 503 us on
        // block 1) -- the others 4 drain + up to 11 stencil warps (block 2: 395 -> 384 us).
        const int nd = cin <= 16 ? 8 : 4;
        for (int NS = 1; NS <= (480 - 32 * nd) / CP; ++NS) {
          const int thr = CP * NS;
          if (thr < 128) continue;
          const size_t smem = 1024 + 2 * (size_t)a_stride + w_stride + 2 * (size_t)exp_bytes + 2 * (size_t)((thr + 31) & ~31) * 8 +
                              CW * 4 + 64;
          if (smem > 200 * 1024) continue;
          const int passes = cdiv(n_strips, NS);
          const double util = (double)Ho * Wo / ((double)passes * NS * R * SX * tiles_x * tiles_y);
          // the expand epilogue is recomputed on halo pixels: weigh the halo more than the pure stencil does
          const double score = 1.0 / util + 0.8 * (halo - 1.0) + (thr < 192 ? 0.1 : 0.0) - 1e-4 * CW +
                               0.15 * (MB * 128.0 / n_pix - 1.0);
          if (score < best) {
            best = score;
            found = true;
            DwSimtGeom& d = g.d;
            d.CW = CW; d.CP = CP; d.n_cchunks = C / CW; d.NS = NS; d.threads = thr;
            d.TH = TH; d.TW = TW; d.IH = IH; d.IW = IW;
            d.tiles_x = tiles_x; d.tiles_y = tiles_y; d.tiles = tiles_x * tiles_y;
            d.strips_x = cdiv(TW, SX); d.n_strips = n_strips;
            d.tile_bytes = n_pix * row_bytes; d.tile_stride = a_stride;
            g.kbox = kbox; g.row_bytes = row_bytes; g.ksteps = cdiv(cin, 16);
            g.n_pix = n_pix; g.MB = MB; g.a_stride = a_stride; g.w_stride = w_stride;
            g.pitch = pitch; g.exp_bytes = exp_bytes;
            g.nd = nd;
            g.smem = smem;
          }
        }
      }
    }
  }
  if (!found) return false;
  int tm = 32;
  while (tm < 2 * g.MB * g.d.CW) tm <<= 1;
  g.tmem_cols = tm;
  g.d.smem = g.smem;
  g.d.workers = 1;
  (void)n_img; (void)num_sms;
  *out = g;
  return true;
}

// K-major operand rows of 32 / 64 / 128 bytes written by TMA with the matching swizzle: 8-row atoms, so the
// stride between 8-row groups is 8 * row_bytes; layout field 6 / 4 / 2 (SWIZZLE_32B / 64B / 128B); version 1.
__device__ __forceinline__ uint64_t umma_smem_desc_rows(uint32_t smem_addr, int row_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((8 * row_bytes) >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6)) << 61;
  return d;
}

// Warp roles (one block per SM, mbarrier hand-offs only -- no block-wide barrier in the steady state):
//   warp 0        (one thread) TMA producer of the input tiles + tcgen05.mma issuer
//   warps 1..nd   drain: TMEM -> BN shift + swish (MUFU) -> padding mask -> bf16 tile in shared memory;
//                 warp w reads the TMEM lane quadrant w % 4; nd = 8 (the two warps of a quadrant alternate 128-pixel
//                 blocks) for drain-heavy shapes -- one drain warp per SM sub-partition leaves the MUFU pipe half
//                 idle -- nd = 4 for stencil-heavy ones (fused_geom's cost model decides)
//   warps nd+1..  stencil: dw_tile on the finished tile (FFMA2), output + pool partial sums
// Input tiles, TMEM accumulators and expanded tiles are all double buffered, so the MUFU-bound drain of tile
// i+1 overlaps the FMA-bound stencil of tile i on the same SM.

template <int K, int S, int CW>
__global__ void __launch_bounds__(512, 1)
mbconv_front_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_w,
                    const float* __restrict__ exp_shift, const float* __restrict__ w_dw,
                    const float* __restrict__ dw_shift, bf16* __restrict__ out, float* __restrict__ pool_part, int n_img,
                    int H, int W, int Ho, int Wo, int C, int pad_lo, FusedGeom g) {
  constexpr int CP = CW / 2, PSTRIDE = CW * 2 + 16;
  extern __shared__ __align__(1024) uint8_t fsm_raw[];
  uint8_t* sm = fsm_raw + ((1024u - (ptx::smem_u32(fsm_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  uint8_t* a_buf = sm;                                   // [2][a_stride]   input tiles (UMMA A operand)
  uint8_t* w_buf = a_buf + 2 * g.a_stride;               // [CW][row_bytes] expand weights (UMMA B operand)
  uint8_t* exp_buf = w_buf + g.w_stride;                 // [2][exp_bytes]  expanded activations, bf16, pitch PSTRIDE
  const int nd = g.nd;                                   // drain warps 1..nd
  const int dw0 = 32 + 32 * nd;                          // first stencil thread
  const int n_dw = g.d.threads;                          // CP * NS stencil threads (the launch rounds up to a warp)
  const int n_dw_pad = (n_dw + 31) & ~31;
  float2* red = reinterpret_cast<float2*>(exp_buf + 2 * g.exp_bytes);   // [2][n_dw_pad]
  float* hsh_exp = reinterpret_cast<float*>(red + 2 * n_dw_pad);        // [CW] expand BN shift / 2
  __shared__ uint64_t a_full[2], a_empty[2], t_full[2], t_empty[2], e_full[2], e_empty[2], bar_w;
  __shared__ uint32_t tmem_ptr_smem;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cbase = blockIdx.y * CW;
  const int tiles = g.d.tiles;
  const int n_work = n_img * tiles;
  const int my_steps = (n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int IW = g.d.IW;
  const int acc_cols = g.MB * CW;                        // TMEM columns of one accumulator buffer
  const uint32_t iw_magic = (1u << 20) / (uint32_t)IW + 1u;   // row / IW == (row * iw_magic) >> 20 for row < 4096

  auto tile_origin = [&](int step, int& img, int& t, int& x0, int& y0) {
    const int work = blockIdx.x + step * gridDim.x;
    img = work / tiles; t = work - img * tiles;
    x0 = (t % g.d.tiles_x) * g.d.TW * S - pad_lo;
    y0 = (t / g.d.tiles_x) * g.d.TH * S - pad_lo;
  };

  if (tid == 0) {
    ptx::prefetch_tmap(&tmap_in);
    ptx::prefetch_tmap(&tmap_w);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&a_full[i], 1);
      ptx::mbar_init(&a_empty[i], 1);
      ptx::mbar_init(&t_full[i], 1);
      ptx::mbar_init(&t_empty[i], (uint32_t)nd);         // one arrival per drain warp
      ptx::mbar_init(&e_full[i], (uint32_t)nd);
      ptx::mbar_init(&e_empty[i], (uint32_t)(n_dw_pad >> 5));   // one arrival per stencil warp
    }
    ptx::mbar_init(&bar_w, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_ptr_smem, (uint32_t)g.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (tid < CW) hsh_exp[tid] = 0.5f * exp_shift[cbase + tid];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_ptr_smem;

  if (warp == 0) {
    // ===================================================================== TMA producer + MMA issuer
    if (lane == 0 && my_steps > 0) {
      auto issue_a = [&](int step) {
        int img, t, x0, y0;
        tile_origin(step, img, t, x0, y0);
        uint64_t* bar = &a_full[step & 1];
        ptx::mbar_arrive_expect_tx(bar, (uint32_t)(g.n_pix * g.row_bytes));
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(ptx::smem_u32(a_buf + (step & 1) * g.a_stride)), "l"(reinterpret_cast<uint64_t>(&tmap_in)),
              "r"(ptx::smem_u32(bar)), "r"(0), "r"(x0), "r"(y0), "r"(img)
            : "memory");
      };
      ptx::mbar_arrive_expect_tx(&bar_w, (uint32_t)(CW * g.row_bytes));
      ptx::tma_load_2d(w_buf, &tmap_w, &bar_w, 0, cbase);
      issue_a(0);
      if (my_steps > 1) issue_a(1);
      ptx::mbar_wait(&bar_w, 0);
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(128, CW);
      const uint32_t b0 = ptx::smem_u32(w_buf);
      for (int step = 0; step < my_steps; ++step) {
        const int s = step & 1;
        const uint32_t ph = (step >> 1) & 1;
        ptx::mbar_wait(&a_full[s], ph);                  // input tile landed
        ptx::mbar_wait(&t_empty[s], ph ^ 1);             // accumulator buffer drained (passes on first use)
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_u32(a_buf + s * g.a_stride);
        for (int mb = 0; mb < g.MB; ++mb)
          for (int k = 0; k < g.ksteps; ++k)
            ptx::umma_bf16_ss(tmem_base + (uint32_t)(s * acc_cols + mb * CW),
                              umma_smem_desc_rows(a0 + mb * 128 * g.row_bytes + k * 32, g.row_bytes),
                              umma_smem_desc_rows(b0 + k * 32, g.row_bytes), idesc, k != 0 ? 1u : 0u);
        ptx::umma_commit(&t_full[s]);                    // accumulators ready for the drain warps
        ptx::umma_commit(&a_empty[s]);                   // input buffer reusable once these MMAs retire
        if (step + 2 < my_steps) {
          ptx::mbar_wait(&a_empty[s], ph);
          issue_a(step + 2);
        }
      }
    }
    __syncwarp();
  } else if (warp <= nd) {
    // ===================================================================== drain warps (TMEM -> swish -> smem)
    const int quad = warp & 3, half = (warp - 1) >> 2;
    for (int step = 0; step < my_steps; ++step) {
      const int s = step & 1;
      const uint32_t ph = (step >> 1) & 1;
      int img, t, x0, y0;
      tile_origin(step, img, t, x0, y0);
      ptx::mbar_wait(&e_empty[s], ph ^ 1);               // the stencil warps are done with this tile buffer
      ptx::mbar_wait(&t_full[s], ph);
      ptx::tc_fence_after();
      uint8_t* exp_tile = exp_buf + s * g.exp_bytes;
      // Software pipeline over this warp's 128-pixel blocks: the tcgen05.ld of the NEXT block is in flight while the
      // current one goes through the swish (one block = 32 elements per lane = 256 MUFU cycles of the sub-partition;
      // an exposed TMEM round trip per block cost about as much again).
      const int mstep = nd >> 2;
      const uint32_t tquad = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(s * acc_cols);
      auto ld_block = [&](uint32_t (&r)[CW / 16][16], int mb) {
#pragma unroll
        for (int c16 = 0; c16 < CW / 16; ++c16) ptx::tmem_ld_32x32b_x16(tquad + (uint32_t)(mb * CW + c16 * 16), r[c16]);
      };
      auto release_tmem = [&]() {                          // this warp's last TMEM read of the tile has completed
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&t_empty[s]);
      };
      auto swish_block = [&](const uint32_t (&r)[CW / 16][16], int mb) {
        const int row = mb * 128 + quad * 32 + lane;     // pixel of the input tile
        if (row >= g.n_pix) return;
        const int iy = (int)(((uint32_t)row * iw_magic) >> 20), ix = row - iy * IW;     // row / IW (row < 4096)
        const int gy = y0 + iy, gx = x0 + ix;
        const bool in_img = gy >= 0 && gy < H && gx >= 0 && gx < W;
        uint8_t* dst = exp_tile + (size_t)row * PSTRIDE;
#pragma unroll
        for (int c16 = 0; c16 < CW / 16; ++c16) {
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const float4 b0 = *reinterpret_cast<const float4*>(hsh_exp + c16 * 16 + h8 * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(hsh_exp + c16 * 16 + h8 * 8 + 4);
            const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                                  make_float2(b1.z, b1.w)};
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 v = silu2(make_float2(__uint_as_float(r[c16][h8 * 8 + 2 * i]), __uint_as_float(r[c16][h8 * 8 + 2 * i + 1])), bb[i]);
              o[i] = in_img ? pack_bf16x2(v.x, v.y) : 0u;
            }
            *reinterpret_cast<uint4*>(dst + c16 * 32 + h8 * 16) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      };
      if constexpr (CW <= 48) {
        uint32_t ra[CW / 16][16], rb[CW / 16][16];
        int mb = half;
        if (mb < g.MB) ld_block(ra, mb);
        while (mb < g.MB) {
          const int mb2 = mb + mstep;
          ptx::tmem_ld_wait();                             // ra landed
          if (mb2 < g.MB) ld_block(rb, mb2); else release_tmem();
          swish_block(ra, mb);
          if (mb2 >= g.MB) break;
          const int mb3 = mb2 + mstep;
          ptx::tmem_ld_wait();                             // rb landed
          if (mb3 < g.MB) ld_block(ra, mb3); else release_tmem();
          swish_block(rb, mb2);
          mb = mb3;
        }
      } else {                                             // 64-channel chunks: two buffers would not fit the registers
        uint32_t ra[CW / 16][16];
        for (int mb = half; mb < g.MB; mb += mstep) {
          ld_block(ra, mb);
          ptx::tmem_ld_wait();
          if (mb + mstep >= g.MB) release_tmem();
          swish_block(ra, mb);
        }
      }
      if (half >= g.MB) {                                // (no block for this warp in a one-block tile)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&t_empty[s]);
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&e_full[s]);       // (release: publishes this warp's tile rows)
    }
  } else {
    // ===================================================================== stencil warps
    const int dtid = tid - dw0;
    const bool active = dtid < n_dw;
    const int cp = active ? dtid % CP : 0, slot = active ? dtid / CP : 0;
    const int c = cbase + 2 * cp;
    float2 wv[K * K];
#pragma unroll
    for (int t = 0; t < K * K; ++t) wv[t] = *reinterpret_cast<const float2*>(w_dw + (size_t)t * C + c);
    float2 hsh = *reinterpret_cast<const float2*>(dw_shift + c);   // depthwise BN shift, pre-halved for silu2
    hsh.x *= 0.5f; hsh.y *= 0.5f;
    for (int step = 0; step < my_steps; ++step) {
      const int s = step & 1;
      const uint32_t ph = (step >> 1) & 1;
      int img, t, x0, y0;
      tile_origin(step, img, t, x0, y0);
      ptx::mbar_wait(&e_full[s], ph);
      const uint8_t* exp_tile = exp_buf + s * g.exp_bytes;
      const int oy_t = (t / g.d.tiles_x) * g.d.TH, ox_t = (t % g.d.tiles_x) * g.d.TW;
      const int oy_end = min(g.d.TH, Ho - oy_t), ox_end = min(g.d.TW, Wo - ox_t);
      float2 psum = make_float2(0.f, 0.f);
      if (active)
        psum = dw_tile<K, S>(exp_tile + cp * 4, PSTRIDE, IW * PSTRIDE, wv, hsh, out, img, oy_t, ox_t, oy_end, ox_end, Ho, Wo,
                             C, c, slot, g.d.n_strips, g.d.strips_x, g.d.NS);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&e_empty[s]);      // tile buffer free for the drain of step + 2
      float2* rb = red + s * n_dw_pad;
      rb[dtid] = psum;
      asm volatile("bar.sync 1, %0;" ::"r"(n_dw_pad) : "memory");   // stencil warps only
      if (active && slot == 0) {
        float2 sum = rb[cp];
        for (int sl = 1; sl < g.d.NS; ++sl) { const float2 v = rb[sl * CP + cp]; sum.x += v.x; sum.y += v.y; }
        *reinterpret_cast<float2*>(pool_part + ((size_t)img * tiles + t) * C + c) = sum;   // one writer per entry
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
  }
}

}  // namespace mt
