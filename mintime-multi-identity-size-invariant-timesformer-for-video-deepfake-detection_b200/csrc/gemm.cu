// out = epilogue(A[M][K] * W[N][K]^T): the contraction behind every 1x1 convolution of EfficientNet-B0
// (reference model.py:101,118,286 via utils.py:273-276) and every nn.Linear of the TimeSformer
// (size_invariant_timesformer.py:68-73,102-106,175).
//
//  * gemm_tc_kernel (bf16): persistent, warp-specialised tcgen05 GEMM.
//      warp 0      TMA producer: A (128 x 64) and W (block_n x 64) tiles, 128-byte swizzle, mbarrier ring
//      warp 1      MMA issuer: tcgen05.mma kind::f16, fp32 accumulators in TMEM, double buffered
//      warps 2-5   epilogue: tcgen05.ld -> bias / swish / GEGLU in registers -> swizzled smem staging ->
//      (2-9 for    TMA store (bf16) or TMA reduce-add (fp32 residual stream), double buffered, so global
//       bf16       writes are full 128-byte lines issued by the copy engine instead of per-row stores; the
//       stores)    8-warp bf16 form has no block-wide barrier (per-warp slabs, straight-line 32 x 32 units)
//      last 4      (GATED only) scale the landed A tile by the squeeze-excite gate (bf16, HMUL2) before the MMA reads it
//    Operands are K-major (A [M][K], W [N][K]); mn_major = 1 takes both with the contraction index slow ([K][M], [K][N]:
//    the weight gradient dW = dY^T X straight from the row-major tensors).  The direct-store epilogue (TMA_OUT = false)
//    is kept for the patch-embedding (+pos/size embedding rows) and for outputs whose pitch TMA cannot address.
//  * gemm_simt_kernel (fp32 / any T): plain FFMA tiles with the same epilogues -- the exact path.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include <stdio.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mt {
namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                          // 64 bf16 = one 128-byte swizzle row
constexpr int kAStageBytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kStagingBytes = kBlockM * 128;         // one epilogue staging buffer: 128 rows x 128 B
constexpr int kEpiThread0 = 64;                      // first epilogue thread (warp 2, lane 0)

struct TcParams {
  int M, N, K;
  int block_n;      // multiple of 16, <= 256 (multiple of 128 for GEGLU)
  int stages;
  int tiles_m, tiles_n;
  int tmem_cols;    // power of two >= acc_stages * block_n
  int acc_stages;   // TMEM accumulator buffers (2, or 1 when two CTAs share the SM's 512 columns at block_n > 128)
  int gate_imgs;    // > 0: SE gate rows of up to this many images are staged in smem per tile (GATED)
  int b_resident;   // 1: the block's weight slice (one column tile, all k-blocks) is loaded into smem once per block
  int mn_major;     // 1: both operands MN-major (64-wide atoms, see ptx::umma_smem_desc_mn_sw128)
  int n_outer;      // 1: 2-D grid, blockIdx.y = the block's (fixed) column tile, blockIdx.x strides over the row tiles
  int splits;       // split-K (EPI_RESID_F32 through the TMA reduce-add only): every output tile is worked on by
  int kb_per_split; // `splits` tiles, each over kb_per_split k-blocks; the cp.reduce .add epilogue sums them in L2
  const float* gate;
  int rows_per_gate;
  EpiParams epi;
};

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int stages) {
    if (++stage == stages) { stage = 0; phase ^= 1; }
  }
};

// erf via Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7): 2 MUFU + ~10 FMA instead of erff's ~25
// instructions -- the GEGLU epilogue evaluates 128 of them per thread per tile.
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.0f - p * t * __expf(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float fast_gelu(float x) { return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752f)); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void bar_sync_epi() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// staging buffer = 128 rows of 128 B with the TMA 128-byte swizzle: 16-byte piece p of row r lives at
// r*128 + ((p ^ (r & 7)) << 4)  (conflict-free for a warp writing 32 different rows)
__device__ __forceinline__ uint4* staging_piece(uint8_t* buf, int row, int piece) {
  return reinterpret_cast<uint4*>(buf + row * 128 + ((piece ^ (row & 7)) << 4));
}

// epilogue warps: 8 for the bf16 TMA-store epilogue (two per TMEM lane quadrant, each taking half of the
// columns of a 64-column chunk) -- the extractor's narrow 1x1 convolutions are bound by epilogue issue
// slots and MUFU latency, not by the MMA -- 4 for the other epilogues
template <int KIND, bool TMA_OUT>
constexpr int epi_warps() { return (KIND == EPI_STORE && TMA_OUT) ? 8 : 4; }
template <int KIND, bool GATED, bool TMA_OUT>
constexpr int tc_threads() { return 64 + 32 * epi_warps<KIND, TMA_OUT>() + (GATED ? 128 : 0); }

// One 32-row x 32-column unit of the 8-warp bf16 epilogue: accumulator registers -> BN shift (+ swish) (+ skip row) ->
// bf16 -> the warp's staging slab (64-byte swizzle: 16-byte piece g of row `lane` sits at piece g ^ ((lane >> 1) & 3)).
// ACT / RES / FULL are compile-time so that the unit is straight-line code; the arithmetic is the library's one
// swish form (h = x/2 + shift/2, h + h*tanh(h); bias_c holds the pre-halved shift when ACT).
template <bool ACT, bool RES, bool FULL>
__device__ __forceinline__ void epi8_unit(const uint32_t (&r)[32], const float* bias_c, const bf16* rrow, int groups,
                                          uint8_t* stg, int lane) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float2 v[4];
    if (FULL || g < groups) {
      const float4 b0 = *reinterpret_cast<const float4*>(bias_c + g * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(bias_c + g * 8 + 4);
      const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                            make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x = make_float2(__uint_as_float(r[g * 8 + 2 * i]), __uint_as_float(r[g * 8 + 2 * i + 1]));
        if (ACT) {
          const float2 h = __ffma2_rn(x, make_float2(0.5f, 0.5f), bb[i]);
          float2 t;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
          v[i] = __ffma2_rn(h, t, h);
        } else {
          v[i] = __fadd2_rn(x, bb[i]);
        }
      }
      if (RES) {
        if (rrow != nullptr) {                       // MBConv skip connection (model.py:123-127), added after BN
          const uint4 q = *reinterpret_cast<const uint4*>(rrow + g * 8);
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
            v[i] = __fadd2_rn(v[i], make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u)));
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = make_float2(0.f, 0.f);
    }
    *reinterpret_cast<uint4*>(stg + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
        make_uint4(pack_bf16(v[0].x, v[0].y), pack_bf16(v[1].x, v[1].y), pack_bf16(v[2].x, v[2].y),
                   pack_bf16(v[3].x, v[3].y));
  }
}

template <int N>
__device__ __forceinline__ void bar_sync_epi_n() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

template <typename T, int KIND, bool GATED, bool TMA_OUT>
__global__ void __launch_bounds__((tc_threads<KIND, GATED, TMA_OUT>()), (epi_warps<KIND, TMA_OUT>() == 8 ? 2 : 1))
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_out, TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up: [A stages][B stages][2 staging buffers (TMA_OUT)][barriers][tmem ptr]
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned; stays a shared-space pointer (LDS/STS, not generic LD/ST)
  const int b_stage_bytes = p.block_n * kBlockK * 2;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.stages * kAStageBytes;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  uint8_t* smem_stage = smem_b + (p.b_resident ? num_kb : p.stages) * b_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + (TMA_OUT ? 2 * kStagingBytes : 0));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* gated_bar = bars + 2 * kMaxStages;
  uint64_t* tmem_full = bars + 3 * kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* b_full = tmem_empty + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_full + 1);
  float* bias_s = reinterpret_cast<float*>(tmem_ptr_smem + 2);   // [N] (8-warp epilogue only)
  float* gate_s = bias_s + ((p.N + 3) & ~3);                     // [gate_imgs][K] (GATED, when staged)
  constexpr int kEW = epi_warps<KIND, TMA_OUT>();
  constexpr int kGateThread0 = 64 + 32 * kEW;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
    if (TMA_OUT) ptx::prefetch_tmap(&tmap_out);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
      ptx::mbar_init(&gated_bar[s], 4);       // one arrival per gate warp
    }
    ptx::mbar_init(b_full, 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      // 8-warp epilogue: one arrival per warp; 4 warps own a buffer when there are two buffers, else all 8
      ptx::mbar_init(&tmem_empty[s], kEW == 8 ? (p.acc_stages == 2 ? 4 : 8) : 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr_smem, (uint32_t)p.tmem_cols);
    ptx::tmem_relinquish();
  }
  if (kEW == 8) {
    // bias (BN shift) staged once per block; pre-halved when the swish follows (silu2 takes shift/2)
    const float bs = p.epi.act == 1 ? 0.5f : 1.0f;
    for (int i = threadIdx.x; i < p.N; i += blockDim.x) bias_s[i] = p.epi.bias ? p.epi.bias[i] * bs : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int mn_tiles = p.tiles_m * p.tiles_n;
  const int num_tiles = mn_tiles * p.splits;      // tile = split * mn_tiles + (m-tile, n-tile)
  // tile walk: grid-stride over all tiles, or (n_outer) a fixed column tile per block (blockIdx.y) with a stride over
  // the row tiles, so that the block's weight slice can stay in shared memory
  const int tile0 = p.n_outer ? (int)blockIdx.x * p.tiles_n + (int)blockIdx.y : (int)blockIdx.x;
  const int tstep = p.n_outer ? (int)gridDim.x * p.tiles_n : (int)gridDim.x;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      PipeState ps;
      // The extractor's early 1x1 convolutions have a weight matrix of a few KiB: re-fetching it for every
      // 128-row tile makes every SM hammer the same handful of L2 lines (measured: 250 of 580 us at N=96,
      // K=16), so it is loaded ONCE per block when it fits.
      if (p.b_resident) {
        ptx::mbar_arrive_expect_tx(b_full, (uint32_t)(num_kb * b_stage_bytes));
        for (int kb = 0; kb < num_kb; ++kb)
          ptx::tma_load_2d(smem_b + kb * b_stage_bytes, &tmap_b, b_full, kb * kBlockK, (int)blockIdx.y * p.block_n);
      }
      const uint32_t tx_bytes = (uint32_t)(kAStageBytes + (p.b_resident ? 0 : b_stage_bytes));
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        const int split = tile / mn_tiles, t2 = tile - split * mn_tiles;
        const int m0 = (t2 / p.tiles_n) * kBlockM;
        const int n0 = (t2 % p.tiles_n) * p.block_n;
        const int kb0 = split * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&empty_bar[ps.stage], ps.phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[ps.stage], tx_bytes);
          if (p.mn_major) {
            // 64-channel atoms (8 KiB: 64 tokens x 128 bytes) side by side; rows beyond K / columns beyond M, N zero-fill
            for (int j = 0; j < kBlockM / 64; ++j)
              ptx::tma_load_2d(smem_a + ps.stage * kAStageBytes + j * 8192, &tmap_a, &full_bar[ps.stage], m0 + j * 64,
                               kb * kBlockK);
            for (int j = 0; j < p.block_n / 64; ++j)
              ptx::tma_load_2d(smem_b + ps.stage * b_stage_bytes + j * 8192, &tmap_b, &full_bar[ps.stage], n0 + j * 64,
                               kb * kBlockK);
          } else {
            ptx::tma_load_2d(smem_a + ps.stage * kAStageBytes, &tmap_a, &full_bar[ps.stage], kb * kBlockK, m0);
            if (!p.b_resident)
              ptx::tma_load_2d(smem_b + ps.stage * b_stage_bytes, &tmap_b, &full_bar[ps.stage], kb * kBlockK, n0);
          }
          ps.advance(p.stages);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      PipeState ps;
      const uint32_t idesc = p.mn_major ? ptx::umma_idesc_bf16_f32_mn(kBlockM, (uint32_t)p.block_n)
                                        : ptx::umma_idesc_bf16_f32(kBlockM, (uint32_t)p.block_n);
      if (p.b_resident) ptx::mbar_wait(b_full, 0);
      int it = 0;
      for (int tile = tile0; tile < num_tiles; tile += tstep, ++it) {
        const int as = p.acc_stages == 2 ? (it & 1) : 0;
        const uint32_t aphase = (p.acc_stages == 2 ? (it >> 1) : it) & 1;
        ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * p.block_n);
        const int kb0 = (tile / mn_tiles) * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(GATED ? &gated_bar[ps.stage] : &full_bar[ps.stage], ps.phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem_a + ps.stage * kAStageBytes);
          const uint32_t b_addr = ptx::smem_u32(smem_b + (p.b_resident ? kb : ps.stage) * b_stage_bytes);
          const int k_left = p.K - kb * kBlockK;
          const int ksteps = k_left >= kBlockK ? 4 : (k_left + 15) / 16;  // TMA zero-filled the K tail
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t adesc = p.mn_major ? ptx::umma_smem_desc_mn_sw128(a_addr + k * 2048, 8192)
                                              : ptx::umma_smem_desc_sw128(a_addr + k * 32);
            const uint64_t bdesc = p.mn_major ? ptx::umma_smem_desc_mn_sw128(b_addr + k * 2048, 8192)
                                              : ptx::umma_smem_desc_sw128(b_addr + k * 32);
            ptx::umma_bf16_ss(tmem_d, adesc, bdesc, idesc, ((kb - kb0) | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[ps.stage]);  // smem slot free once these MMAs retire
          ps.advance(p.stages);
        }
        ptx::umma_commit(&tmem_full[as]);          // accumulator ready for the epilogue warps
      }
    }
    __syncwarp();
  } else if (kEW == 8 && warp < 10) {
    // ===================================================================== epilogue (8 warps, bf16 TMA store)
    // No block-wide barriers: a warp owns the 32 accumulator rows of its TMEM lane quadrant.  With two
    // accumulator buffers the two warps of a quadrant take alternate TILES (one warp set per buffer: two
    // independent MMA -> epilogue pipelines per block); with one buffer they split the tile's 32-column units.
    // Per unit: tcgen05.ld -> BN shift / swish / skip in registers -> its own 2 KiB staging slab (64-byte
    // swizzle) -> its own TMA store of a 32 x 32 box.  Slabs are double buffered per warp.
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;
    const int trow = quad * 32 + lane;
    const EpiParams& e = p.epi;
    const bool act = e.act == 1;
    const bf16* resid = reinterpret_cast<const bf16*>(e.resid);
    uint8_t* slab = smem_stage + (warp - 2) * 4096;          // 2 x 2 KiB
    const int n_units = (p.block_n + 31) >> 5;
    uint32_t store_it = 0;
    int it = 0;
    for (int tile = tile0; tile < num_tiles; tile += tstep, ++it) {
      const int as = p.acc_stages == 2 ? (it & 1) : 0;
      const uint32_t aphase = (p.acc_stages == 2 ? (it >> 1) : it) & 1;
      if (p.acc_stages == 2 && as != half) continue;
      const int t2 = tile % mn_tiles;
      const int m0 = (t2 / p.tiles_n) * kBlockM;
      const int n0 = (t2 % p.tiles_n) * p.block_n;
      const int row = m0 + trow;
      ptx::mbar_wait(&tmem_full[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * p.block_n);
      const int u0 = p.acc_stages == 2 ? 0 : ((half + it) & 1), ustep = p.acc_stages == 2 ? 1 : 2;
      const int units_here = min(n_units, (p.N - n0 + 31) >> 5);     // the last column tile may be narrower
      for (int u = u0; u < units_here; u += ustep) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(taddr + u * 32, r);
        uint8_t* stg = slab + (store_it & 1) * 2048;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // slab's previous store drained
        __syncwarp();
        const int col = n0 + u * 32;
        const int groups = min(4, (p.N - col) >> 3);                 // valid 8-column groups (N is a multiple of 8)
        const bf16* rrow = (resid != nullptr && row < p.M) ? resid + (size_t)row * e.ldo + col : nullptr;
        // one warp-uniform dispatch per unit instead of per-element predicates (the epilogue is issue bound)
        const int variant = (act ? 4 : 0) | (resid != nullptr ? 2 : 0) | (groups == 4 ? 1 : 0);
        ptx::tmem_ld_wait();
        switch (variant) {
          case 0: epi8_unit<false, false, false>(r, bias_s + col, rrow, groups, stg, lane); break;
          case 1: epi8_unit<false, false, true>(r, bias_s + col, rrow, groups, stg, lane); break;
          case 2: epi8_unit<false, true, false>(r, bias_s + col, rrow, groups, stg, lane); break;
          case 3: epi8_unit<false, true, true>(r, bias_s + col, rrow, groups, stg, lane); break;
          case 4: epi8_unit<true, false, false>(r, bias_s + col, rrow, groups, stg, lane); break;
          case 5: epi8_unit<true, false, true>(r, bias_s + col, rrow, groups, stg, lane); break;
          case 6: epi8_unit<true, true, false>(r, bias_s + col, rrow, groups, stg, lane); break;
          default: epi8_unit<true, true, true>(r, bias_s + col, rrow, groups, stg, lane); break;
        }
        ptx::fence_proxy_async_smem();             // slab writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(ptx::smem_u32(stg)), "r"(col), "r"(m0 + quad * 32)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++store_it;
      }
      // all TMEM reads of this warp for the tile have completed (tcgen05.wait::ld above): hand it back
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
  } else if (kEW == 4 && warp < 6) {
    // ===================================================================== epilogue (4 warps)
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may read
    const int trow = quad * 32 + lane;             // row of the tile this thread owns
    const EpiParams& e = p.epi;
    uint32_t store_it = 0;                         // staging buffer ring position (TMA_OUT)
    int it = 0;
    for (int tile = tile0; tile < num_tiles; tile += tstep, ++it) {
      const int as = p.acc_stages == 2 ? (it & 1) : 0;
      const uint32_t aphase = (p.acc_stages == 2 ? (it >> 1) : it) & 1;
      const int t2 = tile % mn_tiles;
      const int m0 = (t2 / p.tiles_n) * kBlockM;
      const int n0 = (t2 % p.tiles_n) * p.block_n;
      const int row = m0 + trow;
      const bool row_ok = row < p.M;
      ptx::mbar_wait(&tmem_full[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * p.block_n);

      if (TMA_OUT) {
        // accumulator columns consumed per staging buffer: 64 (bf16 store), 128 (GEGLU -> 64 outputs),
        // 32 (fp32 reduce-add); each fills 128 B per row
        constexpr int kChunk = KIND == EPI_GEGLU ? 128 : (KIND == EPI_RESID_F32 ? 32 : 64);
        for (int c = 0; c < p.block_n; c += kChunk) {
          uint8_t* stg = smem_stage + (store_it & 1) * kStagingBytes;
          if (threadIdx.x == kEpiThread0)
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // buffer's previous store drained
          bar_sync_epi();
          if (KIND == EPI_GEGLU) {
#pragma unroll
            for (int ob = 0; ob < 2; ++ob) {        // two 64-wide interleave blocks: [32 u | 32 gate]
              uint32_t ru[32], rg[32];
              ptx::tmem_ld_32x32b_x32(taddr + c + ob * 64, ru);
              ptx::tmem_ld_32x32b_x32(taddr + c + ob * 64 + 32, rg);
              ptx::tmem_ld_wait();
              const int col = n0 + c + ob * 64;     // packed column of u[0]
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float u[8], gg[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { u[i] = __uint_as_float(ru[g * 8 + i]); gg[i] = __uint_as_float(rg[g * 8 + i]); }
                if (e.bias && col + 64 <= p.N) {
                  float b[8];
                  load8(e.bias + col + g * 8, b);
#pragma unroll
                  for (int i = 0; i < 8; ++i) u[i] += b[i];
                  load8(e.bias + col + 32 + g * 8, b);
#pragma unroll
                  for (int i = 0; i < 8; ++i) gg[i] += b[i];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {       // same packed form as the CTA-pair kernel: identical bits
                  const float2 r2 = geglu2(make_float2(u[2 * i], u[2 * i + 1]), make_float2(gg[2 * i], gg[2 * i + 1]));
                  u[2 * i] = r2.x; u[2 * i + 1] = r2.y;
                }
                *staging_piece(stg, trow, ob * 4 + g) =
                    make_uint4(pack_bf16(u[0], u[1]), pack_bf16(u[2], u[3]), pack_bf16(u[4], u[5]), pack_bf16(u[6], u[7]));
              }
            }
          } else {
            constexpr int kQ = kChunk / 16;         // 16-column quarters in this chunk
            uint32_t r[kQ][16];
#pragma unroll
            for (int q = 0; q < kQ; ++q)
              if (c + q * 16 < p.block_n) ptx::tmem_ld_32x32b_x16(taddr + c + q * 16, r[q]);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < kQ; ++q) {
              if (c + q * 16 >= p.block_n) continue;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int col = n0 + c + q * 16 + hh * 8;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[q][hh * 8 + i]);
                float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (e.bias && col < p.N) load8(e.bias + col, b);
                if (KIND == EPI_STORE && e.act == 1) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = swish_shift_fast(v[i], b[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] += b[i];
                }
                if (KIND == EPI_STORE) {
                  *staging_piece(stg, trow, q * 2 + hh) =
                      make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                } else {                            // EPI_RESID_F32: 8 fp32 = two 16-byte pieces
                  *staging_piece(stg, trow, q * 4 + hh * 2) =
                      make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
                  *staging_piece(stg, trow, q * 4 + hh * 2 + 1) =
                      make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
                }
              }
            }
          }
          if (c + kChunk >= p.block_n) {            // last TMEM read of this tile: hand the accumulator back
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tmem_empty[as]);
          }
          ptx::fence_proxy_async_smem();            // staging writes -> visible to the TMA engine
          bar_sync_epi();
          if (threadIdx.x == kEpiThread0) {
            const int ocol = KIND == EPI_GEGLU ? (n0 + c) / 2 : n0 + c;
            if (KIND == EPI_RESID_F32)
              asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(ptx::smem_u32(stg)), "r"(ocol), "r"(m0)
                           : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(ptx::smem_u32(stg)), "r"(ocol), "r"(m0)
                           : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++store_it;
        }
      } else {
        // ------------------------------------------------------------- direct (per-row) stores
        for (int c = 0; c < p.block_n; c += 16) {
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(taddr + c, r);
          ptx::tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int col = n0 + c + g * 8;
              if (col < p.N) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
                epi_store8<T, KIND>(e, row, col, v);
              }
            }
          }
        }
        ptx::tc_fence_before();
        ptx::mbar_arrive(&tmem_empty[as]);
      }
    }
    if (TMA_OUT && threadIdx.x == kEpiThread0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else if (GATED) {
    // ===================================================================== SE-gate warps (4 warps)
    // Multiply the freshly landed A tile by gate[image(row)][k] in shared memory (reference
    // model.py:115: x = sigmoid(se) * x, ahead of the project conv), then hand it to the MMA warp.
    // The gate rows of the (at most gate_imgs) images a 128-row tile touches are staged in shared memory once
    // per tile with coalesced loads, so the TMA -> gate -> MMA critical path holds no global-memory latency.
    const int r = threadIdx.x - kGateThread0;  // tile row 0..127
    PipeState ps;
    for (int tile = tile0; tile < num_tiles; tile += tstep) {
      const int m0 = ((tile % mn_tiles) / p.tiles_n) * kBlockM;
      const int row = m0 + r;
      const int img = (row < p.M ? row : p.M - 1) / p.rows_per_gate;
      const float* grow = p.gate + (size_t)img * p.K;   // global row (fallback when the gates are not staged)
      const bf16* gsm = nullptr;
      if (p.gate_imgs > 0) {
        // staged gates are held as bf16 so that the scaling is one packed HMUL2 per channel pair (the fp32 form
        // costs 5x the instructions and made the gate warps the bottleneck of the long-K project convs:
        // 49.6 vs 26.7 us ungated at K = 1152, N = 192)
        const int img0 = m0 / p.rows_per_gate;
        const int img1 = min(m0 + kBlockM - 1, p.M - 1) / p.rows_per_gate;
        asm volatile("bar.sync 2, 128;" ::: "memory");   // everyone is done with the previous tile's gates
        const float4* src = reinterpret_cast<const float4*>(p.gate + (size_t)img0 * p.K);
        const int n4 = (img1 - img0 + 1) * (p.K >> 2);
        bf16* gs = reinterpret_cast<bf16*>(gate_s);
        for (int i = r; i < n4; i += 128) {
          const float4 g4 = src[i];
          *reinterpret_cast<uint2*>(gs + 4 * i) = make_uint2(pack_bf16(g4.x, g4.y), pack_bf16(g4.z, g4.w));
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        gsm = gs + (size_t)(img - img0) * p.K;
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        // Gates as packed bf16 (staged in shared memory, or -- single k-block shapes -- read from the fp32 gate row
        // in global memory and packed here), one HMUL2 per channel pair.  Per half k-block (4 x 16 bytes of the row):
        // all loads first, then the multiplies, then the stores -- the A tile and the staged gates are both shared
        // memory, so an interleaved form serialises on possible aliasing; straight-line code for full halves (per-piece
        // guards made the compiler rematerialise the addresses in every piece).  72 registers per thread.
        const int valid = min(8, (p.K - kb * kBlockK) >> 3);       // 16-byte pieces inside K (K % 8 == 0)
        const bf16* gk = gsm ? gsm + kb * kBlockK : nullptr;
        const float* gf = grow + kb * kBlockK;
        auto gate_piece = [&](int c) -> uint4 {
          if (gk) return *reinterpret_cast<const uint4*>(gk + c * 8);
          const float4 g0 = *reinterpret_cast<const float4*>(gf + c * 8);
          const float4 g1 = *reinterpret_cast<const float4*>(gf + c * 8 + 4);
          return make_uint4(pack_bf16(g0.x, g0.y), pack_bf16(g0.z, g0.w), pack_bf16(g1.x, g1.y), pack_bf16(g1.z, g1.w));
        };
        uint8_t* arow = smem_a + ps.stage * kAStageBytes + r * 128;
        const uint32_t sw = (uint32_t)(r & 7) << 4;
        uint4 gq[4], u[4];
        if (valid >= 4) {
#pragma unroll
          for (int c = 0; c < 4; ++c) gq[c] = gate_piece(c);       // before the wait: overlaps the TMA's latency
        }
        ptx::mbar_wait(&full_bar[ps.stage], ps.phase);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int nv = valid - half * 4;                          // pieces of this half inside K
          if (nv >= 4) {
            if (half == 1) {
#pragma unroll
              for (int c = 0; c < 4; ++c) gq[c] = gate_piece(4 + c);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) u[c] = *reinterpret_cast<const uint4*>(arow + (((half * 4 + c) << 4) ^ sw));
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u[c]);
              const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&gq[c]);
#pragma unroll
              for (int i = 0; i < 4; ++i) h[i] = __hmul2(h[i], gh[i]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(arow + (((half * 4 + c) << 4) ^ sw)) = u[c];
          } else {
            for (int c = 0; c < nv; ++c) {                          // K tail (at most one half per tile)
              const uint4 g1 = gate_piece(half * 4 + c);
              uint4* ptr = reinterpret_cast<uint4*>(arow + ((uint32_t)((half * 4 + c) << 4) ^ sw));
              uint4 v = *ptr;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
              const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&g1);
#pragma unroll
              for (int i = 0; i < 4; ++i) h[i] = __hmul2(h[i], gh[i]);
              *ptr = v;
            }
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&gated_bar[ps.stage]);
        ps.advance(p.stages);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// =====================================================================================================
// SIMT kernel (exact path): 64x64 tile, 256 threads, 4 rows x (2+2) columns per thread so that a
// GEGLU u-column and its gate column (+32) live in the same thread.
// =====================================================================================================
template <typename T, int KIND, bool GATED>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const T* A = reinterpret_cast<const T*>(g.a);
  const T* W = reinterpret_cast<const T*>(g.w);
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;   // (row tiles on x: 512 images x 112^2 pixels exceed grid.y)
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
    for (int e = threadIdx.x; e < BM * BK; e += 256) {
      const int r = e / BK, kk = e % BK;
      const int row = m0 + r, k = k0 + kk;
      float v = 0.f;
      if (row < g.M && k < g.K) {
        v = to_f(A[(size_t)row * g.K + k]);
        if (GATED) v *= g.gate[(size_t)(row / g.rows_per_gate) * g.K + k];
      }
      As[kk][r] = v;
      const int col = n0 + r;
      Bs[kk][r] = (col < g.N && k < g.K) ? to_f(W[(size_t)col * g.K + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
      b[0] = Bs[kk][tx * 2]; b[1] = Bs[kk][tx * 2 + 1]; b[2] = Bs[kk][32 + tx * 2]; b[3] = Bs[kk][33 + tx * 2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  constexpr bool kExact = sizeof(T) == 4;
  const EpiParams& p = g.epi;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= g.M) continue;
    if (KIND == EPI_GEGLU) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = n0 + tx * 2 + j;  // u column; gate at +32
        if (col < g.N) {
          float u = acc[i][j] + (p.bias ? p.bias[col] : 0.f);
          float gt = acc[i][j + 2] + (p.bias ? p.bias[col + 32] : 0.f);
          const int ocol = (col >> 6) * 32 + (col & 63);
          reinterpret_cast<T*>(p.out)[(size_t)row * p.ldo + ocol] = from_f<T>(u * gelu_erf(gt));
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + tx * 2 + (j & 1) + (j >> 1) * 32;
        if (col >= g.N) continue;
        float v = acc[i][j] + (p.bias ? p.bias[col] : 0.f);
        if (KIND == EPI_STORE) {
          if (p.act == 1) v = silu<kExact>(v);
          const size_t off = (size_t)row * p.ldo + col;
          if (p.resid) v += to_f(reinterpret_cast<const T*>(p.resid)[off]);
          reinterpret_cast<T*>(p.out)[off] = from_f<T>(v);
        } else if (KIND == EPI_RESID_F32) {
          float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col;
          *o = *o + v;
        } else if (KIND == EPI_PATCH_EMBED) {
          const int b = row / p.rows_per_batch, t = row - b * p.rows_per_batch;
          const size_t orow = (size_t)row + b + 1;
          const long long pos = p.positions ? p.positions[(size_t)b * (p.rows_per_batch + 1) + 1 + t] : (long long)(1 + t);
          if ((unsigned long long)pos >= (unsigned long long)p.table_rows) __trap();
          v += p.pos_tab[(size_t)pos * p.ldo + col];
          if (p.size_tab) {
            const int si = p.size_idx[b * p.frames + t / p.n_patches];
            if ((unsigned)si >= (unsigned)p.table_rows) __trap();
            v += p.size_tab[(size_t)si * p.ldo + col];
          }
          reinterpret_cast<float*>(p.out)[orow * p.ldo + col] = v;
        }
      }
    }
  }
}

// =====================================================================================================
// host side
// =====================================================================================================
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// 2D row-major [rows][cols] tensor (row pitch = pitch_elems), box = [box_rows][128 bytes of columns],
// 128-byte swizzle, zero fill / clipping out of bounds.
int make_tmap_2d(CUtensorMap* m, const void* base, bool f32, int rows, int cols, int pitch_elems, int box_rows,
                 int box_bytes = 128) {
  auto enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MT_ERR_DRIVER;
  }
  const int es = f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * es};
  cuuint32_t box[2] = {(cuuint32_t)(box_bytes / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   box_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                   : (box_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d pitch=%d box_rows=%d base=%p", (int)r, rows, cols,
              pitch_elems, box_rows, base);
    return MT_ERR_DRIVER;
  }
  return MT_OK;
}

int num_sms() { return current_sms(); }

// largest weight slice kept resident by a one-block-per-SM launch (MINTIME_B200_WRES_KB, read once; 0 disables)
int wres_limit_bytes() {
  static const int v = [] {
    const char* e = getenv("MINTIME_B200_WRES_KB");
    return std::min(e ? atoi(e) : 136, 136) * 1024;     // 136 KiB + staging/bias + 3 A stages = the 226 KiB budget
  }();
  return v;
}

template <int KIND, bool GATED, bool TMA_OUT>
int launch_tc_impl(const GemmArgs& g, cudaStream_t stream) {
  TcParams p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.gate = g.gate; p.rows_per_gate = g.rows_per_gate;
  p.epi = g.epi;
  int bn;
  if (KIND == EPI_GEGLU) {
    bn = 256;
  } else {
    const int parts = (g.N + 255) / 256;              // fewest UMMA-wide (<= 256) column tiles ...
    bn = ((g.N + parts - 1) / parts + 15) / 16 * 16;  // ... of equal width (multiple of 16)
    // several column tiles: a TMA store box is 128 B wide, so tiles must start on 64-column boundaries
    // (a partial last box of tile i would otherwise spill stale staging data into tile i+1's columns)
    if (parts > 1) bn = (bn + 63) / 64 * 64;
    if (g.mn_major) bn = (bn + 63) / 64 * 64;         // MN-major operands: whole 64-channel atoms per tile (TMA zero-fills beyond N)
  }
  p.block_n = bn;
  p.tiles_m = (g.M + kBlockM - 1) / kBlockM;
  p.tiles_n = (g.N + bn - 1) / bn;
  constexpr int kEW = epi_warps<KIND, TMA_OUT>();
  const int num_kb = (g.K + kBlockK - 1) / kBlockK;
  const int b_block = bn * kBlockK * 2;
  p.mn_major = g.mn_major;
  if (g.mn_major && (bn % 64 != 0 || GATED)) {
    set_error("gemm_tc: MN-major operands need 64-column tiles (N = %d) and no SE gate", g.N);
    return MT_ERR_ARG;
  }
  p.b_resident = (!g.mn_major && p.tiles_n == 1 && num_kb * b_block <= 40 * 1024) ? 1 : 0;
  p.n_outer = 0;
  // Wide layers with several k-blocks run one block per SM (below) and were bound by re-streaming the weight slice
  // for every 128-row tile (L2 -> SM: e.g. N=1152 K=192 moves 98 KB of W next to 49 KB of A per tile).  Give each
  // block ONE column tile (2-D grid) and keep that slice in shared memory when it fits beside >= 3 A stages and the
  // block has at least two row tiles to use it for.
  if (KIND == EPI_STORE && TMA_OUT && bn > 128 && num_kb > 1 && g.splits <= 1 && !p.b_resident) {
    const int res_limit = wres_limit_bytes();
    const int gx = std::max(1, num_sms() / p.tiles_n);
    if (num_kb * b_block <= res_limit && p.tiles_m >= 2 * gx) {
      p.b_resident = 1;
      p.n_outer = p.tiles_n > 1 ? 1 : 0;
    }
  }
  // split-K: only where partial tiles can be summed by the epilogue itself (fp32 TMA reduce-add, no bias)
  p.splits = 1;
  p.kb_per_split = num_kb;
  if (KIND == EPI_RESID_F32 && TMA_OUT && !GATED && g.splits > 1 && !g.epi.bias && !p.b_resident) {
    int kbs = (num_kb + g.splits - 1) / g.splits;
    if (kbs < 8) kbs = 8;
    p.kb_per_split = kbs;
    p.splits = (num_kb + kbs - 1) / kbs;           // every split owns at least one k-block
  }
  const int stage_bytes = kAStageBytes + (p.b_resident ? 0 : b_block);
  const int bias_bytes = kEW == 8 ? ((g.N * 4 + 15) & ~15) : 0;
  // SE gates staged per tile: a 128-row tile touches at most (127 / rows_per_gate) + 2 images
  p.gate_imgs = 0;
  int gate_bytes = 0;
  if (GATED && kEW == 8 && g.K % 8 == 0 && num_kb >= 2) {
    const int imgs = (kBlockM - 1) / g.rows_per_gate + 2;
    if (imgs * g.K * 2 <= 20 * 1024) { p.gate_imgs = imgs; gate_bytes = (imgs * g.K * 2 + 15) & ~15; }
  }
  const int fixed = (TMA_OUT ? 2 * kStagingBytes : 0) + 1024 /*align slack*/ + (3 * kMaxStages + 5) * 8 + 16 + bias_bytes +
                    gate_bytes + (p.b_resident ? num_kb * b_block : 0);
  // two blocks per SM (2 x (112 KiB + 1 KiB reserved) <= 228 KiB, 2 x 256 TMEM columns) whenever the tile is at
  // most 128 wide, and for the 8-warp epilogue also up to 256 wide when K fits one k-block: the accumulator is
  // then single buffered and the other block's epilogue covers the MMA latency
  const bool wide_pair = kEW == 8 && bn > 128 && num_kb == 1 && 2 * stage_bytes + fixed <= 112 * 1024;
  const int ctas_per_sm = (bn <= 128 || wide_pair) ? 2 : 1;
  p.acc_stages = (ctas_per_sm == 2 && bn > 128) ? 1 : 2;
  int tmem = 32;
  while (tmem < p.acc_stages * bn) tmem <<= 1;
  p.tmem_cols = tmem;
  const int budget = (ctas_per_sm == 2 ? 112 : 226) * 1024 - fixed;
  int stages = budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + fixed;

  CUtensorMap ta, tb, tout;
  int rc = g.mn_major ? make_tmap_2d(&ta, g.a, false, g.K, g.M, g.M, 64) : make_tmap_2d(&ta, g.a, false, g.M, g.K, g.K, kBlockM);
  if (rc) return rc;
  rc = g.mn_major ? make_tmap_2d(&tb, g.w, false, g.K, g.N, g.N, 64) : make_tmap_2d(&tb, g.w, false, g.N, g.K, g.K, bn);
  if (rc) return rc;
  if (TMA_OUT) {
    const bool f32 = KIND == EPI_RESID_F32;
    const int out_cols = KIND == EPI_GEGLU ? g.N / 2 : g.N;
    if (kEW == 8) rc = make_tmap_2d(&tout, g.epi.out, false, g.M, out_cols, g.epi.ldo, 32, 64);   // per-warp 32 x 32 boxes
    else rc = make_tmap_2d(&tout, g.epi.out, f32, g.M, out_cols, g.epi.ldo, kBlockM);
    if (rc) return rc;
  } else {
    tout = ta;
  }

  auto kern = gemm_tc_kernel<bf16, KIND, GATED, TMA_OUT>;
  if (first_use_on_device(reinterpret_cast<const void*>(kern))) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(gemm_tc)");
  }
  dim3 grid(p.tiles_m * p.tiles_n * p.splits);
  if ((int)grid.x > num_sms() * ctas_per_sm) grid.x = num_sms() * ctas_per_sm;
  if (p.n_outer) grid = dim3(std::min(p.tiles_m, std::max(1, num_sms() * ctas_per_sm / p.tiles_n)), p.tiles_n);
  kern<<<grid, tc_threads<KIND, GATED, TMA_OUT>(), smem, stream>>>(ta, tb, tout, p);
  MT_LAUNCH_CHECK("gemm_tc_kernel");
  return MT_OK;
}

#include "gemm2.cuh"

bool tc2_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MINTIME_B200_NO_TC2");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on == 1;
}

// MINTIME_B200_NO_TC2A=1 (read once): keep the streaming pair kernel for the K = 512 contractions (A/B measurements)
bool tc2a_enabled() {
  static const bool on = [] {
    const char* e = getenv("MINTIME_B200_NO_TC2A");
    return !(e && e[0] == '1');
  }();
  return on;
}

constexpr int tc2a_epi_warps(int kind) { return (kind == EPI_GEGLU || kind == EPI_STORE) ? 8 : 4; }

template <int KIND, bool GATED>
int launch_tc(const GemmArgs& g, cudaStream_t stream) {
  if constexpr (!GATED && (KIND == EPI_STORE || KIND == EPI_GEGLU || KIND == EPI_RESID_F32)) {
    // large transformer contractions: CTA-pair kernel (256x256 tiles, cta_group::2)
    if (tc2_enabled() && g.splits <= 1 && !g.mn_major && tc2_eligible(g)) {
      if constexpr (KIND != EPI_RESID_F32) {
        // (8 epilogue warps for the plain store as well: with 4, a thread drains 256 columns per tile and the bf16
        // store of the 4096-wide FF1 pre-activation ran at 138 us against 91 us for the GEGLU epilogue of the same GEMM)
        if (tc2a_enabled() && tc2a_eligible(g)) return launch_tc2a<KIND, tc2a_epi_warps(KIND)>(g, stream);
      }
      return launch_tc2<KIND, (KIND == EPI_GEGLU ? 8 : 4)>(g, stream);
    }
  }
  // per-row gathers (skip connection, embedding rows) use the direct epilogue; everything else TMA
  if (KIND == EPI_PATCH_EMBED) return launch_tc_impl<KIND, GATED, false>(g, stream);
  if ((g.epi.ldo * (KIND == EPI_RESID_F32 ? 4 : 2)) % 16 != 0 || (reinterpret_cast<uintptr_t>(g.epi.out) & 15))
    return launch_tc_impl<KIND, GATED, false>(g, stream);
  return launch_tc_impl<KIND, GATED, true>(g, stream);
}

template <typename T, int KIND, bool GATED>
int launch_simt(const GemmArgs& g, cudaStream_t stream) {
  dim3 grid((g.M + 63) / 64, (g.N + 63) / 64);
  gemm_simt_kernel<T, KIND, GATED><<<grid, 256, 0, stream>>>(g);
  MT_LAUNCH_CHECK("gemm_simt_kernel");
  return MT_OK;
}

template <int KIND, bool GATED>
int dispatch_prec(int precision, const GemmArgs& g, cudaStream_t stream) {
  if (precision == MT_PREC_BF16) return launch_tc<KIND, GATED>(g, stream);
  if (precision == MT_PREC_FP32) return launch_simt<float, KIND, GATED>(g, stream);
  if (precision == 2) return launch_simt<bf16, KIND, GATED>(g, stream);  // debug: bf16 storage, FFMA math
  set_error("unknown precision %d", precision);
  return MT_ERR_ARG;
}

}  // namespace

int make_tmap_nhwc_bf16(CUtensorMap_st* m, const void* base, int n, int h, int w, int c, int box_w, int box_h) {
  auto enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MT_ERR_DRIVER;
  }
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d) n=%d h=%d w=%d c=%d box=%dx%d base=%p", (int)r, n, h, w, c, box_w,
              box_h, base);
    return MT_ERR_DRIVER;
  }
  return MT_OK;
}

// 4-D map over a bf16 NHWC tensor whose box rows (box_c channels = 32 / 64 / 128 bytes) are swizzled to match a
// K-major UMMA operand; channels beyond c are zero-filled
int make_tmap_nhwc_bf16_kmajor(CUtensorMap_st* m, const void* base, int n, int h, int w, int c, int box_c, int box_w,
                               int box_h) {
  auto enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MT_ERR_DRIVER;
  }
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = box_c == 16 ? CU_TENSOR_MAP_SWIZZLE_32B
                                            : (box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d k-major) failed (%d) n=%d h=%d w=%d c=%d box=%dx%dx%d base=%p", (int)r, n, h, w, c,
              box_c, box_w, box_h, base);
    return MT_ERR_DRIVER;
  }
  return MT_OK;
}

// 2-D map over bf16 weights [rows][cols], box = box_rows x box_cols (32 / 64 / 128 bytes, matching swizzle)
int make_tmap_weights_kmajor(CUtensorMap_st* m, const void* base, int rows, int cols, int box_rows, int box_cols) {
  return make_tmap_2d(m, base, false, rows, cols, cols, box_rows, box_cols * 2);
}

// generic 4-D bf16 map, 128-byte swizzle: dims / box innermost first, strides (bytes) of dims 1..3
int make_tmap_4d_bf16_sw128(CUtensorMap_st* m, const void* base, const unsigned long long dims[4],
                            const unsigned long long strides[3], const unsigned box[4]) {
  auto enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MT_ERR_DRIVER;
  }
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {strides[0], strides[1], strides[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), d, st, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d) dims %llu %llu %llu %llu box %u %u %u %u base=%p", (int)r, dims[0],
              dims[1], dims[2], dims[3], box[0], box[1], box[2], box[3], base);
    return MT_ERR_DRIVER;
  }
  return MT_OK;
}

int make_tmap_nhwc_bf16_plain(CUtensorMap_st* m, const void* base, int n, int h, int w, int c, int box_c, int box_w,
                              int box_h) {
  auto enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return MT_ERR_DRIVER;
  }
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d plain) failed (%d) n=%d h=%d w=%d c=%d box=%dx%dx%d base=%p", (int)r, n, h, w, c,
              box_c, box_w, box_h, base);
    return MT_ERR_DRIVER;
  }
  return MT_OK;
}

int launch_gemm(int precision, const GemmArgs& g, cudaStream_t stream) {
  MT_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  MT_REQUIRE(g.N % 8 == 0 && (g.K % 8 == 0 || g.mn_major), "gemm: N (%d) and K (%d) must be multiples of 8", g.N, g.K);
  MT_REQUIRE(g.a && g.w && g.epi.out, "gemm: null operand");
  const bool gated = g.gate != nullptr;
  MT_REQUIRE(!gated || g.rows_per_gate > 0, "gemm: rows_per_gate must be > 0 with a gate");
  MT_REQUIRE(!g.mn_major || (precision == MT_PREC_BF16 && g.epi.kind == EPI_RESID_F32 && g.M % 8 == 0),
             "gemm: MN-major operands are a bf16 tensor-core path of the fp32-accumulate epilogue only");
  static const char* kKind[] = {"store", "resid", "geglu", "embed"};
  const double es = precision == MT_PREC_FP32 ? 4.0 : 2.0;
  const double mn = (double)g.M * g.N;
  double bytes = ((double)g.M * g.K + (double)g.N * g.K) * es;
  if (g.epi.kind == EPI_STORE) bytes += mn * es * (g.epi.resid ? 2 : 1);
  else if (g.epi.kind == EPI_RESID_F32) bytes += mn * 8;
  else if (g.epi.kind == EPI_GEGLU) bytes += mn / 2 * es;
  else if (g.epi.kind == EPI_PATCH_EMBED) bytes += mn * 4 * (g.epi.size_tab ? 3 : 2);
  // (split-K weight gradients carry M in the name: three shapes share N = 512, K = tokens)
  char wg[24] = "";
  if (g.splits > 1) snprintf(wg, sizeof(wg), " wgrad M%d", g.M);
  ProfScope prof(stream, 2.0 * mn * g.K, bytes, "gemm_%s %s%s%s N%d K%d", precision == MT_PREC_BF16 ? "tc" : "simt",
                 (g.epi.kind >= 0 && g.epi.kind < 4) ? kKind[g.epi.kind] : "?", gated ? "+gate" : "", wg, g.N, g.K);
  switch (g.epi.kind) {
    case EPI_STORE:
      return gated ? dispatch_prec<EPI_STORE, true>(precision, g, stream)
                   : dispatch_prec<EPI_STORE, false>(precision, g, stream);
    case EPI_RESID_F32:
      MT_REQUIRE(!gated, "gemm: gate unsupported for this epilogue");
      return dispatch_prec<EPI_RESID_F32, false>(precision, g, stream);
    case EPI_GEGLU:
      MT_REQUIRE(!gated && g.N % 128 == 0, "gemm: GEGLU needs N %% 128 == 0 (N=%d)", g.N);
      return dispatch_prec<EPI_GEGLU, false>(precision, g, stream);
    case EPI_PATCH_EMBED:
      MT_REQUIRE(!gated, "gemm: gate unsupported for this epilogue");
      return dispatch_prec<EPI_PATCH_EMBED, false>(precision, g, stream);
  }
  set_error("gemm: unknown epilogue %d", g.epi.kind);
  return MT_ERR_ARG;
}

}  // namespace mt
