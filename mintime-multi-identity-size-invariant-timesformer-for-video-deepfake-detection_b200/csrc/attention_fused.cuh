// Fused divided attention (reference size_invariant_timesformer.py:109-144 as ONE kernel, bf16 path):
//     per-head QKV projection of the LayerNorm'd tokens  ->  identity-masked softmax(Q K^T) V  ->  out
// The qkv tensor (3 x inner per token, the largest activation of an attention block) never goes to HBM.
//
// Work decomposition
//   * a TILE is 128 token rows of ONE video that hold whole attention groups plus that video's CLS token:
//       TIME  (groups = the f frames at one patch position):  rows p_local*f + frame for PT = (127 / f) patch positions,
//             row f*PT = CLS.  One 4-D TMA box {64 ch, 1 patch, f frames, 1 video} per group and k-block gathers the
//             group's f token rows (stride n tokens) into f consecutive tile rows.
//       SPACE (groups = the n patches of one frame):  two 64-row slots, slot s = frame fr0+s: rows 0..n-1 its patches,
//             row 56 = CLS (n <= 55).  One 4-D TMA box {64 ch, n patches, 1 frame, 1 video} per slot and k-block.
//     The CLS row of the video is a separate 1-row TMA load, so every group finds its CLS key/value inside its own tile.
//   * a CTA PAIR (cta_group::2) works on two tiles at once (M = 256): each CTA keeps ITS tile's A operand (128 x 512
//     bf16 = 128 KiB) resident in shared memory and loops over the heads; per head the pair streams that head's
//     192 x 512 slice of W_qkv (each CTA loads half, 96 rows) through a 3-stage ring and ONE tcgen05.mma sequence
//     (M 256, N 192 = q|k|v of the head, K 512) fills a 128 x 192 fp32 accumulator in each CTA's TMEM (double buffered).
//   * the linearised (pair tile, head) steps are split evenly over the pairs, so the chip is balanced to one head-step.
//
// Warp roles per CTA (608 threads), mbarrier hand-offs only:
//   warp 0      W producer (TMA, 3-stage ring)                     last warp A producer (TMA: tile k-blocks + CLS rows;
//   warp 1      MMA issuer (leader CTA) + TMEM allocation                    next tile's k-block kb is fetched as soon as
//   warps 2-17  epilogue: drain TMEM -> bf16 q/k/v tiles in shared            the last head has consumed the current one)
//               memory (the swizzled 128-byte rows attention_mma.cuh works on) -> attention core on mma.sync m16n8k16
//               (S and P stay in registers, softmax by warp shuffles; the core is 3 % of the block's FLOPs, a dense
//               128 x 128 tcgen05 S tile would spend 8x the MACs on masked pairs) -> `out` rows; CLS-query partials
//               (m, l, o) per group exactly as the unfused kernels write them (cls_combine_kernel merges them).
// The epilogue of head h overlaps the tcgen05 projection of head h+1.
//
// Measured on B200 (B = 32, f = 16; scripts/prof_fused.py, round 2): with the epilogue and every load switched off the
// MMA issue loop alone costs 1.14 us per video = 185 clk per M256 x N192 x K16 tcgen05.mma -- and exactly the same with
// N = 128 or N = 256 in the instruction descriptor, or with half as many instructions' worth of smem stages: the
// CTA-pair tensor pipe retires one K = 16 instruction per ~185 clk whatever N <= 256 is (cuBLAS's 1616 TFLOP/s is that
// rate at N = 256).  N = 192 (one head's q|k|v) therefore caps this kernel at 75 % of the pipe; the full kernel adds
// 0.23 us per video (epilogue not hidden) and ~18 us per launch (prologue, first A tile, 12-vs-13-step imbalance).
#pragma once
#include <cuda.h>

#include "attention_mma.cuh"
#include "ptx.cuh"

namespace mt {
namespace fattn {

constexpr int kDim = 512;                       // model dim = K of the projection (8 k-blocks of 64)
constexpr int kKB = kDim / 64;
constexpr int kAKbBytes = 128 * 128;            // one k-block of the A tile: 128 rows x 128 B
constexpr int kBRows = 96;                      // this CTA's half of a head's 192 weight rows
constexpr int kBStageBytes = kBRows * 128;
constexpr int kBStages = 3;
constexpr int kTileBytes = 128 * 128;           // q / k / v tile: 128 rows x 64 bf16
constexpr int kEpiWarps = 16;                   // the attention core is a chain of dependent smem / mma.sync latencies per
                                                // warp (~0.12 IPC): 8 warps took 4x the projection's MMA time per head
constexpr int kDrainCols = 192 / (kEpiWarps / 4);
constexpr int kAWarp = 2 + kEpiWarps;           // A producer
constexpr int kThreads = 32 * (2 + kEpiWarps + 1);
constexpr int kSpaceClsRow = 56;                // CLS row inside a 64-row frame slot (a multiple of 8)
constexpr int kAccStride = 256;                 // TMEM columns between the two accumulator buffers

struct Params {
  int B, f, n, heads, N;
  int pt;                 // TIME: patch positions per tile
  int tiles_per_video, n_tiles, n_pair_tiles, total_steps;
  int cls_row;            // TIME: f * pt
  int a_bytes_kb;         // bytes one CTA's TMA loads deliver per k-block of A
  const uint8_t* mask;    // [B][f]
  const uint8_t* idmask;  // [B][f][f]
  bf16* out;              // [B*N][heads*64]
  float* cls_parts;       // [B*heads][G][kClsStride]
  float* cls_scores;      // [B*heads][N] raw scores of the CLS query (the last layer's attention map) or null
  bf16* qkv_cls;          // [B][3*heads*64]: q, k, v of the CLS token (cls_combine_kernel reads them)
};

constexpr size_t smem_bytes() {
  return 1024 + (size_t)kKB * kAKbBytes + (size_t)kBStages * kBStageBytes + 3 * (size_t)kTileBytes + 6144;
}

__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(ptx::smem_u32(bar) & 0xFEFFFFFFu), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------
// 16 query rows [q_row0, +16) of the q tile against KT*16 key rows [key_row0, ...) of the k / v tiles, plus the CLS key
// at row `cls_row` handled apart (its score from 4 extra mma with k_cls broadcast over the B columns, its value as a
// rank-1 update): attention_mma.cuh's attend_mtile_cls_apart with free row offsets.
//   m0 / m1: bit k set <=> key_local k (0..KT*16-1) is allowed for this lane's query rows g / g + 8 (g = lane >> 2);
//   the CLS key is always allowed (:252-253 pads it True).
// The normalised output rows (bf16) are written over the q rows.
// ---------------------------------------------------------------------------------------------------
template <int KT>
__device__ __forceinline__ void attend_rows_cls_apart(uint8_t* qs, uint8_t* ks, uint8_t* vs, int q_row0, int key_row0,
                                                      int cls_row, int lane, uint32_t m0, uint32_t m1) {
  using namespace attn;
  constexpr int NT = KT * 2;
  const int g = lane >> 2, t = lane & 3;
  float s[NT][4], sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    ldmatrix_x4(a, smem_u32(tile_ptr(qs, q_row0 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4))));
#pragma unroll
    for (int jp = 0; jp < KT; ++jp) {
      uint32_t b[4];
      ldmatrix_x4(b, smem_u32(tile_ptr(ks, key_row0 + jp * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1))));
      mma_bf16(s[jp * 2], a, b[0], b[1]);
      mma_bf16(s[jp * 2 + 1], a, b[2], b[3]);
    }
    const uint32_t kb0 = *reinterpret_cast<const uint32_t*>(tile_ptr(ks, cls_row, kk * 2) + t * 4);
    const uint32_t kb1 = *reinterpret_cast<const uint32_t*>(tile_ptr(ks, cls_row, kk * 2 + 1) + t * 4);
    mma_bf16(sc, a, kb0, kb1);
  }
  float mx0 = sc[0], mx1 = sc[2];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int key = j * 8 + t * 2 + e;
      if (!((m0 >> key) & 1u)) s[j][e] = -FLT_MAX;
      if (!((m1 >> key) & 1u)) s[j][2 + e] = -FLT_MAX;
      mx0 = fmaxf(mx0, s[j][e]);
      mx1 = fmaxf(mx1, s[j][2 + e]);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[j][e] = __expf(s[j][e] - mx0);           // masked entries (-FLT_MAX) underflow to exactly 0
      s[j][2 + e] = __expf(s[j][2 + e] - mx1);
      sum0 += s[j][e];
      sum1 += s[j][2 + e];
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float ec0 = __expf(sc[0] - mx0), ec1 = __expf(sc[2] - mx1);
  const float inv0 = 1.0f / (sum0 + ec0), inv1 = 1.0f / (sum1 + ec1);
  const float pc0 = __bfloat162float(__float2bfloat16_rn(ec0 * inv0)), pc1 = __bfloat162float(__float2bfloat16_rn(ec1 * inv1));
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(tile_ptr(vs, cls_row, j) + t * 4);
    const float v0 = __uint_as_float(v << 16), v1 = __uint_as_float(v & 0xffff0000u);
    o[j][0] = pc0 * v0; o[j][1] = pc0 * v1; o[j][2] = pc1 * v0; o[j][3] = pc1 * v1;
  }
#pragma unroll
  for (int kk = 0; kk < KT; ++kk) {
    uint32_t a[4];
    a[0] = pack2(s[2 * kk][0] * inv0, s[2 * kk][1] * inv0);
    a[1] = pack2(s[2 * kk][2] * inv1, s[2 * kk][3] * inv1);
    a[2] = pack2(s[2 * kk + 1][0] * inv0, s[2 * kk + 1][1] * inv0);
    a[3] = pack2(s[2 * kk + 1][2] * inv1, s[2 * kk + 1][3] * inv1);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, smem_u32(tile_ptr(vs, key_row0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4))));
      mma_bf16(o[dp * 2], a, b[0], b[1]);
      mma_bf16(o[dp * 2 + 1], a, b[2], b[3]);
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<uint32_t*>(tile_ptr(qs, q_row0 + g, j) + t * 4) = pack2(o[j][0], o[j][1]);
    *reinterpret_cast<uint32_t*>(tile_ptr(qs, q_row0 + g + 8, j) + t * 4) = pack2(o[j][2], o[j][3]);
  }
  __syncwarp();
}

// CLS-query partial (m, l, o[64]) over the nk keys at rows [key_row0, key_row0 + nk) of the k / v tiles:
// attention_mma.cuh's cls_partial_warp with a free row offset and the CLS query in plain shared memory (q0s, 64 bf16).
template <int MTK, typename Valid, typename Token>
__device__ __forceinline__ void cls_partial_rows(uint8_t* ks, uint8_t* vs, const bf16* q0s, float* es, int key_row0, int nk,
                                                 int max_row, int lane, Valid valid, Token token_of, float* part,
                                                 float* scores) {
  using namespace attn;
  const int g = lane >> 2, t = lane & 3;
  uint32_t qb[4][2];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    qb[kk][0] = *reinterpret_cast<const uint32_t*>(q0s + kk * 16 + 2 * t);
    qb[kk][1] = *reinterpret_cast<const uint32_t*>(q0s + kk * 16 + 2 * t + 8);
  }
  float s[MTK][2];
  float m = -FLT_MAX;
#pragma unroll
  for (int mt = 0; mt < MTK; ++mt) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int row = min(key_row0 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, max_row);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      ldmatrix_x4(a, smem_u32(tile_ptr(ks, row, kk * 2 + (lane >> 4))));
      mma_bf16(acc, a, qb[kk][0], qb[kk][1]);
    }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int k = mt * 16 + g + hr * 8;
      float v = -FLT_MAX;
      if (k < nk) {
        if (valid(k)) v = acc[hr * 2];
        if (scores && t == 0) scores[token_of(k)] = v;
      }
      s[mt][hr] = v;
      m = fmaxf(m, v);
    }
  }
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float l = 0.f;
#pragma unroll
  for (int mt = 0; mt < MTK; ++mt)
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const float e = s[mt][hr] > -FLT_MAX ? __expf(s[mt][hr] - m) : 0.f;
      l += e;
      if (t == 0) es[mt * 16 + g + hr * 8] = e;
    }
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  __syncwarp();
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < MTK; ++kk) {
    uint32_t a[4];
    a[0] = a[1] = pack2(es[kk * 16 + 2 * t], es[kk * 16 + 2 * t + 1]);
    a[2] = a[3] = pack2(es[kk * 16 + 2 * t + 8], es[kk * 16 + 2 * t + 9]);
    const int row = min(key_row0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, max_row);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, smem_u32(tile_ptr(vs, row, dp * 2 + (lane >> 4))));
      mma_bf16(o[dp * 2], a, b[0], b[1]);
      mma_bf16(o[dp * 2 + 1], a, b[2], b[3]);
    }
  }
  if (lane == 0) { part[0] = m; part[1] = l; }
  if (g == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<float2*>(part + 2 + j * 8 + 2 * t) = make_float2(o[j][0], o[j][1]);
  }
  __syncwarp();
}

__device__ __forceinline__ void epi_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kEpiWarps * 32) : "memory"); }

// MODE: MT_ATTN_TIME / MT_ATTN_SPACE.  KT (TIME only): 16-key tiles per query tile (1: f <= 16, 2: f = 32).
template <int MODE, int KT>
__global__ void __launch_bounds__(kThreads, 1)
fused_attn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_cls,
                  const __grid_constant__ CUtensorMap tmap_w, Params p) {
  extern __shared__ __align__(1024) uint8_t fa_raw[];
  uint8_t* sm = fa_raw + ((1024u - (ptx::smem_u32(fa_raw) & 1023u)) & 1023u);
  uint8_t* a_buf = sm;                                        // [kKB][128 rows][128 B]
  uint8_t* b_buf = a_buf + kKB * kAKbBytes;                   // [kBStages][96 rows][128 B]
  uint8_t* q_t = b_buf + kBStages * kBStageBytes;             // q / k / v tiles of the current (tile, head)
  uint8_t* k_t = q_t + kTileBytes;
  uint8_t* v_t = k_t + kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_t + kTileBytes);
  uint64_t* a_full = bars;                                    // [kKB]  (leader)
  uint64_t* a_empty = a_full + kKB;                           // [kKB]  (both, multicast commit)
  uint64_t* b_full = a_empty + kKB;                           // [kBStages] (leader)
  uint64_t* b_empty = b_full + kBStages;                      // [kBStages] (both)
  uint64_t* acc_full = b_empty + kBStages;                    // [2] (both)
  uint64_t* acc_empty = acc_full + 2;                         // [2] (leader; every epilogue thread of the pair arrives)
  uint8_t* misc = reinterpret_cast<uint8_t*>(bars) + 256;                                     // (26 barriers = 208 B)
  bf16* q0s = reinterpret_cast<bf16*>(misc);                                                  // [64] CLS query, 16-B aligned
  uint32_t* allow_bits = reinterpret_cast<uint32_t*>(misc + 256);                             // [32] TIME: per query frame, bit k = key frame k
  uint8_t* frame_ok = misc + 512;                                                             // [64]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(misc + 576);
  float* es_all = reinterpret_cast<float*>(misc + 768);                                       // [kEpiWarps][64]
  static_assert(768 + kEpiWarps * 64 * 4 + 256 <= 6144, "misc region");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int s0 = (int)((long long)p.total_steps * pair / num_pairs);
  const int s1 = (int)((long long)p.total_steps * (pair + 1) / num_pairs);

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_cls);
    ptx::prefetch_tmap(&tmap_w);
    for (int i = 0; i < kKB; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < kBStages; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 2 * kEpiWarps); }   // one arrival per epilogue warp of the pair
    ptx::fence_mbar_init();
  }
  // rows of the A tile no TMA box ever writes must be finite (they are multiplied into accumulator rows nobody reads,
  // but whose q/k/v land in the tiles next to real keys): zero the whole A region once, before any TMA is issued
  for (int i = threadIdx.x; i < kKB * kAKbBytes / 16; i += kThreads) reinterpret_cast<uint4*>(a_buf)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async_smem();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tmem_alloc_2sm(tmem_ptr_smem, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  auto my_tile_of = [&](int step) { return min(2 * (step / p.heads) + (int)rank, p.n_tiles - 1); };

  if (warp == kAWarp) {
    // ===================================================================== A producer
    if (lane == 0) {
      int gen = 0;
      for (int step = s0; step < s1; ++gen) {
        const int h = step % p.heads;
        const int tile = my_tile_of(step);
        const int b = tile / p.tiles_per_video, t = tile - b * p.tiles_per_video;
        for (int kb = 0; kb < kKB; ++kb) {
          ptx::mbar_wait(&a_empty[kb], (gen & 1) ^ 1);
          if (leader) ptx::mbar_arrive_expect_tx(&a_full[kb], 2u * (uint32_t)p.a_bytes_kb);
          uint8_t* dst = a_buf + kb * kAKbBytes;
          if (MODE == MT_ATTN_TIME) {
            for (int g = 0; g < p.pt; ++g)           // (patch positions past n are zero-filled by the TMA unit)
              tma_load_4d_2sm(dst + g * p.f * 128, &tmap_a, &a_full[kb], kb * 64, t * p.pt + g, 0, b);
            ptx::tma_load_2d_2sm(dst + p.cls_row * 128, &tmap_cls, &a_full[kb], kb * 64, b * p.N);
          } else {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              tma_load_4d_2sm(dst + s * 64 * 128, &tmap_a, &a_full[kb], kb * 64, 0, t * 2 + s, b);
              ptx::tma_load_2d_2sm(dst + (s * 64 + kSpaceClsRow) * 128, &tmap_cls, &a_full[kb], kb * 64, b * p.N);
            }
          }
        }
        step += min(p.heads - h, s1 - step);
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // ===================================================================== W producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int step = s0; step < s1; ++step) {
        const int h = step % p.heads;
        for (int kb = 0; kb < kKB; ++kb) {
          ptx::mbar_wait(&b_empty[stage], phase ^ 1);
          if (leader) ptx::mbar_arrive_expect_tx(&b_full[stage], 2u * kBStageBytes);
          ptx::tma_load_2d_2sm(b_buf + stage * kBStageBytes, &tmap_w, &b_full[stage], kb * 64, h * 192 + (int)rank * kBRows);
          if (++stage == kBStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA)
    if (leader && lane == 0) {
      const uint32_t idesc = ptx::umma_idesc_bf16_f32(256, 192);
      int stage = 0, gen = 0, it = 0;
      uint32_t phase = 0;
      for (int step = s0; step < s1; ++step, ++it) {
        const int h = step % p.heads;
        const bool first = step == s0 || h == 0, last = h == p.heads - 1 || step == s1 - 1;
        const int as = it & 1;
        ptx::mbar_wait(&acc_empty[as], ((it >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * kAccStride);
        for (int kb = 0; kb < kKB; ++kb) {
          if (first) ptx::mbar_wait(&a_full[kb], gen & 1);
          ptx::mbar_wait(&b_full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(a_buf + kb * kAKbBytes);
          const uint32_t b_addr = ptx::smem_u32(b_buf + stage * kBStageBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ss_2sm(tmem_d, ptx::umma_smem_desc_sw128(a_addr + k * 32), ptx::umma_smem_desc_sw128(b_addr + k * 32),
                                  idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit_2sm(&b_empty[stage], 0x3);
          if (last) ptx::umma_commit_2sm(&a_empty[kb], 0x3);   // the A k-block is free for the next tile in both CTAs
          if (++stage == kBStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_2sm(&acc_full[as], 0x3);
        if (last) ++gen;
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue: drain + attention core
    const int ew = warp - 2;
    const int quad = warp & 3, part = ew >> 2;     // `part`: which kDrainCols-wide slice of q | k | v this warp drains
    const int row = quad * 32 + lane;
    const int etid = ew * 32 + lane;
    const int inner = p.heads * 64;
    float* es = es_all + ew * 64;
    int it = 0;
    for (int step = s0; step < s1; ++step, ++it) {
      const int h = step % p.heads;
      const bool first = step == s0 || h == 0;
      const int as = it & 1;
      const int tile = my_tile_of(step);
      const int b = tile / p.tiles_per_video, t = tile - b * p.tiles_per_video;
      const int bh = b * p.heads + h;
      epi_bar(1);                                  // everyone is done with the previous step's tiles and masks
      if (first && etid < 64) {
        frame_ok[etid] = etid < p.f ? p.mask[b * p.f + etid] : 0;
        if (MODE == MT_ATTN_TIME && etid < 32) {
          uint32_t bits = 0u;
          if (etid < p.f)
            for (int k = 0; k < p.f; ++k)
              if (p.mask[b * p.f + k] && p.idmask[((size_t)b * p.f + etid) * p.f + k]) bits |= 1u << k;
          allow_bits[etid] = bits;
        }
      }
      ptx::mbar_wait(&acc_full[as], (it >> 1) & 1);
      ptx::tc_fence_after();
      // ---- drain: this thread's row, columns [part*kDrainCols, +kDrainCols) of q | k | v -> bf16 -> swizzled tiles
      {
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * kAccStride + part * kDrainCols);
#pragma unroll
        for (int u = 0; u < kDrainCols / 16; ++u) {
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(taddr + u * 16, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c8 = 0; c8 < 2; ++c8) {
            const int col = part * kDrainCols + u * 16 + c8 * 8;  // 0..191: q 0-63, k 64-127, v 128-191
            uint8_t* tl = col < 64 ? q_t : (col < 128 ? k_t : v_t);
            const uint4 pk = make_uint4(attn::pack2(__uint_as_float(r[c8 * 8]), __uint_as_float(r[c8 * 8 + 1])),
                                        attn::pack2(__uint_as_float(r[c8 * 8 + 2]), __uint_as_float(r[c8 * 8 + 3])),
                                        attn::pack2(__uint_as_float(r[c8 * 8 + 4]), __uint_as_float(r[c8 * 8 + 5])),
                                        attn::pack2(__uint_as_float(r[c8 * 8 + 6]), __uint_as_float(r[c8 * 8 + 7])));
            *reinterpret_cast<uint4*>(attn::tile_ptr(tl, row, (col & 63) >> 3)) = pk;
            // the CLS query also goes to plain shared memory: attention tasks overwrite q rows with their outputs
            if (col < 64) {
              if (MODE == MT_ATTN_TIME) {
                if (row == p.cls_row) *reinterpret_cast<uint4*>(q0s + col) = pk;
              } else {
                if (row == kSpaceClsRow) *reinterpret_cast<uint4*>(q0s + col) = pk;
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(&acc_empty[as], 0);   // accumulator buffer back to the leader's MMA warp
      epi_bar(2);                                    // q / k / v tiles and masks complete

      if (MODE == MT_ATTN_TIME) {
        const int f = p.f, lf = 31 - __clz(f), p0 = t * p.pt;
        const int n_groups = min(p.pt, p.n - p0);
        const int rows_used = n_groups * f;
        const int n_mt = (rows_used + 15) >> 4;
        // tasks: [0, n_mt) query tiles, then one CLS-query partial per group
        for (int task = ew; task < n_mt + n_groups; task += kEpiWarps) {
          if (task < n_mt) {
            const int q_row0 = task * 16;
            const int key_row0 = q_row0 & ~(KT * 16 - 1);
            // this lane's two query rows: the keys of row qr are the f rows of its own group (f is a power of two)
            uint32_t mrow[2];
#pragma unroll
            for (int hr = 0; hr < 2; ++hr) {
              const int qr = q_row0 + (lane >> 2) + hr * 8;
              const int gq = qr >> lf;
              mrow[hr] = qr < rows_used ? allow_bits[qr & (f - 1)] << (gq * f - key_row0) : 0u;
            }
            attend_rows_cls_apart<KT>(q_t, k_t, v_t, q_row0, key_row0, p.cls_row, lane, mrow[0], mrow[1]);
            for (int e = lane; e < 16 * 8; e += 32) {
              const int r = q_row0 + (e >> 3), c = e & 7;
              if (r < rows_used) {
                const int g = r >> lf, fr = r & (f - 1);
                *reinterpret_cast<uint4*>(p.out + ((size_t)b * p.N + 1 + fr * p.n + p0 + g) * inner + h * 64 + c * 8) =
                    *reinterpret_cast<const uint4*>(attn::tile_ptr(q_t, r, c));
              }
            }
          } else {
            const int g = task - n_mt, pp = p0 + g;
            cls_partial_rows<KT>(k_t, v_t, q0s, es, g * f, f, 127, lane, [&](int k) { return frame_ok[k] != 0; },
                                 [&](int k) { return 1 + k * p.n + pp; }, p.cls_parts + ((size_t)bh * p.n + pp) * attn::kClsStride,
                                 p.cls_scores ? p.cls_scores + (size_t)bh * p.N : nullptr);
          }
        }
        if (t == 0 && ew == kEpiWarps - 1 && lane < 24) {     // q, k, v of the CLS token for cls_combine_kernel
          const int m = lane >> 3, c = lane & 7;
          const uint4 v = m == 0 ? *reinterpret_cast<const uint4*>(q0s + c * 8)
                                 : *reinterpret_cast<const uint4*>(attn::tile_ptr(m == 1 ? k_t : v_t, p.cls_row, c));
          *reinterpret_cast<uint4*>(p.qkv_cls + (size_t)b * 3 * inner + m * inner + h * 64 + c * 8) = v;
        }
      } else {
        const int n = p.n, nk_mt = (n + 15) >> 4;
        // tasks: slot s in {0, 1}: nk_mt query tiles each, then one CLS-query partial per slot
        const int per_slot = nk_mt + 1;
        for (int task = ew; task < 2 * per_slot; task += kEpiWarps) {
          const int s = task / per_slot, tt = task - s * per_slot;
          const int fr = t * 2 + s;
          if (fr >= p.f) continue;
          uint8_t* qs = q_t + s * 64 * 128;
          uint8_t* ks = k_t + s * 64 * 128;
          uint8_t* vs = v_t + s * 64 * 128;
          const int tok0 = 1 + fr * n;
          if (tt < nk_mt) {
            attn::attend_mtile<4>(qs, ks, vs, tt * 16, lane, [&](int, int key) { return key < n || key == kSpaceClsRow; });
            for (int e = lane; e < 16 * 8; e += 32) {
              const int r = tt * 16 + (e >> 3), c = e & 7;
              if (r < n)
                *reinterpret_cast<uint4*>(p.out + ((size_t)b * p.N + tok0 + r) * inner + h * 64 + c * 8) =
                    *reinterpret_cast<const uint4*>(attn::tile_ptr(qs, r, c));
            }
          } else {
            const bool ok = frame_ok[fr] != 0;
            cls_partial_rows<4>(ks, vs, q0s, es, 0, n, 63, lane, [&](int) { return ok; }, [&](int k) { return tok0 + k; },
                                p.cls_parts + ((size_t)bh * p.f + fr) * attn::kClsStride,
                                p.cls_scores ? p.cls_scores + (size_t)bh * p.N : nullptr);
          }
        }
        if (t == 0 && ew == kEpiWarps - 1 && lane < 24) {
          const int m = lane >> 3, c = lane & 7;
          const uint4 v = m == 0 ? *reinterpret_cast<const uint4*>(q0s + c * 8)
                                 : *reinterpret_cast<const uint4*>(attn::tile_ptr(m == 1 ? k_t : v_t, kSpaceClsRow, c));
          *reinterpret_cast<uint4*>(p.qkv_cls + (size_t)b * 3 * inner + m * inner + h * 64 + c * 8) = v;
        }
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace fattn
}  // namespace mt
