// Tensor-core backward of the divided attention groups (bf16 path), reference size_invariant_timesformer.py:122-135.
//
// Same shape of problem as the forward (attention_mma.cuh): groups of f x (f+1) or n x (n+1) with head dim 64, far too
// small for a tcgen05 tile, so the five products of the backward run on warp-level mma.sync m16n8k16:
//   S = Q K^T, dP = dO V^T          (per 16-query m-tile, accumulators in registers)
//   P = softmax(S + mask), D = rowsum(P o dP), dS = P o (dP - D)
//   dQ = dS K                       (dS re-used from registers as the A operand)
//   dK = dS^T Q, dV = P^T dO        (contraction over the queries: P^T / dS^T go through shared memory as bf16 tiles
//                                    [key][query], written transposed from the accumulator fragments)
// The fragment / ldmatrix address patterns are the ones of attend_mtile (A from a row-major tile, B = rows of K for
// Q K^T, B = .trans rows of V for P V); tiles are 128-byte rows with the 16-byte chunks XOR-swizzled by (row & 7).
// The CLS query's contribution to dK / dV of the group's keys comes in through ws_kv as (dS, p) per key and is rebuilt
// as rank-1 products with q_cls / dO_cls (attn_cls_bwd_kernel), the group's
// contribution to dK / dV of the CLS key leaves through ws_cls (attn_cls_finish_kernel sums them): see train.cu.
#pragma once
#include <float.h>

#include "attention_mma.cuh"

namespace mt {
namespace attn {

// MODE: MT_ATTN_TIME / MT_ATTN_SPACE.  MT = query m-tiles per group, NKT = 16-key tiles per group (CLS = key 0),
// GPB = groups per block; MT * GPB = 4 warps.
template <int MODE, int MT, int NKT, int GPB>
__global__ void __launch_bounds__(128) attn_group_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                 const uint8_t* __restrict__ mask,
                                                                 const uint8_t* __restrict__ idmask, bf16* __restrict__ dqkv,
                                                                 const float* __restrict__ ws_kv, float* __restrict__ ws_cls,
                                                                 int f, int n, int heads, int total_groups) {
  static_assert(MT * GPB == 4, "four warps per block");
  extern __shared__ __align__(1024) uint8_t dsm[];
  constexpr int kQB = MT * 16 * 128, kKB = NKT * 16 * 128;
  constexpr int kSlot = 2 * kQB + 4 * kKB;
  constexpr int NT = NKT * 2;
  __shared__ unsigned long long allow_bits[GPB][MT * 16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gi = warp / MT, mi = warp % MT;
  uint8_t* slot = dsm + gi * kSlot;
  uint8_t* qs = slot;                 // [MT*16][64]  Q
  uint8_t* os = qs + kQB;             // [MT*16][64]  dO
  uint8_t* ks = os + kQB;             // [NKT*16][64] K   (row 0 = CLS)
  uint8_t* vs = ks + kKB;             // [NKT*16][64] V
  uint8_t* pt = vs + kKB;             // [NKT*16][64] P^T   (columns = queries)
  uint8_t* st = pt + kKB;             // [NKT*16][64] dS^T
  const int G = MODE == MT_ATTN_TIME ? n : f;
  const int Gq = MODE == MT_ATTN_TIME ? f : n;
  const int Gk = Gq + 1;
  const int gidx = blockIdx.x * GPB + gi;
  const bool valid = gidx < total_groups;
  const int g = valid ? gidx % G : 0;
  const int h = valid ? (gidx / G) % heads : 0;
  const int b = valid ? gidx / (G * heads) : 0;
  const int N = 1 + f * n, inner = heads * 64, ld = 3 * inner;
  const bf16* base = qkv + (size_t)b * N * ld + h * 64;              // (token 0: base[0..63] is the CLS query)
  const bf16* dcls = dout + (size_t)b * N * inner + h * 64;           // dO of the CLS row
  auto token = [&](int j) -> int {   // j = 0 CLS, j >= 1 the (j-1)-th member of the group
    if (j == 0) return 0;
    return MODE == MT_ATTN_TIME ? 1 + (j - 1) * n + g : 1 + g * n + (j - 1);
  };
  // ---- stage the group's tiles (the MT warps of the group share the work)
  const int tg = mi * 32 + lane;
  for (int e = tg; e < MT * 16 * 8; e += MT * 32) {
    const int r = e >> 3, c = e & 7;
    if (valid && r < Gq) {
      const int tok = token(r + 1);
      cp_async16(tile_ptr(qs, r, c), base + (size_t)tok * ld + c * 8);
      cp_async16(tile_ptr(os, r, c), dout + ((size_t)b * N + tok) * inner + h * 64 + c * 8);
    } else {
      *reinterpret_cast<uint4*>(tile_ptr(qs, r, c)) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(tile_ptr(os, r, c)) = make_uint4(0, 0, 0, 0);
    }
  }
  for (int e = tg; e < NKT * 16 * 8; e += MT * 32) {
    const int r = e >> 3, c = e & 7;
    if (valid && r < Gk) {
      const bf16* kr = base + (size_t)token(r) * ld + c * 8;
      cp_async16(tile_ptr(ks, r, c), kr + inner);
      cp_async16(tile_ptr(vs, r, c), kr + 2 * inner);
    } else {
      *reinterpret_cast<uint4*>(tile_ptr(ks, r, c)) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(tile_ptr(vs, r, c)) = make_uint4(0, 0, 0, 0);
    }
  }
  if (lane < 16) {
    const int q = mi * 16 + lane;
    unsigned long long bits = 1ull;                        // the CLS key is always allowed (:254)
    if (valid && q < Gq) {
      for (int k = 1; k < Gk; ++k) {
        bool ok = true;
        if (MODE == MT_ATTN_TIME) ok = mask[b * f + (k - 1)] && idmask[((size_t)b * f + q) * f + (k - 1)];
        if (ok) bits |= 1ull << k;
      }
    }
    allow_bits[gi][q] = bits;
  }
  cp_async_wait_all();
  __syncthreads();

  const int gr = lane >> 2, t = lane & 3;
  const int q_row0 = mi * 16;
  // ---- phase 1: S = Q K^T and dP = dO V^T for this warp's 16 queries
  float s[NT][4], dp[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
  }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {              // 16 head dims per step
    uint32_t aq[4], ao[4];
    ldmatrix_x4(aq, smem_u32(tile_ptr(qs, q_row0 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4))));
    ldmatrix_x4(ao, smem_u32(tile_ptr(os, q_row0 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4))));
#pragma unroll
    for (int jp = 0; jp < NKT; ++jp) {
      uint32_t bk[4], bv[4];
      ldmatrix_x4(bk, smem_u32(tile_ptr(ks, jp * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1))));
      ldmatrix_x4(bv, smem_u32(tile_ptr(vs, jp * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1))));
      mma_bf16(s[jp * 2], aq, bk[0], bk[1]);
      mma_bf16(s[jp * 2 + 1], aq, bk[2], bk[3]);
      mma_bf16(dp[jp * 2], ao, bv[0], bv[1]);
      mma_bf16(dp[jp * 2 + 1], ao, bv[2], bv[3]);
    }
  }
  // masked softmax, rows gr (c0, c1) and gr + 8 (c2, c3)
  const unsigned long long al0 = allow_bits[gi][q_row0 + gr], al1 = allow_bits[gi][q_row0 + gr + 8];
  float mx0 = -FLT_MAX, mx1 = -FLT_MAX;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int key = j * 8 + t * 2 + e;
      if (!((al0 >> key) & 1ull)) s[j][e] = -FLT_MAX;
      if (!((al1 >> key) & 1ull)) s[j][2 + e] = -FLT_MAX;
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
      s[j][e] = __expf(s[j][e] - mx0);          // masked entries: exp(-FLT_MAX - mx) == 0
      s[j][2 + e] = __expf(s[j][2 + e] - mx1);
      sum0 += s[j][e];
      sum1 += s[j][2 + e];
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s[j][e] *= inv0;
      s[j][2 + e] *= inv1;
      d0 = fmaf(s[j][e], dp[j][e], d0);
      d1 = fmaf(s[j][2 + e], dp[j][2 + e], d1);
    }
  }
  d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
  d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
  // dS = P o (dP - D) (kept in dp); P^T and dS^T to shared memory: tile row = key, column = query
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      dp[j][e] = s[j][e] * (dp[j][e] - d0);
      dp[j][2 + e] = s[j][2 + e] * (dp[j][2 + e] - d1);
      const int key = j * 8 + t * 2 + e;
      const int qa = q_row0 + gr, qb = q_row0 + gr + 8;
      *reinterpret_cast<bf16*>(tile_ptr(pt, key, qa >> 3) + (qa & 7) * 2) = __float2bfloat16_rn(s[j][e]);
      *reinterpret_cast<bf16*>(tile_ptr(pt, key, qb >> 3) + (qb & 7) * 2) = __float2bfloat16_rn(s[j][2 + e]);
      *reinterpret_cast<bf16*>(tile_ptr(st, key, qa >> 3) + (qa & 7) * 2) = __float2bfloat16_rn(dp[j][e]);
      *reinterpret_cast<bf16*>(tile_ptr(st, key, qb >> 3) + (qb & 7) * 2) = __float2bfloat16_rn(dp[j][2 + e]);
    }
  }
  // ---- dQ = dS K  (this warp's 16 queries x 64 dims)
  {
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NKT; ++kk) {          // 16 keys per step
      uint32_t a[4];
      a[0] = pack2(dp[2 * kk][0], dp[2 * kk][1]);
      a[1] = pack2(dp[2 * kk][2], dp[2 * kk][3]);
      a[2] = pack2(dp[2 * kk + 1][0], dp[2 * kk + 1][1]);
      a[3] = pack2(dp[2 * kk + 1][2], dp[2 * kk + 1][3]);
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) {          // 16 head dims per ldmatrix.x4.trans
        uint32_t bb[4];
        ldmatrix_x4_trans(bb, smem_u32(tile_ptr(ks, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dd * 2 + (lane >> 4))));
        mma_bf16(o[dd * 2], a, bb[0], bb[1]);
        mma_bf16(o[dd * 2 + 1], a, bb[2], bb[3]);
      }
    }
    if (valid) {
      const int qa = q_row0 + gr, qb = q_row0 + gr + 8;
      if (qa < Gq) {
        bf16* row = dqkv + ((size_t)b * N + token(qa + 1)) * ld + h * 64 + t * 2;
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<uint32_t*>(row + j * 8) = pack2(o[j][0], o[j][1]);
      }
      if (qb < Gq) {
        bf16* row = dqkv + ((size_t)b * N + token(qb + 1)) * ld + h * 64 + t * 2;
#pragma unroll
        for (int j = 0; j < 8; ++j) *reinterpret_cast<uint32_t*>(row + j * 8) = pack2(o[j][2], o[j][3]);
      }
    }
  }
  __syncthreads();
  // ---- phase 2: dK = dS^T Q, dV = P^T dO for 16-key tiles, contraction over the MT * 16 queries
  const size_t bh = (size_t)b * heads + h;
  for (int kt = mi; kt < NKT; kt += MT) {
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
      dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
    }
#pragma unroll
    for (int qq = 0; qq < MT; ++qq) {           // 16 queries per step
      uint32_t as_[4], ap[4];
      ldmatrix_x4(as_, smem_u32(tile_ptr(st, kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, qq * 2 + (lane >> 4))));
      ldmatrix_x4(ap, smem_u32(tile_ptr(pt, kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, qq * 2 + (lane >> 4))));
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) {
        uint32_t bq[4], bo[4];
        ldmatrix_x4_trans(bq, smem_u32(tile_ptr(qs, qq * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dd * 2 + (lane >> 4))));
        ldmatrix_x4_trans(bo, smem_u32(tile_ptr(os, qq * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dd * 2 + (lane >> 4))));
        mma_bf16(dk[dd * 2], as_, bq[0], bq[1]);
        mma_bf16(dk[dd * 2 + 1], as_, bq[2], bq[3]);
        mma_bf16(dv[dd * 2], ap, bo[0], bo[1]);
        mma_bf16(dv[dd * 2 + 1], ap, bo[2], bo[3]);
      }
    }
    if (!valid) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int key = kt * 16 + gr + half * 8;
      if (key >= Gk) continue;
      if (key == 0) {
        float* w = ws_cls + (bh * G + g) * 128 + t * 2;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          *reinterpret_cast<float2*>(w + j * 8) = make_float2(dk[j][half * 2], dk[j][half * 2 + 1]);
          *reinterpret_cast<float2*>(w + 64 + j * 8) = make_float2(dv[j][half * 2], dv[j][half * 2 + 1]);
        }
      } else {
        const int tok = token(key);
        // the CLS query's rank-1 contribution: dS_cls,tok * q_cls and p_cls,tok * dO_cls
        const float2 sp = *reinterpret_cast<const float2*>(ws_kv + (bh * N + tok) * 2);
        bf16* row = dqkv + ((size_t)b * N + tok) * ld + h * 64 + t * 2;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 qc = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(base + t * 2 + j * 8));
          const float2 oc = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dcls + t * 2 + j * 8));
          *reinterpret_cast<uint32_t*>(row + inner + j * 8) =
              pack2(fmaf(sp.x, qc.x, dk[j][half * 2]), fmaf(sp.x, qc.y, dk[j][half * 2 + 1]));
          *reinterpret_cast<uint32_t*>(row + 2 * inner + j * 8) =
              pack2(fmaf(sp.y, oc.x, dv[j][half * 2]), fmaf(sp.y, oc.y, dv[j][half * 2 + 1]));
        }
      }
    }
  }
}

template <int MODE, int MT, int NKT, int GPB>
int launch_group_bwd_mma(const bf16* qkv, const bf16* dout, const uint8_t* mask, const uint8_t* idmask, bf16* dqkv,
                         const float* ws_kv, float* ws_cls, int B, int f, int n, int heads, cudaStream_t st) {
  constexpr int kSlot = 2 * MT * 16 * 128 + 4 * NKT * 16 * 128;
  const int G = MODE == MT_ATTN_TIME ? n : f;
  const int total = B * heads * G;
  auto kern = attn_group_bwd_mma_kernel<MODE, MT, NKT, GPB>;
  if (first_use_on_device(reinterpret_cast<const void*>(kern))) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GPB * kSlot);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(attn_group_bwd_mma)");
  }
  kern<<<(total + GPB - 1) / GPB, 128, GPB * kSlot, st>>>(qkv, dout, mask, idmask, dqkv, ws_kv, ws_cls, f, n, heads, total);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_status(e, "attn_group_bwd_mma_kernel");
  count_launch();
  return MT_OK;
}

// Returns MT_ERR_UNSUPPORTED when no instantiation covers (mode, f, n): the caller runs the SIMT kernel.
inline int launch_attn_group_bwd_mma(int mode, const bf16* qkv, const bf16* dout, const uint8_t* mask, const uint8_t* idmask,
                                     bf16* dqkv, const float* ws_kv, float* ws_cls, int B, int f, int n, int heads,
                                     cudaStream_t st) {
#define MT_BWD_ARGS qkv, dout, mask, idmask, dqkv, ws_kv, ws_cls, B, f, n, heads, st
  if (mode == MT_ATTN_TIME) {
    if (f <= 15) return launch_group_bwd_mma<MT_ATTN_TIME, 1, 1, 4>(MT_BWD_ARGS);
    if (f == 16) return launch_group_bwd_mma<MT_ATTN_TIME, 1, 2, 4>(MT_BWD_ARGS);
    if (f <= 31) return launch_group_bwd_mma<MT_ATTN_TIME, 2, 2, 2>(MT_BWD_ARGS);
    if (f <= 32) return launch_group_bwd_mma<MT_ATTN_TIME, 2, 3, 2>(MT_BWD_ARGS);
    return MT_ERR_UNSUPPORTED;
  }
  if (n <= 15) return launch_group_bwd_mma<MT_ATTN_SPACE, 1, 1, 4>(MT_BWD_ARGS);
  if (n >= 48 && n <= 63) return launch_group_bwd_mma<MT_ATTN_SPACE, 4, 4, 1>(MT_BWD_ARGS);
  return MT_ERR_UNSUPPORTED;
#undef MT_BWD_ARGS
}

}  // namespace attn
}  // namespace mt
