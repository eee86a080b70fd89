// Tensor-core attention core for the bf16 path (reference size_invariant_timesformer.py:122-135).
//
// The divided attention groups are tiny -- time: f queries x (f+1) keys, space: n x (n+1), head dim 64 --
// 3 % of an attention block's FLOPs, far below one 128-row tcgen05 tile, so the core runs on warp-level
// mma.sync m16n8k16 (bf16 in, fp32 accumulate) with the softmax held in registers between the two
// products (QK^T accumulator fragments are re-used as the A operand of PV).  The kernels are bound by
// streaming qkv from HBM/L2 (cp.async into swizzled shared memory), not by math.
#pragma once
#include <float.h>

#include "common.cuh"

namespace mt {
namespace attn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D(16x8, f32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// tile of 64-wide bf16 rows (128 B per row), 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint8_t* tile_ptr(uint8_t* base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}

// One 16-query m-tile against NKT*16 (padded) keys held in shared memory.
//   allow(q_local, key) -> bool decides the mask; keys >= n_keys must be rejected by the caller's allow().
// Writes the normalised output rows (bf16) back over the Q tile rows (own rows only), 16 x 64.
template <int NKT, typename Allow>
__device__ __forceinline__ void attend_mtile(uint8_t* qs, uint8_t* ks, uint8_t* vs, int q_row0, int lane, Allow allow) {
  constexpr int NT = NKT * 2;                    // 8-key n-tiles of S
  float s[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
  // ---- S = Q K^T
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {              // 16 head dims per step
    uint32_t a[4];
    ldmatrix_x4(a, smem_u32(tile_ptr(qs, q_row0 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4))));
#pragma unroll
    for (int jp = 0; jp < NKT; ++jp) {          // 16 keys per ldmatrix.x4
      uint32_t b[4];
      ldmatrix_x4(b, smem_u32(tile_ptr(ks, jp * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1))));
      mma_bf16(s[jp * 2], a, b[0], b[1]);
      mma_bf16(s[jp * 2 + 1], a, b[2], b[3]);
    }
  }
  // ---- masked softmax over keys, rows g (c0,c1) and g+8 (c2,c3)
  const int g = lane >> 2, t = lane & 3;
  float mx0 = -FLT_MAX, mx1 = -FLT_MAX;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int key = j * 8 + t * 2 + e;
      if (!allow(g, key)) s[j][e] = -FLT_MAX;
      if (!allow(g + 8, key)) s[j][2 + e] = -FLT_MAX;
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
      // masked entries hold -FLT_MAX: exp underflows to exactly 0 (the CLS key keeps every row's max finite)
      s[j][e] = __expf(s[j][e] - mx0);
      s[j][2 + e] = __expf(s[j][2 + e] - mx1);
      sum0 += s[j][e];
      sum1 += s[j][2 + e];
    }
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  // ---- O = P V
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NKT; ++kk) {            // 16 keys per step
    uint32_t a[4];
    a[0] = pack2(s[2 * kk][0] * inv0, s[2 * kk][1] * inv0);
    a[1] = pack2(s[2 * kk][2] * inv1, s[2 * kk][3] * inv1);
    a[2] = pack2(s[2 * kk + 1][0] * inv0, s[2 * kk + 1][1] * inv0);
    a[3] = pack2(s[2 * kk + 1][2] * inv1, s[2 * kk + 1][3] * inv1);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {            // 16 head dims per ldmatrix.x4.trans
      uint32_t b[4];
      ldmatrix_x4_trans(b, smem_u32(tile_ptr(vs, kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4))));
      mma_bf16(o[dp * 2], a, b[0], b[1]);
      mma_bf16(o[dp * 2 + 1], a, b[2], b[3]);
    }
  }
  // ---- stage O (bf16) over this m-tile's own Q rows so the caller can write full 128-byte rows
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    *reinterpret_cast<uint32_t*>(tile_ptr(qs, q_row0 + g, j) + t * 4) = pack2(o[j][0], o[j][1]);
    *reinterpret_cast<uint32_t*>(tile_ptr(qs, q_row0 + g + 8, j) + t * 4) = pack2(o[j][2], o[j][3]);
  }
  __syncwarp();
}

// One 16-query m-tile against KT*16 patch keys (tile rows 1 .. KT*16) PLUS the CLS key (tile row 0) handled apart:
// with f = 16 frames the 17 keys of a time group would otherwise pad to two 16-key tiles (32 mma per group); here
// the frame keys are exactly KT tiles (QK^T and PV on mma.sync), the CLS key's score comes from KT-independent 4 mma
// with k_cls broadcast over the B columns, and its value row is added as a rank-1 update in registers (20 mma).
//   allow(q_local, key) as in attend_mtile: key 0 = CLS (always allowed by the callers), key k+1 = frame k.
template <int KT, typename Allow>
__device__ __forceinline__ void attend_mtile_cls_apart(uint8_t* qs, uint8_t* ks, uint8_t* vs, int q_row0, int lane, Allow allow) {
  constexpr int NT = KT * 2;
  const int g = lane >> 2, t = lane & 3;
  float s[NT][4], sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {              // 16 head dims per step
    uint32_t a[4];
    ldmatrix_x4(a, smem_u32(tile_ptr(qs, q_row0 + (lane & 7) + ((lane >> 3) & 1) * 8, kk * 2 + (lane >> 4))));
#pragma unroll
    for (int jp = 0; jp < KT; ++jp) {
      uint32_t b[4];
      ldmatrix_x4(b, smem_u32(tile_ptr(ks, 1 + jp * 16 + (lane & 7) + (lane >> 4) * 8, kk * 2 + ((lane >> 3) & 1))));
      mma_bf16(s[jp * 2], a, b[0], b[1]);
      mma_bf16(s[jp * 2 + 1], a, b[2], b[3]);
    }
    // CLS key: every B column = k_cls, so c0 / c2 of every lane = q_row . k_cls for rows g / g + 8
    const uint32_t kb0 = *reinterpret_cast<const uint32_t*>(tile_ptr(ks, 0, kk * 2) + t * 4);
    const uint32_t kb1 = *reinterpret_cast<const uint32_t*>(tile_ptr(ks, 0, kk * 2 + 1) + t * 4);
    mma_bf16(sc, a, kb0, kb1);
  }
  float mx0 = sc[0], mx1 = sc[2];               // the CLS key is always allowed: it keeps every row's max finite
#pragma unroll
  for (int j = 0; j < NT; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int key = 1 + j * 8 + t * 2 + e;
      if (!allow(g, key)) s[j][e] = -FLT_MAX;
      if (!allow(g + 8, key)) s[j][2 + e] = -FLT_MAX;
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
  const float ec0 = __expf(sc[0] - mx0), ec1 = __expf(sc[2] - mx1);   // CLS key, counted once per row
  const float inv0 = 1.0f / (sum0 + ec0), inv1 = 1.0f / (sum1 + ec1);
  // P in bf16 for the tensor-core product; the CLS probability goes through bf16 as well so that every key is
  // weighted at the same precision
  const float pc0 = __bfloat162float(__float2bfloat16_rn(ec0 * inv0)), pc1 = __bfloat162float(__float2bfloat16_rn(ec1 * inv1));
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {                 // rank-1 start: o = p_cls * v_cls
    const uint32_t v = *reinterpret_cast<const uint32_t*>(tile_ptr(vs, 0, j) + t * 4);
    const float v0 = __uint_as_float(v << 16), v1 = __uint_as_float(v & 0xffff0000u);
    o[j][0] = pc0 * v0; o[j][1] = pc0 * v1; o[j][2] = pc1 * v0; o[j][3] = pc1 * v1;
  }
#pragma unroll
  for (int kk = 0; kk < KT; ++kk) {             // 16 frame keys per step
    uint32_t a[4];
    a[0] = pack2(s[2 * kk][0] * inv0, s[2 * kk][1] * inv0);
    a[1] = pack2(s[2 * kk][2] * inv1, s[2 * kk][3] * inv1);
    a[2] = pack2(s[2 * kk + 1][0] * inv0, s[2 * kk + 1][1] * inv0);
    a[3] = pack2(s[2 * kk + 1][2] * inv1, s[2 * kk + 1][3] * inv1);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, smem_u32(tile_ptr(vs, 1 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4))));
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

// ---------------------------------------------------------------------------------------------------
// CLS row of Attention.forward (:117-120): query 0 of each (b, h) attends ALL N keys with the padded-frame
// mask (cls_attn_mask :258-260).  The grouped kernels already hold every patch key/value of their group in
// shared memory, so ONE extra warp per group computes the CLS query's flash-style partial over those keys
//   m = max_k s_k,  l = sum_k exp(s_k - m),  o[d] = sum_k exp(s_k - m) v[k][d]       (s_k = q0 . k_k)
// (and the raw scores for the attention map) and cls_combine_kernel merges the groups + the CLS key itself:
// qkv is not streamed a second time for the CLS row.  Keys = rows 1 .. nk of the ks / vs tiles.
// ---------------------------------------------------------------------------------------------------
constexpr int kClsStride = 66;                   // floats per partial: m, l, o[64]

// MTK = 16-key tiles covering the nk keys; `es` = 16*MTK floats of per-warp scratch; q0 = the CLS query (global,
// bf16, 64 values); rows beyond `max_row` are never addressed (rows past the keys are zero-filled tiles).
// Both products run on mma.sync: scores = K q0 (K rows as the A operand, q0 broadcast over the 8 B columns),
// o = e^T V (e broadcast over the 16 A rows, V through ldmatrix.trans exactly as in attend_mtile).
template <int MTK, typename Valid, typename Token>
__device__ __forceinline__ void cls_partial_warp(uint8_t* ks, uint8_t* vs, const bf16* q0, float* es, int nk, int max_row,
                                                 int lane, Valid valid, Token token_of, float* part, float* scores) {
  const int g = lane >> 2, t = lane & 3;
  uint32_t qb[4][2];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    qb[kk][0] = *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 2 * t);
    qb[kk][1] = *reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 2 * t + 8);
  }
  float s[MTK][2];
  float m = -FLT_MAX;
#pragma unroll
  for (int mt = 0; mt < MTK; ++mt) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int row = min(1 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, max_row);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      ldmatrix_x4(a, smem_u32(tile_ptr(ks, row, kk * 2 + (lane >> 4))));
      mma_bf16(acc, a, qb[kk][0], qb[kk][1]);
    }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {             // this lane: keys mt*16 + g (c0) and + 8 (c2)
      const int k = mt * 16 + g + hr * 8;
      float v = -FLT_MAX;
      if (k < nk) {
        if (valid(k)) v = acc[hr * 2];           // masked_fill(~mask, -finfo.max) (:83-84)
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
      const float e = s[mt][hr] > -FLT_MAX ? __expf(s[mt][hr] - m) : 0.f;   // (all keys masked: l = 0)
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
    const int row = min(1 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, max_row);
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
}

// Merge of the per-group partials with the CLS key itself; one block (64 threads) per (b, h).
// Writes the CLS row of `out` and, when `cls_attn` is given (it holds the raw scores of the patch keys),
// normalises it in place into the attention map the model returns (:271).
// `qkv_rows` = rows per video of the buffer `qkv` points at: N for the full qkv tensor, 1 for the fused kernel's
// compact [B][3*inner] copy of the CLS token's q, k, v.
static __global__ void __launch_bounds__(64) cls_combine_kernel(const bf16* __restrict__ qkv, int qkv_rows,
                                                                const float* __restrict__ parts, bf16* __restrict__ out,
                                                                float* __restrict__ cls_attn, int N, int G, int heads) {
  __shared__ float red[2];
  __shared__ __align__(16) float ps[64 * kClsStride];            // the (b, h)'s partials (G <= 63)
  const int bh = blockIdx.x, b = bh / heads, h = bh % heads;
  const int inner = heads * 64, ld = 3 * inner;
  const bf16* base = qkv + (size_t)b * qkv_rows * ld + h * 64;
  const int d = threadIdx.x;
  {                                              // all loads independent: one L2 round trip, not one per group
    const float2* src = reinterpret_cast<const float2*>(parts + (size_t)bh * G * kClsStride);
    float2* dst = reinterpret_cast<float2*>(ps);
    for (int i = d; i < G * (kClsStride / 2); i += 64) dst[i] = src[i];
  }
  float p = __bfloat162float(base[d]) * __bfloat162float(base[inner + d]);     // q0 . k0
  const float v0 = __bfloat162float(base[2 * inner + d]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  if ((d & 31) == 0) red[d >> 5] = p;
  __syncthreads();
  const float s0 = red[0] + red[1];
  float M = s0;
  for (int g = 0; g < G; ++g)
    if (ps[g * kClsStride + 1] > 0.f) M = fmaxf(M, ps[g * kClsStride]);
  const float e0 = expf(s0 - M);
  float L = e0, acc = e0 * v0;
  for (int g = 0; g < G; ++g) {                  // fixed order: deterministic
    const float lg = ps[g * kClsStride + 1];
    if (lg > 0.f) {
      const float w = expf(ps[g * kClsStride] - M);
      L = fmaf(w, lg, L);
      acc = fmaf(w, ps[g * kClsStride + 2 + d], acc);
    }
  }
  const float inv = 1.0f / L;
  out[(size_t)b * N * inner + h * 64 + d] = __float2bfloat16_rn(acc * inv);
  if (cls_attn) {
    float* row = cls_attn + (size_t)bh * N;
    for (int j = d; j < N; j += 64) {
      const float sj = j == 0 ? s0 : row[j];
      row[j] = sj > -FLT_MAX ? expf(sj - M) * inv : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// SPACE: one block (4 warps) per group (b, h, frame): 49 queries x (CLS + 49) keys, no mask (:266).
// ---------------------------------------------------------------------------------------------------
// A fifth warp computes the group's partial of the CLS row (cls_partial_warp).
static __global__ void __launch_bounds__(160) attn_space_mma_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ mask,
                                                             bf16* __restrict__ out, float* __restrict__ cls_parts,
                                                             float* __restrict__ cls_scores, int f, int n, int heads) {
  __shared__ __align__(1024) uint8_t sm[3 * 64 * 128];
  __shared__ float es[64];
  uint8_t* qs = sm;
  uint8_t* ks = sm + 64 * 128;
  uint8_t* vs = sm + 2 * 64 * 128;
  const int fr = blockIdx.x % f;
  const int h = (blockIdx.x / f) % heads;
  const int b = blockIdx.x / (f * heads);
  const int N = 1 + f * n, inner = heads * 64, ld = 3 * inner;
  const bf16* base = qkv + (size_t)b * N * ld + h * 64;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tok0 = 1 + fr * n;                   // first patch token of the frame
  for (int e = tid; e < 64 * 8; e += 160) {
    const int r = e >> 3, c = e & 7;
    // queries: row r = patch r; keys/values: row 0 = CLS, row r = patch r-1
    if (r < n) cp_async16(tile_ptr(qs, r, c), base + (size_t)(tok0 + r) * ld + c * 8);
    else *reinterpret_cast<uint4*>(tile_ptr(qs, r, c)) = make_uint4(0, 0, 0, 0);
    if (r <= n) {
      const bf16* kr = base + (size_t)(r == 0 ? 0 : tok0 + r - 1) * ld + c * 8;
      cp_async16(tile_ptr(ks, r, c), kr + inner);
      cp_async16(tile_ptr(vs, r, c), kr + 2 * inner);
    } else {
      *reinterpret_cast<uint4*>(tile_ptr(ks, r, c)) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(tile_ptr(vs, r, c)) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const int nk = n + 1;
  if (warp == 4) {
    const int bh = b * heads + h;
    const bool frame_ok = mask[b * f + fr] != 0;
    // (base = token 0's q for this head: the CLS query, already scaled by dim_head^-1/2 through the weights)
    cls_partial_warp<4>(ks, vs, base, es, n, 63, lane, [&](int) { return frame_ok; }, [&](int k) { return tok0 + k; },
                        cls_parts + ((size_t)bh * f + fr) * kClsStride, cls_scores ? cls_scores + (size_t)bh * N : nullptr);
  } else if (warp * 16 < n) {
    attend_mtile<4>(qs, ks, vs, warp * 16, lane, [&](int, int key) { return key < nk; });
    // rows warp*16 .. +15 of qs now hold O; 8 lanes write one 128-byte row
    for (int e = lane; e < 16 * 8; e += 32) {
      const int r = warp * 16 + (e >> 3), c = e & 7;
      if (r < n)
        *reinterpret_cast<uint4*>(out + ((size_t)b * N + tok0 + r) * inner + h * 64 + c * 8) =
            *reinterpret_cast<const uint4*>(tile_ptr(qs, r, c));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// TIME: one warp per group (b, h, patch): f queries x (CLS + f) keys with the identity mask (:252-255):
// key k allowed iff mask[b][k] & identities_mask[b][q][k]; the CLS key always.  4 groups per block.
// NKT = ceil((f + 1) / 16) key tiles, MT = ceil(f / 16) query tiles.
// ---------------------------------------------------------------------------------------------------
// Warps 4-7 compute the CLS-row partials of the groups of warps 0-3 (cls_partial_warp).
template <int NKT, int MT>
__global__ void __launch_bounds__(256) attn_time_mma_kernel(const bf16* __restrict__ qkv, const uint8_t* __restrict__ mask,
                                                            const uint8_t* __restrict__ idmask, bf16* __restrict__ out,
                                                            float* __restrict__ cls_parts, float* __restrict__ cls_scores,
                                                            int f, int n, int heads) {
  extern __shared__ __align__(1024) uint8_t dsm[];
  constexpr int kQBytes = MT * 16 * 128, kKBytes = NKT * 16 * 128;
  constexpr int kWarpBytes = kQBytes + 2 * kKBytes;
  __shared__ unsigned long long allow_bits[64];  // per query frame: bit k set <=> key k (0 = CLS) allowed
  __shared__ float es_all[4][16 * NKT];
  __shared__ uint8_t frame_ok[64];
  const int tid = threadIdx.x, lane = tid & 31;
  const bool cls_warp = tid >= 128;
  const int warp = (tid >> 5) & 3;               // group slot; warps 4-7 mirror warps 0-3
  // blocks are laid out per (b, h): ceil(n / 4) blocks each, one warp per patch position
  const int blocks_per_bh = (n + 3) / 4;
  const int bh = blockIdx.x / blocks_per_bh;
  const int p = (blockIdx.x % blocks_per_bh) * 4 + warp;
  const int b = bh / heads, h = bh % heads;
  if (tid >= 128 && tid < 192) {
    const int d = tid - 128;
    frame_ok[d] = d < f ? mask[b * f + d] : 0;
  }
  if (tid < 64) {
    unsigned long long bits = 1ull;              // CLS key
    if (tid < f)
      for (int k = 0; k < f; ++k)
        if (mask[b * f + k] && idmask[((size_t)b * f + tid) * f + k]) bits |= 1ull << (k + 1);
    allow_bits[tid] = bits;
  }
  uint8_t* qs = dsm + warp * kWarpBytes;
  uint8_t* ks = qs + kQBytes;
  uint8_t* vs = ks + kKBytes;
  const int N = 1 + f * n, inner = heads * 64, ld = 3 * inner;
  const bf16* base = qkv + (size_t)b * N * ld + h * 64;
  const bool active = p < n;
  if (active && !cls_warp) {
    for (int e = lane; e < MT * 16 * 8; e += 32) {
      const int r = e >> 3, c = e & 7;
      if (r < f) cp_async16(tile_ptr(qs, r, c), base + (size_t)(1 + r * n + p) * ld + c * 8);
      else *reinterpret_cast<uint4*>(tile_ptr(qs, r, c)) = make_uint4(0, 0, 0, 0);
    }
    for (int e = lane; e < NKT * 16 * 8; e += 32) {
      const int r = e >> 3, c = e & 7;
      if (r <= f) {
        const bf16* kr = base + (size_t)(r == 0 ? 0 : 1 + (r - 1) * n + p) * ld + c * 8;
        cp_async16(tile_ptr(ks, r, c), kr + inner);
        cp_async16(tile_ptr(vs, r, c), kr + 2 * inner);
      } else {
        *reinterpret_cast<uint4*>(tile_ptr(ks, r, c)) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(tile_ptr(vs, r, c)) = make_uint4(0, 0, 0, 0);
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();                               // allow_bits + (per-warp) tiles visible
  if (!active) return;
  if (cls_warp) {                                // keys of this group: frame k at patch p -> token 1 + k*n + p
    cls_partial_warp<NKT>(ks, vs, base, es_all[warp], f, NKT * 16 - 1, lane, [&](int k) { return frame_ok[k] != 0; },
                          [&](int k) { return 1 + k * n + p; }, cls_parts + ((size_t)bh * n + p) * kClsStride,
                          cls_scores ? cls_scores + (size_t)bh * N : nullptr);
    return;
  }
  auto allow = [&](int mt, int ql, int key) {
    const int q = mt * 16 + ql;
    return q < f ? ((allow_bits[q] >> key) & 1ull) != 0 : key == 0;
  };
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    // f a multiple of 16: the frame keys fill (NKT - 1) whole tiles and the CLS key is handled apart
    if (f == (NKT - 1) * 16 && NKT >= 2)
      attend_mtile_cls_apart<(NKT >= 2 ? NKT - 1 : 1)>(qs, ks, vs, mt * 16, lane, [&](int ql, int key) { return allow(mt, ql, key); });
    else
      attend_mtile<NKT>(qs, ks, vs, mt * 16, lane, [&](int ql, int key) { return allow(mt, ql, key); });
  }
  for (int e = lane; e < MT * 16 * 8; e += 32) {
    const int r = e >> 3, c = e & 7;
    if (r < f)
      *reinterpret_cast<uint4*>(out + ((size_t)b * N + 1 + r * n + p) * inner + h * 64 + c * 8) =
          *reinterpret_cast<const uint4*>(tile_ptr(qs, r, c));
  }
}

}  // namespace attn
}  // namespace mt
