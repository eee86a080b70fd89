// Host side of the fused divided-attention kernel (attention_fused.cuh): tile geometry, tensor maps, launch, and the
// merge of the CLS-query partials.  Reference: size_invariant_timesformer.py:109-144 (Attention.forward).
#include <cuda.h>

#include <stdlib.h>

#include <algorithm>

#include "attention_fused.cuh"
#include "common.cuh"

namespace mt {
namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct FusedWs { size_t parts, qkv_cls, total; };
FusedWs fused_ws_layout(int B, int f, int n, int heads) {
  FusedWs l;
  size_t off = 0;
  l.parts = off;   off += align_up((size_t)B * heads * (size_t)std::max(f, n) * attn::kClsStride * sizeof(float), 256);
  l.qkv_cls = off; off += align_up((size_t)B * 3 * heads * 64 * sizeof(bf16), 256);
  l.total = off;
  return l;
}

bool fused_supported(int f, int n, int heads, int dim_head, int dim) {
  return dim == fattn::kDim && dim_head == 64 && heads >= 1 && heads <= 16 && (f == 8 || f == 16 || f == 32) && n >= 1 &&
         n <= 55;
}

template <int MODE, int KT>
int launch_fused(const CUtensorMap& ta, const CUtensorMap& tcls, const CUtensorMap& tw, const fattn::Params& p,
                 cudaStream_t st) {
  auto kern = fattn::fused_attn_kernel<MODE, KT>;
  const size_t smem = fattn::smem_bytes();
  if (first_use_on_device(reinterpret_cast<const void*>(kern))) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "cudaFuncSetAttribute(fused_attn)");
  }
  const int pairs = std::max(1, std::min(p.total_steps, current_sms() / 2));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(fattn::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tcls, tw, p);
  if (e != cudaSuccess) return cuda_status(e, "cudaLaunchKernelEx(fused_attn)");
  MT_LAUNCH_CHECK("fused_attn_kernel");
  return MT_OK;
}

}  // namespace
}  // namespace mt

using namespace mt;

extern "C" int mt_fused_attn_supported(int f, int n, int heads, int dim_head, int dim) {
  return fused_supported(f, n, heads, dim_head, dim) ? 1 : 0;
}

extern "C" size_t mt_fused_attn_workspace_bytes(int batch, int f, int n, int heads) {
  if (batch <= 0 || f <= 0 || n <= 0 || heads <= 0) return 0;
  return fused_ws_layout(batch, f, n, heads).total;
}

extern "C" int mt_fused_attn_fwd(const void* xn, const void* w_qkv_heads, const uint8_t* mask, const uint8_t* identities_mask,
                                 int mode, void* out, float* cls_attn, int batch, int f, int n, int heads, int dim_head,
                                 int dim, void* workspace, size_t workspace_bytes, void* stream) {
  MT_REQUIRE(xn && w_qkv_heads && mask && out && workspace, "fused_attn: null pointer");
  MT_REQUIRE(mode == MT_ATTN_TIME || mode == MT_ATTN_SPACE, "fused_attn: unknown mode %d", mode);
  MT_REQUIRE(mode == MT_ATTN_SPACE || identities_mask, "fused_attn: time mode needs identities_mask");
  MT_REQUIRE(batch > 0, "fused_attn: empty batch");
  if (!fused_supported(f, n, heads, dim_head, dim)) {
    set_error("fused_attn: no fused schedule for f=%d n=%d heads=%d dim_head=%d dim=%d", f, n, heads, dim_head, dim);
    return MT_ERR_UNSUPPORTED;
  }
  const FusedWs l = fused_ws_layout(batch, f, n, heads);
  if (workspace_bytes < l.total) {
    set_error("fused_attn: workspace too small (%zu < %zu)", workspace_bytes, l.total);
    return MT_ERR_WORKSPACE;
  }
  MT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(xn) & 15) == 0,
             "fused_attn: workspace must be 256-byte aligned, xn 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int N = 1 + f * n, inner = heads * 64;
  const bf16* x = reinterpret_cast<const bf16*>(xn);

  fattn::Params p{};
  p.B = batch; p.f = f; p.n = n; p.heads = heads; p.N = N;
  p.mask = mask; p.idmask = identities_mask;
  p.out = reinterpret_cast<bf16*>(out);
  p.cls_parts = reinterpret_cast<float*>(ws + l.parts);
  p.cls_scores = cls_attn;
  p.qkv_cls = reinterpret_cast<bf16*>(ws + l.qkv_cls);

  CUtensorMap ta, tcls, tw;
  int rc;
  const unsigned long long row_b = (unsigned long long)dim * 2;
  if (mode == MT_ATTN_TIME) {
    p.pt = 127 / f;
    p.tiles_per_video = (n + p.pt - 1) / p.pt;
    p.cls_row = f * p.pt;
    p.a_bytes_kb = (f * p.pt + 1) * 128;
    // token 1 + frame*n + patch of video b: dims {channel, patch, frame, video}; a box is one patch position x f frames
    const unsigned long long dims[4] = {(unsigned long long)dim, (unsigned long long)n, (unsigned long long)f,
                                        (unsigned long long)batch};
    const unsigned long long strides[3] = {row_b, (unsigned long long)n * row_b, (unsigned long long)N * row_b};
    const unsigned box[4] = {64, 1, (unsigned)f, 1};
    rc = make_tmap_4d_bf16_sw128(&ta, x + dim, dims, strides, box);
  } else {
    p.pt = 0;
    p.tiles_per_video = (f + 1) / 2;
    p.cls_row = fattn::kSpaceClsRow;
    p.a_bytes_kb = 2 * (n + 1) * 128;
    const unsigned long long dims[4] = {(unsigned long long)dim, (unsigned long long)n, (unsigned long long)f,
                                        (unsigned long long)batch};
    const unsigned long long strides[3] = {row_b, (unsigned long long)n * row_b, (unsigned long long)N * row_b};
    const unsigned box[4] = {64, (unsigned)n, 1, 1};
    rc = make_tmap_4d_bf16_sw128(&ta, x + dim, dims, strides, box);
  }
  if (rc) return rc;
  rc = make_tmap_weights_kmajor(&tcls, xn, batch * N, dim, 1, 64);
  if (rc) return rc;
  rc = make_tmap_weights_kmajor(&tw, w_qkv_heads, heads * 192, dim, fattn::kBRows, 64);
  if (rc) return rc;
  p.n_tiles = batch * p.tiles_per_video;
  p.n_pair_tiles = (p.n_tiles + 1) / 2;
  p.total_steps = p.n_pair_tiles * heads;

  {
    // algorithmic work: the projection (2 * tokens * dim * 3*inner) + the attention core; bytes: xn in, out out
    const double gq = mode == MT_ATTN_TIME ? f : n, groups = (double)batch * heads * (mode == MT_ATTN_TIME ? n : f);
    ProfScope prof(st, 2.0 * batch * N * (double)dim * 3 * inner + 4.0 * groups * 64.0 * gq * (gq + 1),
                   (double)batch * N * (dim + inner) * 2.0 + 3.0 * inner * dim * 2.0,
                   mode == MT_ATTN_TIME ? "fused_attn_time" : "fused_attn_space");
    if (mode == MT_ATTN_TIME) rc = f == 32 ? launch_fused<MT_ATTN_TIME, 2>(ta, tcls, tw, p, st) : launch_fused<MT_ATTN_TIME, 1>(ta, tcls, tw, p, st);
    else rc = launch_fused<MT_ATTN_SPACE, 1>(ta, tcls, tw, p, st);
    if (rc) return rc;
  }
  attn::cls_combine_kernel<<<batch * heads, 64, 0, st>>>(p.qkv_cls, 1, p.cls_parts, p.out, cls_attn, N, mode == MT_ATTN_SPACE ? f : n, heads);
  MT_LAUNCH_CHECK("cls_combine_kernel");
  return MT_OK;
}
