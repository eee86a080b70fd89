/*
 * mintime_b200 -- C ABI of libmintime_b200.so (hand-written sm_100a kernels for MINTIME's hot path).
 *
 * The reference has no FFI of its own for this path (pure PyTorch): the boundary
 * a maintainer binds is this header, called from the Python nn.Module shims that keep the reference's
 * class names / forward signatures (see INTEGRATION.md for the ctypes stubs).  Every entry point
 *   - takes plain device pointers + sizes + a CUDA stream (cudaStream_t passed as void*),
 *   - allocates nothing and never synchronises: the caller owns all memory incl. the workspace,
 *   - is stateless and re-entrant per stream,
 *   - returns 0 on success, <0 for an invalid argument/shape (mt_last_error() has the text),
 *     or a positive cudaError_t value if a launch failed.
 * "T" below is the activation/weight element type selected by `precision`:
 *   MT_PREC_FP32 -> float (exact path, FFMA kernels), MT_PREC_BF16 -> __nv_bfloat16 (tcgen05 path).
 *
 * Each function names the reference code it replaces (paths relative to the reference repo).
 */
#ifndef MINTIME_B200_H
#define MINTIME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MT_ABI_VERSION 5   /* 2: mt_divided_attn_fwd takes a workspace; + mt_expand_dwconv_*, mt_clip_meta_fwd
                            * 3: + the backward entry points of the transformer (mt_*_bwd, mt_grad_prep, mt_geglu_*)
                            * 4: mt_clip_meta_fwd mask_padding semantics; fused divided attention; extractor training
                            * 5: + mt_linear_wgrad_nt (weight gradient from row-major operands), mt_geglu_bwd_colsum, mt_layernorm_copy_fwd, mt_xception_* */

enum { MT_PREC_FP32 = 0, MT_PREC_BF16 = 1 };
enum { MT_OK = 0, MT_ERR_ARG = -1, MT_ERR_WORKSPACE = -2, MT_ERR_UNSUPPORTED = -3, MT_ERR_DRIVER = -4 };
enum { MT_ATTN_TIME = 0, MT_ATTN_SPACE = 1 };
enum { MT_IN_F32 = 0, MT_IN_U8 = 1 };

int mt_abi_version(void);
/* Last error text of the calling thread ("" if none). */
const char* mt_last_error(void);
/* 0 if the current CUDA device can run the library (compute capability 10.x), else MT_ERR_UNSUPPORTED. */
int mt_device_check(void);

/* ---------------------------------------------------------------------------------------------
 * Packed weights (device pointers).  Packing (BatchNorm folding, layout, dtype) is done once at load
 * time by the host shim (weights.py); layouts:
 *   pointwise conv / linear  : w  T [Cout][Cin]   (BN scale folded into rows), shift f32 [Cout]
 *   depthwise conv           : w  f32 [k*k][C]    (BN scale folded),            shift f32 [C]
 *   stem conv                : w  f32 [27][32]    ((ky*3+kx)*3+ci major, BN scale folded), shift f32 [32]
 *   squeeze-excite           : f32, reduce [Sq][C] + bias [Sq], expand TRANSPOSED [Sq][C] + bias [C]
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* w;      /* T   [cout][cin] */
  const float* shift; /* f32 [cout] (BN beta - mean*scale, or the linear bias); may be NULL */
} mt_pw_t;

typedef struct {
  mt_pw_t expand;          /* w == NULL when expand_ratio == 1 (block 0)   model.py:57-64   */
  const float* dw_w;       /* [k*k][cexp]                                   model.py:66-73   */
  const float* dw_shift;   /* [cexp] */
  const float* se_reduce_w; /* [sq][cexp]                                   model.py:76-80   */
  const float* se_reduce_b; /* [sq] */
  const float* se_expand_w; /* [sq][cexp] (transposed _se_expand.weight) */
  const float* se_expand_b; /* [cexp] */
  mt_pw_t project;         /* [cout][cexp]                                  model.py:83-86   */
} mt_mbconv_t;

typedef struct {
  const float* stem_w;     /* model.py:160-165 */
  const float* stem_shift;
  mt_mbconv_t blocks[16];
  mt_pw_t head;            /* model.py:193-197 */
} mt_effnet_b0_weights_t;

typedef struct {
  const float* ln_g; const float* ln_b;    /* PreNorm LayerNorm            size_invariant_timesformer.py:18-26 */
  const void* w_qkv;                       /* T [3*inner][dim], q rows pre-scaled by dim_head^-0.5 (:114) */
  const void* w_out; const float* b_out;   /* T [dim][inner], f32 [dim]                                   (:103-106) */
  const void* w_qkv_heads;                 /* bf16 [heads][3*dim_head][dim]: the rows of w_qkv regrouped per head (q | k | v
                                            * of head h contiguous) for the fused attention kernel; NULL = unfused path */
} mt_attn_weights_t;

typedef struct {
  const float* ln_g; const float* ln_b;
  const void* w1; const float* b1;         /* T [8*dim][dim] rows interleaved in blocks of 32 (x | gate), f32 [8*dim] same order (:68-73) */
  const void* w2; const float* b2;         /* T [dim][4*dim], f32 [dim] */
} mt_ff_weights_t;

typedef struct {
  int dim, depth, heads, dim_head, num_frames, num_patches, channels, num_classes;
  int enable_pos_emb, enable_size_emb;
} mt_tsf_cfg_t;

#define MT_TSF_MAX_DEPTH 16
typedef struct {
  const void* w_patch; const float* b_patch;   /* T [dim][channels], f32 [dim]   (:175) */
  const float* cls_token;                      /* f32 [dim]                      (:176) */
  const float* pos_emb;                        /* f32 [num_frames*channels+1][dim] (:178) */
  const float* size_emb;                       /* f32 same shape or NULL          (:179-180) */
  mt_attn_weights_t time_attn[MT_TSF_MAX_DEPTH];
  mt_attn_weights_t space_attn[MT_TSF_MAX_DEPTH];
  mt_ff_weights_t ff[MT_TSF_MAX_DEPTH];
  const float* out_ln_g; const float* out_ln_b; /* to_out.0 (:195-198) */
  const float* out_w; const float* out_b;       /* f32 [num_classes][dim], [num_classes] */
} mt_tsf_weights_t;

/* Xception (models/xception.py:93-137), the alternative 2048-channel extractor (train.py:129-133):
 *   conv1 / conv2      : dense 3x3 as im2col GEMMs: w T [cout][kp], column (ky*3+kx)*cin + ci, kp = 32 (27 padded) / 288,
 *                        bn1 / bn2 folded
 *   sep[34]            : the separable convolutions in forward order -- block1..block12 `rep` entries, conv3, conv4:
 *                        dw_w f32 [9][cin] (SeparableConv2d.conv1, tap-major), pw = pointwise 1x1 with the following
 *                        BatchNorm folded
 *   skip[4]            : block1/2/3/12 `skip` 1x1 stride-2 projection with `skipbn` folded */
typedef struct {
  const float* dw_w;
  mt_pw_t pw;
} mt_xc_sep_t;

typedef struct {
  mt_pw_t conv1, conv2;
  mt_xc_sep_t sep[34];
  mt_pw_t skip[4];
} mt_xception_weights_t;

/* ---------------------------------------------------------------------------------------------
 * Whole-model entry points
 * ------------------------------------------------------------------------------------------- */

/* Xception.forward == Xception.features (models/xception.py:146-184, :196-198), eval mode.
 *   x      : frames NHWC [n_img][in_hw][in_hw][3], f32 or u8, as given to the module (no normalisation inside)
 *   feats  : T [n_img * o * o][2048], o = mt_xception_out_hw(in_hw) (7 for 224): the reference's (n_img,2048,o,o) permuted
 * Workspace: mt_xception_workspace_bytes(n_img, in_hw, precision), 1024-byte aligned. */
int mt_xception_out_hw(int in_hw);
size_t mt_xception_workspace_bytes(int n_img, int in_hw, int precision);
int mt_xception_fwd(const mt_xception_weights_t* w, const void* x, int x_dtype, void* feats, int n_img, int in_hw,
                    int precision, void* workspace, size_t workspace_bytes, void* stream);

/* EfficientNet.forward (models/efficientnet/efficientnet_pytorch/model.py:267-288), eval mode.
 *   x      : frames, NHWC [n_img][224][224][3], f32 (MT_IN_F32) or u8 (MT_IN_U8), raw 0..255
 *   feats  : T [n_img*49][1280]  == NHWC feature map == the 'b (f h w) c' token layout of
 *            size_invariant_timesformer.py:227, i.e. the reference's (n_img,1280,7,7) output permuted.
 * Workspace: mt_effnet_b0_workspace_bytes(n_img, precision). */
size_t mt_effnet_b0_workspace_bytes(int n_img, int precision);
int mt_effnet_b0_fwd(const mt_effnet_b0_weights_t* w, const void* x, int x_dtype, void* feats, int n_img,
                     int precision, void* workspace, size_t workspace_bytes, void* stream);

/* SizeInvariantTimeSformer.forward (models/size_invariant_timesformer.py:224-276).
 *   feats          : T [B][f*49][channels]
 *   mask           : u8 [B][f]           (bool)       identities_mask : u8 [B][f][f]
 *   size_embedding : i32 [B][f]                        positions       : i64 [B][1+f*49]
 *   logits         : f32 [B][num_classes]
 *   space_attn/time_attn : f32 [B*heads][1+f*49] each (last layer's CLS attention, (b h) row order,
 *                    :271) or NULL when attention maps are not required. */
size_t mt_tsf_workspace_bytes(const mt_tsf_cfg_t* cfg, int batch, int precision);
int mt_tsf_fwd(const mt_tsf_weights_t* w, const mt_tsf_cfg_t* cfg, const void* feats, const uint8_t* mask,
               const uint8_t* identities_mask, const int32_t* size_embedding, const int64_t* positions,
               float* logits, float* space_attn, float* time_attn, int batch, int precision, void* workspace,
               size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Building blocks (the whole-model calls are sequences of these; exported for per-kernel parity
 * tests and for callers that fuse differently)
 * ------------------------------------------------------------------------------------------- */

/* 1x1 conv / linear:  out[m][n] = act( sum_k a[m][k]*gate[m/rows_per_gate][k] * w[n][k] + shift[n] ) + residual[m][n]
 * Replaces F.conv2d 1x1 + BatchNorm2d(eval) + swish (+ SE scale on the input, + skip add):
 * model.py:100-103 (expand), :113-119 (SE multiply + project + bn2), :122-127 (skip), :286 (head);
 * and nn.Linear (to_qkv, size_invariant_timesformer.py:111).
 * gate (f32 [m/rows_per_gate][k]) and residual (T [m][n]) may be NULL; act: 0 none, 1 swish. */
int mt_pointwise_fwd(int precision, const void* a, const void* w, const float* shift, const float* gate,
                     int rows_per_gate, const void* residual, int act, void* out, int m, int n, int k, void* stream);

/* x[m][n] += sum_k a[m][k]*w[n][k] + bias[n]   (x: f32 residual stream, in place)
 * Replaces to_out Linear + residual add (size_invariant_timesformer.py:144,265,267) and FF net.3 + residual (:73,268). */
int mt_linear_residual_fwd(int precision, const void* a, const void* w, const float* bias, float* x, int m, int n,
                           int k, void* stream);

/* GEGLU feed-forward first half: out[m][j] = u[j] * gelu_erf(g[j]),  (u|g) = a*W1^T + b1   (:60-63, :68-70)
 * w/bias rows are interleaved in blocks of 32 (32 u-rows, then their 32 gate rows); n = 2*n_out. */
int mt_linear_geglu_fwd(int precision, const void* a, const void* w, const float* bias, void* out, int m, int n,
                        int k, void* stream);

/* Token build (:225-248): x[b][0] = cls + pos[positions[b][0]] + size[0];
 * x[b][1+t] = feats[b][t]*Wp^T + bp + pos[positions[b][1+t]] + size[size_embedding[b][t/49]]   (x: f32 [B][1+f*49][dim]) */
int mt_patch_embed_fwd(int precision, const mt_tsf_weights_t* w, const mt_tsf_cfg_t* cfg, const void* feats,
                       const int32_t* size_embedding, const int64_t* positions, float* x, int batch, void* stream);

/* nn.LayerNorm(dim) over the last axis, eps 1e-5 (:18-26): f32 [rows][dim] -> T [rows][dim] */
int mt_layernorm_fwd(int precision, const float* x, const float* gamma, const float* beta, void* out, int rows,
                     int dim, void* stream);
/* The same, also writing x_copy f32 [rows][dim] = x (the training forward keeps every sub-block's input for the LayerNorm
 * backward; the residual GEMM that follows updates x in place).  x_copy may be NULL. */
int mt_layernorm_copy_fwd(int precision, const float* x, const float* gamma, const float* beta, void* out, float* x_copy,
                          int rows, int dim, void* stream);

/* Attention core of Attention.forward (:114-141) on an already projected qkv (q pre-scaled):
 *   qkv T [B][1+f*n][3*heads*dim_head] -> out T [B][1+f*n][heads*dim_head] (heads merged, before to_out)
 *   mode MT_ATTN_TIME : groups (b,h,patch): frames attend frames (+CLS key), bias = mask[b][k] & identities_mask[b][q][k]
 *   mode MT_ATTN_SPACE: groups (b,h,frame): patches attend patches (+CLS key), no mask
 *   CLS row: attends all tokens with the padded-frame mask; cls_attn f32 [B*heads][1+f*n] (may be NULL).
 *   workspace: mt_divided_attn_workspace_bytes(batch, f, n, heads) bytes, 16-byte aligned; the bf16 path keeps the
 *   per-group partials of the CLS row there (computed inside the grouped kernels, merged by a small combine
 *   kernel).  NULL selects the stand-alone CLS-row kernel (always used by the fp32 path). */
size_t mt_divided_attn_workspace_bytes(int batch, int f, int n, int heads);
int mt_divided_attn_fwd(int precision, const void* qkv, const uint8_t* mask, const uint8_t* identities_mask,
                        int mode, void* out, float* cls_attn, int batch, int f, int n, int heads, int dim_head,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Front half of MBConvBlock.forward in one kernel, bf16 only (model.py:98-107): 1x1 expand conv + BN + swish ->
 * depthwise kxk stride s (TF-SAME padding) + BN + swish, plus the squeeze-excite pool partial sums; the expanded
 * tensor never leaves the SM (tcgen05 contraction into TMEM -> shared memory -> stencil).
 *   in bf16 NHWC [n_img][h][h][cin]; w_exp bf16 [cexp][cin] (BN scale folded), exp_shift f32 [cexp];
 *   w_dw f32 [k*k][cexp] tap-major (BN scale folded), dw_shift f32 [cexp];
 *   out bf16 NHWC [n_img][ho][ho][cexp]; pool_part f32 [n_img][mt_expand_dwconv_chunks(...)][cexp].
 * mt_expand_dwconv_chunks returns 0 when the shape has no fused schedule (cin > 64, ...): callers then run
 * mt_pointwise_fwd + mt_dwconv_fwd.  mt_mbconv_fwd / mt_effnet_b0_fwd make that choice themselves. */
int mt_expand_dwconv_chunks(int h, int cin, int cexp, int k, int s);
int mt_expand_dwconv_fwd(const void* in, const void* w_exp, const float* exp_shift, const float* w_dw,
                         const float* dw_shift, void* out, float* pool_part, int n_img, int h, int cin, int cexp, int k,
                         int s, void* stream);

/* Clip metadata assembled on the device instead of in DeepFakesDataset.__getitem__ (deepfakes_dataset.py:259-330) /
 * predict.py generate_masks (:254-352), from the per-identity slot table of each clip:
 *   slots, n_real i32 [batch][max_identities]: face slots owned by each identity (sum <= f) and how many hold a face;
 *   frame_no, ratio i32 [batch][f]: source frame number and int(face_area*100/video_area) of the face in each slot
 *   mask_padding: 1 -> padded slots get mask 0 (predict.py:300-306); 0 -> every slot is valid, which is what
 *   DeepFakesDataset.__getitem__ returns AS EXECUTED (its test at :283 runs after :276 has padded the list, so the
 *   all-ones branch :286 is always taken; tests/golden/clip_meta_ref.json holds outputs of the reference class).
 *   Padded slots repeat the largest frame number of the clip so far (:277, predict.py:304).
 * -> mask u8 [batch][f], identities_mask u8 [batch][f][f], size_embedding i32 [batch][f] (0 = padding, 1..20 = size
 *    bucket), positions i64 [batch][1 + f*n_patches] -- the tensors mt_tsf_fwd / the forward of the model take. */
int mt_clip_meta_fwd(const int32_t* slots, const int32_t* n_real, const int32_t* frame_no, const int32_t* ratio,
                     int max_identities, int mask_padding, uint8_t* mask, uint8_t* identities_mask,
                     int32_t* size_embedding, int64_t* positions, int batch, int f, int n_patches, void* stream);

/* Attention.forward up to the head merge (:109-141) in ONE kernel, bf16 only: per-head QKV projection of the
 * LayerNorm'd tokens on tcgen05 (CTA pairs, the 128-token tile resident in shared memory, W streamed per head) ->
 * identity-masked softmax(QK^T)V on the tiles in shared memory -> out.  qkv never goes to HBM.
 *   xn bf16 [B][1+f*n][dim] (the PreNorm output), w_qkv_heads as in mt_attn_weights_t, mask / identities_mask / mode /
 *   out / cls_attn as in mt_divided_attn_fwd.  workspace: mt_fused_attn_workspace_bytes(...), 256-byte aligned.
 * mt_fused_attn_supported() tells whether a shape has a fused schedule (dim 512, dim_head 64, f in {8,16,32}, n <= 55);
 * mt_fused_attn_fwd returns MT_ERR_UNSUPPORTED otherwise and mt_tsf_fwd falls back to LayerNorm -> mt_pointwise_fwd ->
 * mt_divided_attn_fwd. */
int mt_fused_attn_supported(int f, int n, int heads, int dim_head, int dim);
size_t mt_fused_attn_workspace_bytes(int batch, int f, int n, int heads);
int mt_fused_attn_fwd(const void* xn, const void* w_qkv_heads, const uint8_t* mask, const uint8_t* identities_mask,
                      int mode, void* out, float* cls_attn, int batch, int f, int n, int heads, int dim_head, int dim,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * EfficientNet-B0 in TRAIN mode (train.py:153-170: the extractor trains unless --freeze_backbone): fp32 NHWC building
 * blocks of MBConvBlock.forward / backward (model.py:89-128) that the eval path folds away.  1x1 convolutions use
 * mt_pointwise_fwd (forward; data gradient on the transposed weight) and mt_grad_prep + mt_linear_wgrad.
 * Orchestrated by mintime_b200/efficientnet_train.py.  `workspace`: mt_extractor_train_workspace_bytes(rows, c).
 * ------------------------------------------------------------------------------------------- */
size_t mt_extractor_train_workspace_bytes(long long rows, int c);
/* nn.BatchNorm2d in train mode (utils.py:520-521): per-channel mean / BIASED variance of x [rows][c]; when running_* are
 * given they move by `momentum` towards mean / UNBIASED variance (in place). */
int mt_bn_stats(const float* x, float* mean, float* var, float* running_mean, float* running_var, float momentum,
                long long rows, int c, void* workspace, size_t workspace_bytes, void* stream);
/* out = act(gamma * (x - mean) / sqrt(var + eps) + beta), act 0 = none, 1 = swish (utils.py:64-69) */
int mt_bn_act_fwd(const float* x, const float* mean, const float* var, const float* gamma, const float* beta, int act,
                  float eps, float* out, long long rows, int c, void* stream);
/* backward of mt_bn_act_fwd through the batch statistics (swish backward utils.py:71-80): dx, dgamma [c], dbeta [c] */
int mt_bn_act_bwd(const float* dy, const float* x, const float* mean, const float* var, const float* gamma,
                  const float* beta, int act, float eps, float* dx, float* dgamma, float* dbeta, long long rows, int c,
                  void* workspace, size_t workspace_bytes, void* stream);
/* raw stem conv 3x3 s2 3->32, TF-SAME padding (utils.py:248-276): x [n][h][w][3], w [27][32] tap-major -> out [n][h/2][w/2][32];
 * mt_stem_wgrad: dw [27][32] = the weight gradient for output gradient dy */
int mt_stem_raw_fwd(const float* x, const float* w, float* out, int n_img, int h, int w_, void* stream);
int mt_stem_wgrad(const float* x, const float* dy, float* dw, int n_img, int h, int w_, void* workspace, size_t workspace_bytes,
                  void* stream);
/* raw depthwise kxk stride s, TF-SAME padding: in [n][h][h][c], w [k*k][c] tap-major -> out [n][ho][ho][c]; data / weight gradients */
int mt_dwconv_raw_fwd(const float* in, const float* w, float* out, int n_img, int h, int c, int k, int s, void* stream);
int mt_dwconv_dgrad(const float* dy, const float* w, float* dx, int n_img, int h, int c, int k, int s, void* stream);
int mt_dwconv_wgrad(const float* in, const float* dy, float* dw, int n_img, int h, int c, int k, int s, void* workspace,
                    size_t workspace_bytes, void* stream);
/* weight gradient of a 1x1 convolution: dw [co][ci] = dy [rows][co]^T a [rows][ci] (any channel counts; fixed-order sums) */
size_t mt_conv1x1_wgrad_workspace_bytes(long long rows, int co, int ci);
int mt_conv1x1_wgrad(const float* dy, const float* a, float* dw, long long rows, int co, int ci, void* workspace,
                     size_t workspace_bytes, void* stream);
/* squeeze-excite (model.py:110-115): pooled mean of x [groups][rows][c]; the two FC layers (wr [sq][c], we [c][sq], the
 * reference's layouts) -> gate [n][c] (+ s_pre [n][sq] kept for the backward); their backward; the gate multiply's. */
int mt_group_mean(const float* x, float* out, int groups, int rows, int c, void* stream);
int mt_se_fc_fwd(const float* mean, const float* wr, const float* br, const float* we, const float* be, float* gate,
                 float* s_pre, int n_img, int c, int sq, void* stream);
int mt_se_fc_bwd(const float* dgate, const float* gate, const float* s_pre, const float* mean, const float* wr,
                 const float* we, float* dmean, float* dwr, float* dbr, float* dwe, float* dbe, int n_img, int c, int sq,
                 void* workspace, size_t workspace_bytes, void* stream);
int mt_gate_mul(const float* x, const float* gate, float* out, int n_img, int rows, int c, void* stream);
/* phase 0: dgate [n][c] = sum_rows dxg * x;  phase 1: dx = dxg * gate + dmean / rows (gate multiply + avg-pool backward) */
int mt_gate_bwd(const float* dxg, const float* x, const float* gate, const float* dmean, float* dgate, float* dx, int n_img,
                int rows, int c, int phase, void* stream);
/* out = x * scale[image] (+ skip): drop-connect (utils.py:129-154) + residual add (model.py:123-127); scale / skip may be NULL */
int mt_scale_add(const float* x, const float* scale, const float* skip, float* out, int n_img, long long per_img, void* stream);

/* Stem: ZeroPad2d(0,1,0,1) + conv 3x3 s2 (3->32) + BN + swish (utils.py:248-276, model.py:276)
 *   x NHWC [n_img][H][W][3] (f32/u8) -> out T NHWC [n_img][H/2][W/2][32] */
int mt_stem_fwd(int precision, const void* x, int x_dtype, const float* w, const float* shift, void* out, int n_img,
                int h, int w_, void* stream);

/* Depthwise kxk stride s conv with TF-SAME padding + BN + swish, plus partial sums for the SE squeeze
 * (model.py:105-107,110): in T NHWC [n_img][h][w][c] -> out T NHWC [n_img][ceil(h/s)][ceil(w/s)][c];
 * pool_part f32 [n_img][n_chunks][c], n_chunks = mt_dwconv_chunks(precision, h, w, c, k, s): sum of out over
 * each chunk (spatial tile) of output pixels (every entry is written; no zero-init; no atomics: deterministic).
 * bf16 path: tensor-core kernel (TMA-staged NHWC tile, block-diagonal mma.sync); fp32 path: FFMA kernel. */
int mt_dwconv_chunks(int precision, int h, int w_, int c, int k, int s);
int mt_dwconv_fwd(int precision, const void* in, const float* w, const float* shift, void* out, float* pool_part,
                  int n_img, int h, int w_, int c, int k, int s, void* stream);

/* SE excitation (model.py:110-115): mean = sum_chunks(pool_part)/hw;  gate[i][c] = sigmoid(We*swish(Wr*mean + br) + be) */
int mt_se_gate_fwd(const float* pool_part, int n_chunks, int hw, const float* wr, const float* br, const float* we_t,
                   const float* be, float* gate, int n_img, int c, int sq, void* stream);

/* One MBConvBlock.forward in eval mode (model.py:89-128): in T NHWC [n_img][hw_in][hw_in][cin] ->
 * out T NHWC [n_img][ceil(hw_in/stride)]^2[cout].  mt_effnet_b0_block_spec(i, &spec) fills the B0 table entry i. */
typedef struct { int kernel, stride, expand, cin, cout, hw_in; } mt_mbconv_spec_t;
int mt_effnet_b0_block_spec(int index, mt_mbconv_spec_t* spec);
size_t mt_mbconv_workspace_bytes(const mt_mbconv_spec_t* spec, int n_img, int precision);
int mt_mbconv_fwd(int precision, const mt_mbconv_spec_t* spec, const mt_mbconv_t* w, const void* in, void* out,
                  int n_img, void* workspace, size_t workspace_bytes, void* stream);

/* Classification head (:195-198,270-276): logits[b] = LayerNorm(x[b][0]) * W^T + bias */
int mt_head_fwd(const float* x, const float* ln_g, const float* ln_b, const float* w, const float* bias, float* logits,
                int batch, int tokens, int dim, int num_classes, void* stream);

/* Attention aggregation of predict.py/test.py (utils.py:68-96), per video: per-token max over heads of the CLS
 * attention maps, np.array_split into num_frames chunks, mean * scale, softmax over frames; for the space map,
 * the time map and their sum.  space_attn/time_attn f32 [batch*heads][tokens] -> out f32 [batch][3][num_frames]. */
int mt_aggregate_attn_fwd(const float* space_attn, const float* time_attn, float* out, int batch, int heads,
                          int num_frames, int tokens, float scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward pass of the transformer: `loss.backward()` of train.py:376-378 for SizeInvariantTimeSformer with a frozen
 * extractor (train.py:344-346).  The two contractions of every Linear reuse the forward GEMMs:
 *   dgrad  dX[m][k]  = dY[m][n] * W[n][k]      -> mt_pointwise_fwd(a = dY, w = W^T stored [k][n])
 *   wgrad  dW[n][k] += dY^T * X                -> mt_linear_residual_fwd(a = dY^T [n][mp], w = X^T [k][mp], x = dW f32)
 * with the transposed operands produced by mt_grad_prep.  mintime_b200/training.py holds the schedule.
 * ------------------------------------------------------------------------------------------- */

/* One pass over src [m][c] (f32 when src_is_f32, else T) producing any of: out_rm T [m][c] (cast), out_t T [c][mp]
 * (transpose; columns m..mp-1 zero; mp a multiple of 64), colsum f32 [c] (column sums = the bias gradient when src is
 * dY).  rows_per_batch > 0 skips the CLS row of every video: source row = i + i / rows_per_batch + 1.
 * workspace: mt_grad_prep_workspace_bytes(mp, c) bytes when colsum is requested. */
size_t mt_grad_prep_workspace_bytes(int m, int c);
int mt_grad_prep(int precision, const void* src, int src_is_f32, void* out_rm, void* out_t, float* colsum, int m, int c,
                 int mp, int rows_per_batch, void* workspace, size_t workspace_bytes, void* stream);

/* Weight gradient of a Linear: dw f32 [n_out][k_in] += dy_t [n_out][mp] * x_t [k_in][mp]^T  (both operands T, from
 * mt_grad_prep; contraction over the mp tokens).  Same GEMM as mt_linear_residual_fwd with the k-blocks of every
 * output tile split over the SMs (partial tiles are summed by the TMA reduce-add of the epilogue, so the summation order
 * of the fp32 partials is not fixed). */
int mt_linear_wgrad(int precision, const void* dy_t, const void* x_t, float* dw, int n_out, int k_in, int mp, void* stream);

/* The same weight gradient straight from the row-major operands: dw f32 [n_out][k_in] += dy [m][n_out]^T * x [m][k_in]
 * (bf16 only; n_out, k_in multiples of 8; any m).  Both operands enter tcgen05.mma MN-major (the contraction index is
 * the slow one in memory), so no transposed copies (mt_grad_prep out_t) are needed: the backward of
 * size_invariant_timesformer.py:109-144 / :65-76 w.r.t. to_qkv / to_out / net.0 / net.3 weights. */
int mt_linear_wgrad_nt(int precision, const void* dy, const void* x, float* dw, int n_out, int k_in, int m, void* stream);

/* out[c] (+)= sum_r in[r][c], fixed summation order. */
int mt_colsum_f32(const float* in, float* out, int rows, int cols, int accumulate, void* stream);

/* nn.LayerNorm backward of a PreNorm sub-block (size_invariant_timesformer.py:18-26): gx[rows][dim] (f32 gradient of the
 * residual stream) += dLN(dy);  dgamma_dbeta f32 [2*dim] = (dgamma | dbeta) (assigned). */
size_t mt_layernorm_bwd_workspace_bytes(int rows, int dim);
int mt_layernorm_bwd(int precision, const float* x, const float* gamma, const void* dy, float* gx, float* dgamma_dbeta,
                     int rows, int dim, void* workspace, size_t workspace_bytes, void* stream);

/* GEGLU on a stored pre-activation h T [m][2*n_out] in the interleaved column order of mt_ff_weights_t.w1
 * (size_invariant_timesformer.py:60-63): out T [m][n_out] = u * gelu_erf(g); backward: dh T [m][2*n_out] same order.
 * (The training forward keeps h; inference uses the fused mt_linear_geglu_fwd.) */
int mt_geglu_fwd(int precision, const void* h, void* out, int m, int n_out, void* stream);
int mt_geglu_bwd(int precision, const void* h, const void* dout, void* dh, int m, int n_out, void* stream);
/* mt_geglu_bwd that also returns colsum f32 [2*n_out] = the column sums of dh (= d loss / d net.0.bias, interleaved
 * order), summed in a fixed order; workspace: mt_geglu_bwd_colsum_workspace_bytes(m, n_out) bytes. */
size_t mt_geglu_bwd_colsum_workspace_bytes(int m, int n_out);
int mt_geglu_bwd_colsum(int precision, const void* h, const void* dout, void* dh, float* colsum, int m, int n_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Backward of mt_divided_attn_fwd: qkv, mask, identities_mask as in the forward (probabilities are recomputed),
 * dout T [B][1+f*n][heads*dim_head] -> dqkv T [B][1+f*n][3*heads*dim_head] (every element written). */
size_t mt_divided_attn_bwd_workspace_bytes(int batch, int f, int n, int heads);
int mt_divided_attn_bwd(int precision, const void* qkv, const void* dout, const uint8_t* mask,
                        const uint8_t* identities_mask, int mode, void* dqkv, int batch, int f, int n, int heads,
                        int dim_head, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the token build (:225-248) w.r.t. the embedding tables and the CLS token: g0 f32 [B][1+f*n][dim] is
 * scattered (+=, fp32 atomics) into dpos / dsize ([table_rows][dim], may be NULL) and dcls [dim]; the caller zeroes them.
 * An index outside [0, table_rows) traps the kernel (nn.Embedding raises IndexError), in the forward as well. */
int mt_embed_bwd(const float* g0, const int64_t* positions, const int32_t* size_embedding, float* dpos, float* dsize,
                 float* dcls, int batch, int f, int n, int dim, int table_rows, void* stream);

/* Backward of mt_head_fwd: gx[b][0][:] = dL/dx[b][0] (assigned; the other rows of gx are the caller's to zero);
 * grads f32 [classes*dim + classes + 2*dim] = (dW | dbias | dgamma | dbeta). */
size_t mt_head_bwd_workspace_bytes(int batch, int dim, int num_classes);
int mt_head_bwd(const float* x, const float* ln_g, const float* ln_b, const float* w, const float* dlogits, float* gx,
                float* grads, int batch, int tokens, int dim, int num_classes, void* workspace, size_t workspace_bytes,
                void* stream);

/* ---------------------------------------------------------------------------------------------
 * Diagnostics (used by bench.py; off by default, no effect on results)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  char name[64];        /* kernel class, e.g. "gemm_tc geglu N4096 K512" */
  double ms_total;      /* CUDA-event time summed over `count` launches */
  double flops_total;   /* algorithmic flops of those launches */
  double bytes_total;   /* algorithmic HBM bytes of those launches */
  int count;
} mt_prof_entry_t;
/* While enabled, every kernel launch is bracketed by CUDA events on its stream. */
void mt_prof_enable(int on);
void mt_prof_reset(void);
/* Waits for the recorded events; writes per-name aggregates sorted by time; returns #names. */
int mt_prof_collect(mt_prof_entry_t* out, int max_entries);
/* Kernels launched by this library in this process so far. */
unsigned long long mt_prof_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MINTIME_B200_H */
