"""ctypes binding of libmintime_b200.so (C ABI in include/mintime_b200.h).

There is no fallback: if the shared library is missing or the device is not sm_100, every entry
point raises.  The library is built in-tree by ``__graft_entry__.build()`` / ``build.py``.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmintime_b200.so")

PREC_FP32, PREC_BF16 = 0, 1
ATTN_TIME, ATTN_SPACE = 0, 1
IN_F32, IN_U8 = 0, 1
TSF_MAX_DEPTH = 16

EXPORTS = [
    "mt_abi_version", "mt_last_error", "mt_device_check",
    "mt_effnet_b0_workspace_bytes", "mt_effnet_b0_fwd", "mt_xception_out_hw", "mt_xception_workspace_bytes", "mt_xception_fwd", "mt_tsf_workspace_bytes", "mt_tsf_fwd",
    "mt_pointwise_fwd", "mt_linear_residual_fwd", "mt_linear_geglu_fwd", "mt_patch_embed_fwd",
    "mt_layernorm_fwd", "mt_layernorm_copy_fwd", "mt_divided_attn_fwd", "mt_divided_attn_workspace_bytes", "mt_fused_attn_supported",
    "mt_fused_attn_workspace_bytes", "mt_fused_attn_fwd", "mt_stem_fwd", "mt_dwconv_fwd", "mt_se_gate_fwd", "mt_head_fwd",
    "mt_dwconv_chunks", "mt_expand_dwconv_chunks", "mt_expand_dwconv_fwd", "mt_effnet_b0_block_spec", "mt_mbconv_workspace_bytes", "mt_mbconv_fwd",
    "mt_aggregate_attn_fwd", "mt_clip_meta_fwd",
    "mt_grad_prep_workspace_bytes", "mt_grad_prep", "mt_linear_wgrad", "mt_linear_wgrad_nt", "mt_colsum_f32", "mt_layernorm_bwd_workspace_bytes", "mt_layernorm_bwd",
    "mt_geglu_fwd", "mt_geglu_bwd", "mt_geglu_bwd_colsum_workspace_bytes", "mt_geglu_bwd_colsum", "mt_divided_attn_bwd_workspace_bytes", "mt_divided_attn_bwd", "mt_embed_bwd",
    "mt_head_bwd_workspace_bytes", "mt_head_bwd",
    "mt_extractor_train_workspace_bytes", "mt_bn_stats", "mt_bn_act_fwd", "mt_bn_act_bwd", "mt_stem_raw_fwd", "mt_stem_wgrad",
    "mt_dwconv_raw_fwd", "mt_dwconv_dgrad", "mt_dwconv_wgrad", "mt_group_mean", "mt_se_fc_fwd", "mt_se_fc_bwd", "mt_gate_mul",
    "mt_gate_bwd", "mt_scale_add", "mt_conv1x1_wgrad_workspace_bytes", "mt_conv1x1_wgrad", "mt_prof_enable", "mt_prof_reset", "mt_prof_collect", "mt_prof_launch_count",
]

vp, fp, i32, sz = C.c_void_p, C.c_void_p, C.c_int, C.c_size_t   # all device pointers travel as void*


class PW(C.Structure):
    _fields_ = [("w", vp), ("shift", fp)]


class MBConv(C.Structure):
    _fields_ = [("expand", PW), ("dw_w", fp), ("dw_shift", fp), ("se_reduce_w", fp), ("se_reduce_b", fp),
                ("se_expand_w", fp), ("se_expand_b", fp), ("project", PW)]


class EffnetWeights(C.Structure):
    _fields_ = [("stem_w", fp), ("stem_shift", fp), ("blocks", MBConv * 16), ("head", PW)]


class XcSep(C.Structure):
    _fields_ = [("dw_w", fp), ("pw", PW)]


class XceptionWeights(C.Structure):
    _fields_ = [("conv1", PW), ("conv2", PW), ("sep", XcSep * 34), ("skip", PW * 4)]


class AttnWeights(C.Structure):
    _fields_ = [("ln_g", fp), ("ln_b", fp), ("w_qkv", vp), ("w_out", vp), ("b_out", fp), ("w_qkv_heads", vp)]


class FFWeights(C.Structure):
    _fields_ = [("ln_g", fp), ("ln_b", fp), ("w1", vp), ("b1", fp), ("w2", vp), ("b2", fp)]


class TsfCfg(C.Structure):
    _fields_ = [(k, i32) for k in ("dim", "depth", "heads", "dim_head", "num_frames", "num_patches", "channels",
                                   "num_classes", "enable_pos_emb", "enable_size_emb")]


class TsfWeights(C.Structure):
    _fields_ = [("w_patch", vp), ("b_patch", fp), ("cls_token", fp), ("pos_emb", fp), ("size_emb", fp),
                ("time_attn", AttnWeights * TSF_MAX_DEPTH), ("space_attn", AttnWeights * TSF_MAX_DEPTH),
                ("ff", FFWeights * TSF_MAX_DEPTH), ("out_ln_g", fp), ("out_ln_b", fp), ("out_w", fp), ("out_b", fp)]


class MBConvSpec(C.Structure):
    _fields_ = [(k, i32) for k in ("kernel", "stride", "expand", "cin", "cout", "hw_in")]


class ProfEntry(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("ms_total", C.c_double), ("flops_total", C.c_double),
                ("bytes_total", C.c_double), ("count", i32)]


_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"mintime_b200: {LIB_PATH} is missing -- build it with `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (nvcc, sm_100a).  There is no CPU / PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    lib.mt_abi_version.restype = i32
    lib.mt_last_error.restype = C.c_char_p
    lib.mt_device_check.restype = i32
    lib.mt_effnet_b0_workspace_bytes.restype = sz
    lib.mt_effnet_b0_workspace_bytes.argtypes = [i32, i32]
    lib.mt_effnet_b0_fwd.argtypes = [C.POINTER(EffnetWeights), vp, i32, vp, i32, i32, vp, sz, vp]
    lib.mt_xception_out_hw.argtypes = [i32]
    lib.mt_xception_out_hw.restype = i32
    lib.mt_xception_workspace_bytes.argtypes = [i32, i32, i32]
    lib.mt_xception_workspace_bytes.restype = sz
    lib.mt_xception_fwd.argtypes = [C.POINTER(XceptionWeights), vp, i32, vp, i32, i32, i32, vp, sz, vp]
    lib.mt_tsf_workspace_bytes.restype = sz
    lib.mt_tsf_workspace_bytes.argtypes = [C.POINTER(TsfCfg), i32, i32]
    lib.mt_tsf_fwd.argtypes = [C.POINTER(TsfWeights), C.POINTER(TsfCfg), vp, vp, vp, vp, vp, vp, vp, vp, i32, i32,
                               vp, sz, vp]
    lib.mt_pointwise_fwd.argtypes = [i32, vp, vp, fp, fp, i32, vp, i32, vp, i32, i32, i32, vp]
    lib.mt_linear_residual_fwd.argtypes = [i32, vp, vp, fp, fp, i32, i32, i32, vp]
    lib.mt_linear_geglu_fwd.argtypes = [i32, vp, vp, fp, vp, i32, i32, i32, vp]
    lib.mt_patch_embed_fwd.argtypes = [i32, C.POINTER(TsfWeights), C.POINTER(TsfCfg), vp, vp, vp, fp, i32, vp]
    lib.mt_layernorm_fwd.argtypes = [i32, fp, fp, fp, vp, i32, i32, vp]
    lib.mt_layernorm_copy_fwd.argtypes = [i32, fp, fp, fp, vp, fp, i32, i32, vp]
    lib.mt_divided_attn_fwd.argtypes = [i32, vp, vp, vp, i32, vp, fp, i32, i32, i32, i32, i32, vp, sz, vp]
    lib.mt_divided_attn_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.mt_divided_attn_workspace_bytes.restype = sz
    lib.mt_fused_attn_supported.argtypes = [i32, i32, i32, i32, i32]
    lib.mt_fused_attn_supported.restype = i32
    lib.mt_fused_attn_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.mt_fused_attn_workspace_bytes.restype = sz
    lib.mt_fused_attn_fwd.argtypes = [vp, vp, vp, vp, i32, vp, fp, i32, i32, i32, i32, i32, i32, vp, sz, vp]
    lib.mt_stem_fwd.argtypes = [i32, vp, i32, fp, fp, vp, i32, i32, i32, vp]
    lib.mt_dwconv_fwd.argtypes = [i32, vp, fp, fp, vp, fp, i32, i32, i32, i32, i32, i32, vp]
    lib.mt_se_gate_fwd.argtypes = [fp, i32, i32, fp, fp, fp, fp, fp, i32, i32, i32, vp]
    lib.mt_expand_dwconv_chunks.argtypes = [i32, i32, i32, i32, i32]
    lib.mt_expand_dwconv_chunks.restype = i32
    lib.mt_expand_dwconv_fwd.argtypes = [vp, vp, fp, fp, fp, vp, fp, i32, i32, i32, i32, i32, i32, vp]
    lib.mt_dwconv_chunks.argtypes = [i32, i32, i32, i32, i32, i32]
    lib.mt_dwconv_chunks.restype = i32
    lib.mt_effnet_b0_block_spec.argtypes = [i32, C.POINTER(MBConvSpec)]
    lib.mt_effnet_b0_block_spec.restype = i32
    lib.mt_mbconv_workspace_bytes.argtypes = [C.POINTER(MBConvSpec), i32, i32]
    lib.mt_mbconv_workspace_bytes.restype = sz
    lib.mt_mbconv_fwd.argtypes = [i32, C.POINTER(MBConvSpec), C.POINTER(MBConv), vp, vp, i32, vp, sz, vp]
    lib.mt_head_fwd.argtypes = [fp, fp, fp, fp, fp, fp, i32, i32, i32, i32, vp]
    lib.mt_aggregate_attn_fwd.argtypes = [fp, fp, fp, i32, i32, i32, i32, C.c_float, vp]
    lib.mt_clip_meta_fwd.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.mt_grad_prep_workspace_bytes.argtypes = [i32, i32]
    lib.mt_grad_prep_workspace_bytes.restype = sz
    lib.mt_grad_prep.argtypes = [i32, vp, i32, vp, vp, fp, i32, i32, i32, i32, vp, sz, vp]
    lib.mt_colsum_f32.argtypes = [fp, fp, i32, i32, i32, vp]
    lib.mt_linear_wgrad.argtypes = [i32, vp, vp, fp, i32, i32, i32, vp]
    lib.mt_linear_wgrad_nt.argtypes = [i32, vp, vp, fp, i32, i32, i32, vp]
    lib.mt_layernorm_bwd_workspace_bytes.argtypes = [i32, i32]
    lib.mt_layernorm_bwd_workspace_bytes.restype = sz
    lib.mt_layernorm_bwd.argtypes = [i32, fp, fp, vp, fp, fp, i32, i32, vp, sz, vp]
    lib.mt_geglu_fwd.argtypes = [i32, vp, vp, i32, i32, vp]
    lib.mt_geglu_bwd.argtypes = [i32, vp, vp, vp, i32, i32, vp]
    lib.mt_geglu_bwd_colsum_workspace_bytes.argtypes = [i32, i32]
    lib.mt_geglu_bwd_colsum_workspace_bytes.restype = C.c_size_t
    lib.mt_geglu_bwd_colsum.argtypes = [i32, vp, vp, vp, fp, i32, i32, vp, C.c_size_t, vp]
    lib.mt_geglu_bwd_colsum.restype = i32
    lib.mt_divided_attn_bwd_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.mt_divided_attn_bwd_workspace_bytes.restype = sz
    lib.mt_divided_attn_bwd.argtypes = [i32, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, vp, sz, vp]
    lib.mt_embed_bwd.argtypes = [fp, vp, vp, fp, fp, fp, i32, i32, i32, i32, i32, vp]
    lib.mt_head_bwd_workspace_bytes.argtypes = [i32, i32, i32]
    lib.mt_head_bwd_workspace_bytes.restype = sz
    lib.mt_head_bwd.argtypes = [fp, fp, fp, fp, fp, fp, fp, i32, i32, i32, i32, vp, sz, vp]
    i64, f32 = C.c_longlong, C.c_float
    lib.mt_extractor_train_workspace_bytes.argtypes = [i64, i32]
    lib.mt_extractor_train_workspace_bytes.restype = sz
    lib.mt_bn_stats.argtypes = [fp, fp, fp, fp, fp, f32, i64, i32, vp, sz, vp]
    lib.mt_bn_act_fwd.argtypes = [fp, fp, fp, fp, fp, i32, f32, fp, i64, i32, vp]
    lib.mt_bn_act_bwd.argtypes = [fp, fp, fp, fp, fp, fp, i32, f32, fp, fp, fp, i64, i32, vp, sz, vp]
    lib.mt_stem_raw_fwd.argtypes = [fp, fp, fp, i32, i32, i32, vp]
    lib.mt_stem_wgrad.argtypes = [fp, fp, fp, i32, i32, i32, vp, sz, vp]
    lib.mt_dwconv_raw_fwd.argtypes = [fp, fp, fp, i32, i32, i32, i32, i32, vp]
    lib.mt_dwconv_dgrad.argtypes = [fp, fp, fp, i32, i32, i32, i32, i32, vp]
    lib.mt_dwconv_wgrad.argtypes = [fp, fp, fp, i32, i32, i32, i32, i32, vp, sz, vp]
    lib.mt_conv1x1_wgrad_workspace_bytes.argtypes = [i64, i32, i32]
    lib.mt_conv1x1_wgrad_workspace_bytes.restype = sz
    lib.mt_conv1x1_wgrad.argtypes = [fp, fp, fp, i64, i32, i32, vp, sz, vp]
    lib.mt_group_mean.argtypes = [fp, fp, i32, i32, i32, vp]
    lib.mt_se_fc_fwd.argtypes = [fp, fp, fp, fp, fp, fp, fp, i32, i32, i32, vp]
    lib.mt_se_fc_bwd.argtypes = [fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, i32, i32, i32, vp, sz, vp]
    lib.mt_gate_mul.argtypes = [fp, fp, fp, i32, i32, i32, vp]
    lib.mt_gate_bwd.argtypes = [fp, fp, fp, fp, fp, fp, i32, i32, i32, i32, vp]
    lib.mt_scale_add.argtypes = [fp, fp, fp, fp, i32, i64, vp]
    lib.mt_prof_enable.argtypes = [i32]
    lib.mt_prof_enable.restype = None
    lib.mt_prof_reset.restype = None
    lib.mt_prof_collect.argtypes = [C.POINTER(ProfEntry), i32]
    lib.mt_prof_collect.restype = i32
    lib.mt_prof_launch_count.restype = C.c_ulonglong
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name.endswith(("_fwd", "_bwd")) or name in ("mt_grad_prep", "mt_colsum_f32", "mt_linear_wgrad", "mt_linear_wgrad_nt", "mt_bn_stats",
                                                       "mt_stem_wgrad", "mt_dwconv_dgrad", "mt_dwconv_wgrad", "mt_group_mean",
                                                       "mt_gate_mul", "mt_scale_add", "mt_conv1x1_wgrad"):
            fn.restype = i32
    if lib.mt_abi_version() != 5:
        raise RuntimeError("mintime_b200: ABI version mismatch between _lib.py and libmintime_b200.so")
    _lib = lib
    return lib


class MintimeError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mt_last_error().decode("utf-8", "replace")
        raise MintimeError(f"{what or 'mintime_b200'} failed (status {rc}): {msg}")


_device_ok = set()


def require_device(device) -> None:
    """Raise unless `device` is a CUDA device the library supports (compute capability 10.x)."""
    import torch
    if isinstance(device, torch.device) and device.type == "cuda" and device.index in _device_ok:
        return                      # hot path of the eager training step: ~500 calls per step
    if not torch.cuda.is_available():
        raise MintimeError("mintime_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise MintimeError(f"mintime_b200 tensors must live on a CUDA device, got {dev}")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx in _device_ok:
        return
    with torch.cuda.device(idx):
        check(load().mt_device_check(), "mt_device_check")
    _device_ok.add(idx)


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def prec_id(precision: str) -> int:
    try:
        return {"fp32": PREC_FP32, "bf16": PREC_BF16}[precision]
    except KeyError:
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")


def torch_dtype(precision: str):
    import torch
    return torch.float32 if precision == "fp32" else torch.bfloat16


def profile_collect(max_entries: int = 256):
    """[(name, ms_total, flops_total, bytes_total, count)] sorted by time, from the mt_prof_* hooks."""
    lib = load()
    buf = (ProfEntry * max_entries)()
    n = min(lib.mt_prof_collect(buf, max_entries), max_entries)
    return [(buf[i].name.decode(), buf[i].ms_total, buf[i].flops_total, buf[i].bytes_total, buf[i].count)
            for i in range(n)]
