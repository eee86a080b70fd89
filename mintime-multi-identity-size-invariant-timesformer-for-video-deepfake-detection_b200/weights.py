"""One-time weight packing: reference ``state_dict`` -> the device layouts of include/mintime_b200.h.

Runs when a checkpoint is loaded (not on the hot path): folds eval-mode BatchNorm into the
preceding convolution (reference model.py:62,73,86,165,197; eps utils.py:521), lays the depthwise /
stem filters out tap-major for NHWC kernels, pre-scales the query rows of ``to_qkv`` by
dim_head^-0.5 (size_invariant_timesformer.py:114, exact: a power of two) and interleaves the GEGLU
halves of ``net.0`` in blocks of 32 rows so one GEMM tile holds a value column and its gate.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import _lib
from .spec import B0_BLOCKS, BN_EPS


def _strip(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Accept DataParallel checkpoints ('module.' prefix, reference predict.py:378-388)."""
    if any(k.startswith("module.") for k in sd):
        return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    return sd


def _bn_fold(sd, prefix: str, eps: float = BN_EPS) -> Tuple[torch.Tensor, torch.Tensor]:
    g, b = sd[prefix + ".weight"].double(), sd[prefix + ".bias"].double()
    m, v = sd[prefix + ".running_mean"].double(), sd[prefix + ".running_var"].double()
    scale = g / torch.sqrt(v + eps)
    return scale, b - m * scale


class Packed:
    """Owns the packed device tensors and the ctypes struct pointing at them."""

    def __init__(self):
        self.keep: List[torch.Tensor] = []
        self.struct = None

    def dev(self, t: torch.Tensor, dtype, device) -> torch.Tensor:
        t = t.to(device=device, dtype=dtype).contiguous()
        assert t.data_ptr() % 16 == 0
        self.keep.append(t)
        return t

    def p(self, t: torch.Tensor, dtype, device) -> int:
        return self.dev(t, dtype, device).data_ptr()


def pack_effnet(sd: Dict[str, torch.Tensor], precision: str, device) -> Packed:
    sd = {k: v.detach().cpu() for k, v in _strip(sd).items()}
    T = _lib.torch_dtype(precision)
    f32 = torch.float32
    pk = Packed()
    W = _lib.EffnetWeights()

    def pw(conv_key: str, bn_prefix: str) -> _lib.PW:
        scale, shift = _bn_fold(sd, bn_prefix)
        w = sd[conv_key].double().flatten(1) * scale[:, None]          # [cout][cin]
        s = _lib.PW()
        s.w = pk.p(w, T, device)
        s.shift = pk.p(shift, f32, device)
        return s

    scale, shift = _bn_fold(sd, "_bn0")
    stem = sd["_conv_stem.weight"].double() * scale[:, None, None, None]   # (co, ci, ky, kx)
    W.stem_w = pk.p(stem.permute(2, 3, 1, 0).reshape(27, 32), f32, device)  # [(ky,kx,ci)][co]
    W.stem_shift = pk.p(shift, f32, device)
    for b in B0_BLOCKS:
        p = f"_blocks.{b.index}."
        blk = W.blocks[b.index]
        if b.expand != 1:
            blk.expand = pw(p + "_expand_conv.weight", p + "_bn0")
        scale, shift = _bn_fold(sd, p + "_bn1")
        dw = sd[p + "_depthwise_conv.weight"].double()[:, 0] * scale[:, None, None]   # (c, ky, kx)
        blk.dw_w = pk.p(dw.permute(1, 2, 0).reshape(b.kernel * b.kernel, b.cexp), f32, device)
        blk.dw_shift = pk.p(shift, f32, device)
        blk.se_reduce_w = pk.p(sd[p + "_se_reduce.weight"].flatten(1), f32, device)
        blk.se_reduce_b = pk.p(sd[p + "_se_reduce.bias"], f32, device)
        blk.se_expand_w = pk.p(sd[p + "_se_expand.weight"].flatten(1).t(), f32, device)     # [sq][cexp]
        blk.se_expand_b = pk.p(sd[p + "_se_expand.bias"], f32, device)
        blk.project = pw(p + "_project_conv.weight", p + "_bn2")
    W.head = pw("_conv_head.weight", "_bn1")
    pk.struct = W
    return pk


XC_BN_EPS = 1e-5          # nn.BatchNorm2d default: models/xception.py:91 sets BN = nn.BatchNorm2d with no eps argument


def pack_xception(sd: Dict[str, torch.Tensor], precision: str, device) -> Packed:
    """Xception ``state_dict`` (models/xception.py:93-137) -> mt_xception_weights_t: every BatchNorm folded into the 1x1 / dense
    convolution in front of it, dense 3x3 filters as im2col rows ((ky,kx,ci) columns, conv1 padded 27 -> 32), depthwise
    filters tap-major."""
    from .xception import XCEPTION_BLOCKS, sep_channels
    sd = {k: v.detach().cpu() for k, v in _strip(sd).items()}
    T = _lib.torch_dtype(precision)
    f32 = torch.float32
    pk = Packed()
    W = _lib.XceptionWeights()

    def fold(w2d: torch.Tensor, bn_prefix: str) -> _lib.PW:
        scale, shift = _bn_fold(sd, bn_prefix, XC_BN_EPS)
        s = _lib.PW()
        s.w = pk.p(w2d.double() * scale[:, None], T, device)
        s.shift = pk.p(shift, f32, device)
        return s

    def dense3x3(key: str, bn_prefix: str, kp: int) -> _lib.PW:
        w = sd[key]                                                     # (co, ci, ky, kx)
        rows = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)            # [co][(ky,kx,ci)]
        if rows.shape[1] < kp:
            rows = torch.cat([rows, torch.zeros((rows.shape[0], kp - rows.shape[1]), dtype=rows.dtype)], dim=1)
        return fold(rows, bn_prefix)

    def sep(dst, prefix: str, bn_prefix: str):
        dw = sd[prefix + ".conv1.weight"][:, 0]                          # (c, ky, kx)
        dst.dw_w = pk.p(dw.permute(1, 2, 0).reshape(9, dw.shape[0]), f32, device)
        dst.pw = fold(sd[prefix + ".pointwise.weight"].flatten(1), bn_prefix)

    W.conv1 = dense3x3("conv1.weight", "bn1", 32)
    W.conv2 = dense3x3("conv2.weight", "bn2", 288)
    unit, skip = 0, 0
    for bi, (cin, cout, reps, stride, relu0, grow_first) in enumerate(XCEPTION_BLOCKS):
        p = f"block{bi + 1}."
        idx = 1 if relu0 else 0                                          # position of the first SeparableConv2d in `rep`
        for _ in sep_channels(cin, cout, reps, grow_first):
            sep(W.sep[unit], f"{p}rep.{idx}", f"{p}rep.{idx + 1}")
            unit += 1
            idx += 3
        if cout != cin or stride != 1:
            W.skip[skip] = fold(sd[p + "skip.weight"].flatten(1), p + "skipbn")
            skip += 1
    sep(W.sep[unit], "conv3", "bn3")
    sep(W.sep[unit + 1], "conv4", "bn4")
    assert unit + 2 == 34 and skip == 4
    pk.struct = W
    return pk


def geglu_interleave(t: torch.Tensor) -> torch.Tensor:
    """rows [u(0..h) | g(0..h)] -> blocks of 64 = 32 u rows followed by their 32 gate rows."""
    h = t.shape[0] // 2
    assert h % 32 == 0
    u = t[:h].reshape(h // 32, 32, *t.shape[1:])
    g = t[h:].reshape(h // 32, 32, *t.shape[1:])
    return torch.cat([u, g], dim=1).reshape(t.shape)


def qkv_per_head(wqkv: torch.Tensor, heads: int, dim_head: int) -> torch.Tensor:
    """[3*inner][dim] (q rows | k rows | v rows) -> [heads][3*dim_head][dim]: q, k, v rows of head h contiguous -- the
    operand layout of the fused attention kernel (one 192-row B tile per head)."""
    inner = heads * dim_head
    q, k, v = (wqkv[i * inner:(i + 1) * inner].reshape(heads, dim_head, -1) for i in range(3))
    return torch.cat([q, k, v], dim=1).contiguous()


def tsf_cfg_struct(config: dict) -> _lib.TsfCfg:
    m = config["model"]
    c = _lib.TsfCfg()
    c.dim, c.depth, c.heads, c.dim_head = m["dim"], m["depth"], m["heads"], m["dim-head"]
    c.num_frames, c.num_patches, c.channels = m["num-frames"], m["num-patches"], m["channels"]
    c.num_classes = m["num-classes"]
    c.enable_pos_emb = int(bool(m["enable-pos-emb"]))
    c.enable_size_emb = int(bool(m["enable-size-emb"]))
    return c


def pack_tsf(sd: Dict[str, torch.Tensor], config: dict, precision: str, device) -> Packed:
    sd = {k: v.detach().cpu() for k, v in _strip(sd).items()}
    m = config["model"]
    T = _lib.torch_dtype(precision)
    f32 = torch.float32
    inner = m["heads"] * m["dim-head"]
    if m["depth"] > _lib.TSF_MAX_DEPTH:
        raise ValueError(f"depth {m['depth']} > {_lib.TSF_MAX_DEPTH}")
    pk = Packed()
    W = _lib.TsfWeights()
    W.w_patch = pk.p(sd["to_patch_embedding.weight"], T, device)
    W.b_patch = pk.p(sd["to_patch_embedding.bias"], f32, device)
    W.cls_token = pk.p(sd["cls_token"].reshape(-1), f32, device)
    W.pos_emb = pk.p(sd["pos_emb.weight"], f32, device)
    W.size_emb = pk.p(sd["size_emb.weight"], f32, device) if m["enable-size-emb"] else None
    for l in range(m["depth"]):
        for j, arr in ((0, W.time_attn), (1, W.space_attn)):
            p = f"layers.{l}.{j}."
            a = arr[l]
            a.ln_g = pk.p(sd[p + "norm.weight"], f32, device)
            a.ln_b = pk.p(sd[p + "norm.bias"], f32, device)
            wqkv = sd[p + "fn.to_qkv.weight"].clone()
            wqkv[:inner] *= m["dim-head"] ** -0.5
            a.w_qkv = pk.p(wqkv, T, device)
            if precision == "bf16":
                a.w_qkv_heads = pk.p(qkv_per_head(wqkv, m["heads"], m["dim-head"]), T, device)
            a.w_out = pk.p(sd[p + "fn.to_out.0.weight"], T, device)
            a.b_out = pk.p(sd[p + "fn.to_out.0.bias"], f32, device)
        p = f"layers.{l}.2."
        ff = W.ff[l]
        ff.ln_g = pk.p(sd[p + "norm.weight"], f32, device)
        ff.ln_b = pk.p(sd[p + "norm.bias"], f32, device)
        ff.w1 = pk.p(geglu_interleave(sd[p + "fn.net.0.weight"]), T, device)
        ff.b1 = pk.p(geglu_interleave(sd[p + "fn.net.0.bias"]), f32, device)
        ff.w2 = pk.p(sd[p + "fn.net.3.weight"], T, device)
        ff.b2 = pk.p(sd[p + "fn.net.3.bias"], f32, device)
    W.out_ln_g = pk.p(sd["to_out.0.weight"], f32, device)
    W.out_ln_b = pk.p(sd["to_out.0.bias"], f32, device)
    W.out_w = pk.p(sd["to_out.1.weight"], f32, device)
    W.out_b = pk.p(sd["to_out.1.bias"], f32, device)
    pk.struct = W
    return pk
