"""CPU ORACLE for the MINTIME hot path -- TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The shipped path
(``mintime_b200``) never imports it and has no CPU fallback.

It is a flat, functional fp32 restatement (plain torch CPU ops on a ``state_dict``; no nn.Module,
no einops) of the reference's algorithm for

    EfficientNet-B0.forward -> SizeInvariantTimeSformer.forward

Each function cites the reference file:line it follows (paths relative to the reference repo).
The arithmetic itself lives in a third-party dependency of the reference (PyTorch, pinned
torch==1.11.0 in requirements.txt:111; einops 0.4.1) -- this oracle calls the same primitive ops
(conv2d / linear / softmax / layer_norm / gelu-erf) on the same values in the same order.

Parity pinning: the reference holds no golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced by ``oracle/make_golden.py``
importing the unmodified reference modules from /root/reference (fixtures in ``tests/golden/``);
``tests/test_oracle.py`` checks oracle == fixtures (max-abs <= 2e-6 relative to tensor scale).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-3  # models/efficientnet/efficientnet_pytorch/utils.py:521

# models/efficientnet/efficientnet_pytorch/utils.py:502-510
B0_BLOCK_STRINGS = [
    'r1_k3_s11_e1_i32_o16_se0.25', 'r2_k3_s22_e6_i16_o24_se0.25', 'r2_k5_s22_e6_i24_o40_se0.25',
    'r3_k3_s22_e6_i40_o80_se0.25', 'r3_k5_s11_e6_i80_o112_se0.25', 'r4_k5_s22_e6_i112_o192_se0.25',
    'r1_k3_s11_e6_i192_o320_se0.25',
]


def decode_blocks() -> List[dict]:
    """utils.py:372-395 (string -> args) + model.py:171-191 (repeat expansion: later repeats get
    stride 1 and cin = cout)."""
    blocks = []
    for s in B0_BLOCK_STRINGS:
        o = {}
        for op in s.split('_'):
            key = op[0] if not op.startswith('se') else 'se'
            o[key] = op[len(key):]
        r, k, st, e = int(o['r']), int(o['k']), int(o['s'][0]), int(o['e'])
        ci, co = int(o['i']), int(o['o'])
        for j in range(r):
            # skip rule, model.py:123: id_skip and stride == 1 and cin == cout.  For the first block
            # of a stage ``stride`` is still the list [1] (utils.py:394) so ``[1] == 1`` is False
            # there; repeats (j > 0) have stride 1 and cin == cout by construction (model.py:186-187).
            blocks.append(dict(k=k, s=st if j == 0 else 1, e=e, cin=ci if j == 0 else co, cout=co,
                               skip=(j > 0)))
    return blocks


def same_pad(x: Tensor, k: int, s: int) -> Tensor:
    """Conv2dStaticSamePadding, utils.py:254-276: ZeroPad2d((pw//2, pw-pw//2, ph//2, ph-ph//2))."""
    ih, iw = x.shape[-2:]
    oh, ow = math.ceil(ih / s), math.ceil(iw / s)
    ph = max((oh - 1) * s + k - ih, 0)
    pw = max((ow - 1) * s + k - iw, 0)
    if ph > 0 or pw > 0:
        x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    return x


def swish(x: Tensor) -> Tensor:
    """SwishImplementation.forward, utils.py:66-69."""
    return x * torch.sigmoid(x)


def bn_eval(x: Tensor, sd: Dict[str, Tensor], p: str) -> Tensor:
    """nn.BatchNorm2d in eval mode (model.py:62,73,86,165,197), eps 1e-3."""
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'], sd[p + '.weight'],
                        sd[p + '.bias'], False, 0.0, BN_EPS)


def mbconv(x: Tensor, sd: Dict[str, Tensor], p: str, b: dict) -> Tensor:
    """MBConvBlock.forward, model.py:89-128 (eval: drop_connect is the identity, utils.py:141-142)."""
    inp = x
    if b['e'] != 1:                                                         # model.py:100-103
        x = swish(bn_eval(F.conv2d(x, sd[p + '_expand_conv.weight']), sd, p + '_bn0'))
    cexp = x.shape[1]
    x = F.conv2d(same_pad(x, b['k'], b['s']), sd[p + '_depthwise_conv.weight'], None, b['s'], 0, 1, cexp)
    x = swish(bn_eval(x, sd, p + '_bn1'))                                   # model.py:105-107
    sq = F.adaptive_avg_pool2d(x, 1)                                        # model.py:110-115
    sq = swish(F.conv2d(sq, sd[p + '_se_reduce.weight'], sd[p + '_se_reduce.bias']))
    sq = F.conv2d(sq, sd[p + '_se_expand.weight'], sd[p + '_se_expand.bias'])
    x = torch.sigmoid(sq) * x
    x = bn_eval(F.conv2d(x, sd[p + '_project_conv.weight']), sd, p + '_bn2')  # model.py:118-119
    if b['skip']:                                                           # model.py:122-127
        x = x + inp
    return x


def effnet_b0_forward(sd: Dict[str, Tensor], x: Tensor, taps: Optional[dict] = None) -> Tensor:
    """EfficientNet.forward, model.py:267-288: stem -> 16 MBConv -> head; returns the 1280x7x7 map.

    ``x``: (n,3,224,224) fp32, raw 0..255 (no normalisation anywhere on the path)."""
    x = swish(bn_eval(F.conv2d(same_pad(x, 3, 2), sd['_conv_stem.weight'], None, 2), sd, '_bn0'))
    if taps is not None:
        taps['stem'] = x
    for i, b in enumerate(decode_blocks()):
        x = mbconv(x, sd, f'_blocks.{i}.', b)
        if taps is not None:
            taps[f'block{i}'] = x
    x = swish(bn_eval(F.conv2d(x, sd['_conv_head.weight']), sd, '_bn1'))
    if taps is not None:
        taps['head'] = x
    return x


# ----------------------------------------------------------------------------------------------
# EfficientNet-B0 in TRAIN mode (train.py:155-170 unfrozen extractor).  Not yet built in the
# product (the shim raises); this is the pinned checker the backward kernels of the extractor will be
# held to: BatchNorm with batch statistics + running-stat update, drop-connect, autograd for gradients.
# ----------------------------------------------------------------------------------------------
BN_MOMENTUM = 0.01           # 1 - batch_norm_momentum (0.99), model.py:47 / utils.py:520
DROP_CONNECT_RATE = 0.2      # utils.py:523


def bn_train(x: Tensor, sd: Dict[str, Tensor], p: str, new_stats: Optional[dict]) -> Tensor:
    """nn.BatchNorm2d in train mode: normalise with the batch's biased variance; running stats move by
    momentum 0.01 towards the batch mean / UNBIASED variance (returned in ``new_stats``, ``sd`` untouched)."""
    if new_stats is not None:
        with torch.no_grad():
            n = x.numel() // x.shape[1]
            mean = x.mean(dim=(0, 2, 3))
            var_u = x.var(dim=(0, 2, 3), unbiased=True) if n > 1 else torch.zeros_like(mean)
            new_stats[p + '.running_mean'] = (1 - BN_MOMENTUM) * sd[p + '.running_mean'] + BN_MOMENTUM * mean
            new_stats[p + '.running_var'] = (1 - BN_MOMENTUM) * sd[p + '.running_var'] + BN_MOMENTUM * var_u
    return F.batch_norm(x, None, None, sd[p + '.weight'], sd[p + '.bias'], True, 0.0, BN_EPS)


def drop_connect(x: Tensor, p: float) -> Tensor:
    """utils.py:129-154 (training): per-sample keep mask floor(keep_prob + U[0,1)), survivors scaled by 1/keep_prob.
    Draws from torch's global RNG exactly like the reference (same call, same shape, same order)."""
    keep = 1 - p
    r = keep + torch.rand([x.shape[0], 1, 1, 1], dtype=x.dtype, device=x.device)
    return x / keep * torch.floor(r)


def mbconv_train(x: Tensor, sd, p: str, b: dict, drop_rate: float, new_stats: Optional[dict]) -> Tensor:
    """MBConvBlock.forward in train mode, model.py:89-128."""
    inp = x
    if b['e'] != 1:
        x = swish(bn_train(F.conv2d(x, sd[p + '_expand_conv.weight']), sd, p + '_bn0', new_stats))
    cexp = x.shape[1]
    x = F.conv2d(same_pad(x, b['k'], b['s']), sd[p + '_depthwise_conv.weight'], None, b['s'], 0, 1, cexp)
    x = swish(bn_train(x, sd, p + '_bn1', new_stats))
    sq = F.adaptive_avg_pool2d(x, 1)
    sq = swish(F.conv2d(sq, sd[p + '_se_reduce.weight'], sd[p + '_se_reduce.bias']))
    sq = F.conv2d(sq, sd[p + '_se_expand.weight'], sd[p + '_se_expand.bias'])
    x = torch.sigmoid(sq) * x
    x = bn_train(F.conv2d(x, sd[p + '_project_conv.weight']), sd, p + '_bn2', new_stats)
    if b['skip']:
        if drop_rate:                                                       # model.py:125-126
            x = drop_connect(x, drop_rate)
        x = x + inp
    return x


def effnet_b0_forward_train(sd: Dict[str, Tensor], x: Tensor, drop_connect_rate: float = DROP_CONNECT_RATE,
                            new_stats: Optional[dict] = None) -> Tensor:
    """EfficientNet.forward with ``.train()`` (model.py:267-288): per-block drop-connect rate
    ``drop_connect_rate * idx / 16`` (model.py:279-282).  ``new_stats`` receives the updated BatchNorm running stats."""
    x = swish(bn_train(F.conv2d(same_pad(x, 3, 2), sd['_conv_stem.weight'], None, 2), sd, '_bn0', new_stats))
    blocks = decode_blocks()
    for i, b in enumerate(blocks):
        rate = drop_connect_rate * float(i) / len(blocks) if drop_connect_rate else 0.0
        x = mbconv_train(x, sd, f'_blocks.{i}.', b, rate, new_stats)
    return swish(bn_train(F.conv2d(x, sd['_conv_head.weight']), sd, '_bn1', new_stats))


# ----------------------------------------------------------------------------------------------
# Size-Invariant TimeSformer
# ----------------------------------------------------------------------------------------------
def _ln(x: Tensor, sd, p: str) -> Tensor:
    """PreNorm / nn.LayerNorm(dim), size_invariant_timesformer.py:18-26 (eps 1e-5)."""
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def _softmax_attn(q: Tensor, k: Tensor, v: Tensor, allow: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """attn(), size_invariant_timesformer.py:80-87: masked_fill(~mask, -FLT_MAX) then softmax."""
    sim = q @ k.transpose(-1, -2)
    if allow is not None:
        sim = sim.masked_fill(~allow, -torch.finfo(sim.dtype).max)
    a = sim.softmax(dim=-1)
    return a @ v, a


def divided_attention(xn: Tensor, sd, p: str, mode: str, f: int, n: int, heads: int,
                      mask: Tensor, idmask: Tensor) -> Tuple[Tensor, Tensor]:
    """Attention.forward, size_invariant_timesformer.py:109-144, written per (b, h) with explicit
    indices instead of einops regrouping.

    xn: (B, 1+f*n, D) already layer-normed.  Returns (to_out(...) (B,N,D), cls_attn (B*H,1,N)).
    """
    B, N, D = xn.shape
    Wqkv = sd[p + 'fn.to_qkv.weight']
    inner = Wqkv.shape[0] // 3
    dh = inner // heads
    qkv = xn @ Wqkv.t()                                                     # :111
    q, k, v = [t.view(B, N, heads, dh).permute(0, 2, 1, 3) for t in qkv.split(inner, dim=-1)]  # (B,H,N,dh)
    q = q * dh ** -0.5                                                      # :114
    # CLS query attends every key; keys of padded frames are masked (cls_attn_mask :258-260)
    cls_allow = torch.cat([torch.ones(B, 1, dtype=torch.bool), mask.repeat_interleave(n, dim=1)], 1)
    cls_out, cls_att = _softmax_attn(q[:, :, :1], k, v, cls_allow[:, None, None, :])   # :120
    qp = q[:, :, 1:].reshape(B, heads, f, n, dh)
    kp = k[:, :, 1:].reshape(B, heads, f, n, dh)
    vp = v[:, :, 1:].reshape(B, heads, f, n, dh)
    kc, vc = k[:, :, :1], v[:, :, :1]                                       # CLS key/value, :125-129
    if mode == 'time':      # groups (b,h,patch): queries = frames, keys = CLS + frames
        qg, kg, vg = [t.permute(0, 1, 3, 2, 4) for t in (qp, kp, vp)]       # (B,H,n,f,dh)
        kg = torch.cat([kc[:, :, None].expand(B, heads, n, 1, dh), kg], 3)
        vg = torch.cat([vc[:, :, None].expand(B, heads, n, 1, dh), vg], 3)
        # frame_mask :252-255: allow[b,q,k] = mask[b,k] & identities_mask[b,q,k]; CLS key always on
        allow = mask[:, None, :] & idmask                                   # (B,f,f)
        allow = torch.cat([torch.ones(B, f, 1, dtype=torch.bool), allow], 2)
        og, _ = _softmax_attn(qg, kg, vg, allow[:, None, None])
        o = og.permute(0, 1, 3, 2, 4).reshape(B, heads, f * n, dh)
    else:                   # groups (b,h,frame): queries = patches, keys = CLS + patches; no mask (:266)
        kg = torch.cat([kc[:, :, None].expand(B, heads, f, 1, dh), kp], 3)
        vg = torch.cat([vc[:, :, None].expand(B, heads, f, 1, dh), vp], 3)
        og, _ = _softmax_attn(qp, kg, vg, None)
        o = og.reshape(B, heads, f * n, dh)
    o = torch.cat([cls_out, o], 2)                                          # :138
    o = o.permute(0, 2, 1, 3).reshape(B, N, inner)                          # :141
    y = o @ sd[p + 'fn.to_out.0.weight'].t() + sd[p + 'fn.to_out.0.bias']   # :144
    return y, cls_att.reshape(B * heads, 1, N)


def feed_forward(xn: Tensor, sd, p: str) -> Tensor:
    """FeedForward/GEGLU, size_invariant_timesformer.py:60-76: Linear -> x*gelu(gates) -> Linear."""
    h = xn @ sd[p + 'fn.net.0.weight'].t() + sd[p + 'fn.net.0.bias']
    a, g = h.chunk(2, dim=-1)
    return (a * F.gelu(g)) @ sd[p + 'fn.net.3.weight'].t() + sd[p + 'fn.net.3.bias']


def tsf_embed(sd, cfg: dict, feats: Tensor, size_embedding: Tensor, positions: Tensor) -> Tensor:
    """size_invariant_timesformer.py:225-248: tokens + CLS, += pos_emb[positions], += size_emb[idx]."""
    B, f, C, h, w = feats.shape
    n = h * w
    x = feats.permute(0, 1, 3, 4, 2).reshape(B, f * n, C)                   # :227 'b f c h w -> b (f h w) c'
    tok = x @ sd['to_patch_embedding.weight'].t() + sd['to_patch_embedding.bias']
    x = torch.cat([sd['cls_token'][None].expand(B, 1, -1), tok], 1)         # :231-232
    m = cfg['model']
    if m['enable-pos-emb']:
        x = x + sd['pos_emb.weight'][positions]                             # :236
    else:
        x = x + sd['pos_emb.weight'][torch.arange(x.shape[1])]
    if m['enable-size-emb']:                                                # :241-248
        idx = torch.cat([torch.zeros(B, 1, dtype=torch.long),
                         size_embedding.long().repeat_interleave(n, dim=1)], 1)
        x = x + sd['size_emb.weight'][idx]
    return x


def tsf_forward(sd, cfg: dict, feats: Tensor, mask: Tensor, identities_mask: Tensor,
                size_embedding: Tensor, positions: Tensor, taps: Optional[dict] = None):
    """SizeInvariantTimeSformer.forward, size_invariant_timesformer.py:224-276.

    Returns (logits (B,num_classes), [space_attn, time_attn]) -- attention maps of the LAST layer
    only, in that order (:271), each (B*H,1,N)."""
    m = cfg['model']
    B, f, C, h, w = feats.shape
    n, heads = h * w, m['heads']
    assert f == m['num-frames']                                             # :252 uses self.num_frames
    x = tsf_embed(sd, cfg, feats, size_embedding, positions)
    if taps is not None:
        taps['embed'] = x
    ta = sa = None
    for l in range(m['depth']):                                             # :263-268
        y, ta = divided_attention(_ln(x, sd, f'layers.{l}.0.norm'), sd, f'layers.{l}.0.', 'time',
                                  f, n, heads, mask, identities_mask)
        x = x + y
        if taps is not None:
            taps[f'layer{l}.time'] = x
        y, sa = divided_attention(_ln(x, sd, f'layers.{l}.1.norm'), sd, f'layers.{l}.1.', 'space',
                                  f, n, heads, mask, identities_mask)
        x = x + y
        if taps is not None:
            taps[f'layer{l}.space'] = x
        x = feed_forward(_ln(x, sd, f'layers.{l}.2.norm'), sd, f'layers.{l}.2.') + x
        if taps is not None:
            taps[f'layer{l}.ff'] = x
    cls = x[:, 0]
    logits = _ln(cls, sd, 'to_out.0') @ sd['to_out.1.weight'].t() + sd['to_out.1.bias']  # :195-198,273
    return logits, [sa, ta]


def hot_path_forward(esd, tsd, cfg, frames_nhwc: Tensor, mask, identities_mask, size_embedding, positions):
    """The callers' loop body: train.py:341-355 / predict.py:401-406.

    frames_nhwc: (B,f,224,224,3) fp32 0..255."""
    B, f = frames_nhwc.shape[:2]
    x = frames_nhwc.permute(0, 1, 4, 2, 3).reshape(B * f, 3, *frames_nhwc.shape[2:4])   # train.py:341
    feats = effnet_b0_forward(esd, x)
    feats = feats.view(B, f, *feats.shape[1:])                                           # train.py:354
    return tsf_forward(tsd, cfg, feats, mask, identities_mask, size_embedding, positions)
