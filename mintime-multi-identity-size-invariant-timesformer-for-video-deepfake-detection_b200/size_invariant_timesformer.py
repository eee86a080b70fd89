"""Drop-in for the reference's ``SizeInvariantTimeSformer`` (models/size_invariant_timesformer.py).

Same constructor (``config=dict`` with the hyphenated yaml keys, ``require_attention``), same
``forward(x, mask=, identities_mask=, size_embedding=, positions=)`` and return values, same
``state_dict`` names/shapes and initialisation (:148-214).  The sub-modules only hold parameters;
the forward is one call into libmintime_b200.so (mt_tsf_fwd).  When gradients are enabled and a parameter requires
them (train.py:332-378) the forward/backward schedule of ``training.py`` runs instead, as one autograd node, so
``loss.backward()`` fills ``param.grad`` from the library's backward kernels.  There is no PyTorch fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn
from torch.nn.init import trunc_normal_

from . import _lib, weights


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("parameter container; call SizeInvariantTimeSformer.forward")


class GEGLU(_Holder):
    pass


class PreNorm(_Holder):          # :18-26
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)


class FeedForward(_Holder):      # :65-76
    def __init__(self, dim, mult=4, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * mult * 2), GEGLU(), nn.Dropout(dropout),
                                 nn.Linear(dim * mult, dim))


class Attention(_Holder):        # :89-106
    def __init__(self, dim, dim_head=64, heads=8, dropout=0.0):
        super().__init__()
        self.heads = heads
        self.scale = dim_head ** -0.5
        inner_dim = dim_head * heads
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim), nn.Dropout(dropout))


class SizeInvariantTimeSformer(nn.Module):
    def __init__(self, *, config, require_attention=False, precision: str = "bf16"):
        super().__init__()
        m = config["model"]
        self.config = config
        self.dim = m["dim"]
        self.num_frames = m["num-frames"]
        self.max_identities = m["max-identities"]
        self.image_size = m["image-size"]
        self.num_classes = m["num-classes"]
        self.patch_size = m["patch-size"]
        self.num_patches = m["num-patches"]
        self.channels = m["channels"]
        self.depth = m["depth"]
        self.heads = m["heads"]
        self.dim_head = m["dim-head"]
        self.attn_dropout = m["attn-dropout"]
        self.ff_dropout = m["ff-dropout"]
        self.shift_tokens = m["shift-tokens"]
        self.enable_size_emb = m["enable-size-emb"]
        self.enable_pos_emb = m["enable-pos-emb"]
        self.require_attention = require_attention
        self.precision = precision
        if self.shift_tokens:
            # the reference itself raises NameError here (:189, `num_frames` undefined)
            raise NotImplementedError("shift-tokens: True is broken in the reference (NameError at "
                                      "size_invariant_timesformer.py:189) and not supported")
        if self.attn_dropout or self.ff_dropout:
            # (the shipped config/size_invariant_timesformer.yaml has attn-dropout = ff-dropout = 0)
            raise NotImplementedError("dropout > 0 is not supported")

        num_positions = self.num_frames * self.channels                        # :173
        self.to_patch_embedding = nn.Linear(self.channels, self.dim)
        self.cls_token = nn.Parameter(torch.randn(1, self.dim))
        self.pos_emb = nn.Embedding(num_positions + 1, self.dim)
        if self.enable_size_emb:
            self.size_emb = nn.Embedding(num_positions + 1, self.dim)
        self.layers = nn.ModuleList([])
        for _ in range(self.depth):
            ff = FeedForward(self.dim, dropout=self.ff_dropout)
            time_attn = Attention(self.dim, dim_head=self.dim_head, heads=self.heads, dropout=self.attn_dropout)
            spatial_attn = Attention(self.dim, dim_head=self.dim_head, heads=self.heads, dropout=self.attn_dropout)
            self.layers.append(nn.ModuleList([PreNorm(self.dim, t) for t in (time_attn, spatial_attn, ff)]))
        self.to_out = nn.Sequential(nn.LayerNorm(self.dim), nn.Linear(self.dim, self.num_classes))

        trunc_normal_(self.pos_emb.weight, std=.02)                            # :200-205
        trunc_normal_(self.cls_token, std=.02)
        if self.enable_size_emb:
            trunc_normal_(self.size_emb.weight, std=.02)
        self.apply(self._init_weights)

        self._packed: Optional[weights.Packed] = None
        self._packed_key = None
        self._cfg_struct = weights.tsf_cfg_struct(config)
        self._ws = None
        self._train_pack = None
        self._train_pack_key = None

    def _init_weights(self, m):                                                # :207-214
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):                                                 # :216-221
        return {'pos_emb', 'cls_token', 'size_emb'} if self.enable_size_emb else {'pos_emb', 'cls_token'}

    # ------------------------------------------------------------------ packing cache
    def set_precision(self, precision: str):
        _lib.prec_id(precision)
        self.precision = precision
        self._packed = None
        return self

    def _load_from_state_dict(self, *a, **k):
        self._packed = None
        return super()._load_from_state_dict(*a, **k)

    def load_state_dict(self, state_dict, *a, **k):
        """Also accepts a checkpoint saved from ``nn.DataParallel(model)`` ('module.' prefix: train.py:461-464 saves
        them, predict.py:375-386 loads them into a wrapped model) directly into the bare module."""
        return super().load_state_dict(weights._strip(state_dict), *a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _get_packed(self, device) -> weights.Packed:
        key = (self.precision, str(device), tuple(p._version for p in self.parameters()))
        if self._packed is None or self._packed_key != key:
            self._packed = weights.pack_tsf(self.state_dict(), self.config, self.precision, device)
            self._packed_key = key
        return self._packed

    def _get_train_pack(self, device):
        from . import training
        key = (self.precision, str(device), tuple(p._version for p in self.parameters()),
               tuple(p.data_ptr() for p in self.parameters()))
        if self._train_pack is None or self._train_pack_key != key:
            self._train_pack = training.TrainPack(self, self.precision, device)
            self._train_pack_key = key
        return self._train_pack

    # ------------------------------------------------------------------ forward
    def forward(self, x, mask=None, identities_mask=None, size_embedding=None, positions=None):
        """x: (B,f,C,h,w) features (train.py:354).  mask (B,f) bool, identities_mask (B,f,f) bool,
        size_embedding (B,f) int (may arrive on CPU like in the reference, :245), positions (B,1+f*h*w) int64.
        Returns logits (B,num_classes) float32, plus [space_attn, time_attn] (each (B*heads,1,N) float32)
        when ``require_attention`` (:271-276)."""
        if x.dim() != 5:
            raise ValueError(f"expected (b,f,c,h,w) features, got {tuple(x.shape)}")
        b, f, c, h, w = x.shape
        n = h * w
        if f != self.num_frames:
            raise ValueError(f"got {f} frames but config num-frames is {self.num_frames} (the reference's mask "
                             f"repeat at size_invariant_timesformer.py:252 has the same requirement)")
        if c != self.channels or n != self.num_patches:
            raise ValueError(f"features have c={c}, h*w={n}; config says channels={self.channels}, "
                             f"num-patches={self.num_patches}")
        if mask is None or identities_mask is None:
            raise ValueError("mask and identities_mask are required (the reference crashes on None at :252-253)")
        if self.enable_size_emb and size_embedding is None:
            raise ValueError("size_embedding is required when enable-size-emb is on")
        if self.enable_pos_emb and positions is None:
            raise ValueError("positions is required when enable-pos-emb is on")
        dev = x.device
        _lib.require_device(dev)
        lib = _lib.load()
        T = _lib.torch_dtype(self.precision)
        prec = _lib.prec_id(self.precision)
        # token layout 'b (f h w) c' (:227).  The extractor shim already produces this memory order, so
        # this is a view; a true (b,f,c,h,w)-contiguous input costs one transposing copy (plumbing).
        tok = x.permute(0, 1, 3, 4, 2)
        if tok.dtype != T:
            tok = tok.to(dtype=T)
        if not tok.is_contiguous():
            # (Tensor.to(dtype, memory_format=contiguous_format) returns the SAME strided tensor when the dtype
            # already matches, so the copy has to be explicit)
            tok = tok.contiguous()
        mask_u8 = mask.to(device=dev, dtype=torch.uint8).contiguous()
        idm_u8 = identities_mask.to(device=dev, dtype=torch.uint8).contiguous()
        if mask_u8.shape != (b, f) or idm_u8.shape != (b, f, f):
            raise ValueError(f"mask {tuple(mask.shape)} / identities_mask {tuple(identities_mask.shape)} do not match "
                             f"(B,f)=({b},{f})")
        se = pos = None
        if self.enable_size_emb:
            se = size_embedding.to(device=dev, dtype=torch.int32).contiguous()
            if se.shape != (b, f):
                raise ValueError(f"size_embedding must be (B,f), got {tuple(se.shape)}")
        if self.enable_pos_emb:
            pos = positions.to(device=dev, dtype=torch.int64).contiguous()
            if pos.shape != (b, 1 + f * n):
                raise ValueError(f"positions must be (B,1+f*n), got {tuple(pos.shape)}")
        N = 1 + f * n
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training step (train.py:355, 376-378): forward that keeps activations + hand-written backward
            from . import training
            out = training.TsfTrainFunction.apply(self, tok.reshape(b, f * n, c), mask_u8, idm_u8, se, pos,
                                                  *self.parameters())
            if self.require_attention:
                return out[0], [out[1], out[2]]
            return out
        with torch.cuda.device(dev):
            pk = self._get_packed(dev)
            need = lib.mt_tsf_workspace_bytes(self._cfg_struct, b, prec)
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            logits = torch.empty((b, self.num_classes), dtype=torch.float32, device=dev)
            sa = ta = None
            if self.require_attention:
                sa = torch.empty((b * self.heads, 1, N), dtype=torch.float32, device=dev)
                ta = torch.empty((b * self.heads, 1, N), dtype=torch.float32, device=dev)
            rc = lib.mt_tsf_fwd(pk.struct, self._cfg_struct, tok.data_ptr(), mask_u8.data_ptr(), idm_u8.data_ptr(),
                                _lib.ptr(se), _lib.ptr(pos), logits.data_ptr(), _lib.ptr(sa), _lib.ptr(ta), b, prec,
                                self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr())
            _lib.check(rc, "mt_tsf_fwd")
        if self.require_attention:
            return logits, [sa, ta]
        return logits
