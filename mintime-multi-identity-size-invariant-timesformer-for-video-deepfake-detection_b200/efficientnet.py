"""Drop-in for the reference's ``EfficientNet`` feature extractor on the MINTIME path.

Mirrors models/efficientnet/efficientnet_pytorch/model.py: same class name, ``from_name`` /
``from_pretrained`` / ``load_matching_state_dict`` (model.py:345-411), same parameter names and
shapes (360 ``state_dict`` keys, ``_blocks.{i}.``-style names that train.py:159-167 parses), and
``forward(inputs)`` returning the 1280x7x7 feature map (model.py:267-288 -- MINTIME's forward stops
before pooling / ``_fc``).  The modules below only HOLD parameters; the arithmetic runs in
libmintime_b200.so (mt_effnet_b0_fwd).  There is no PyTorch fallback.

Differences a caller can see (documented in INTEGRATION.md):
  * output is NHWC memory viewed as (n,1280,7,7) (what ``rearrange('b f c h w -> b (f h w) c')`` wants),
    dtype float32 for precision='fp32', bfloat16 for precision='bf16' (default);
  * train mode (``.train()``, train.py:153-170) runs batch-statistics BatchNorm + drop-connect and is differentiable
    (efficientnet_train.py); it computes in float32 whatever ``precision`` says and returns float32 features.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

from . import _lib, weights
from .spec import B0_BLOCKS, BN_EPS, BN_MOMENTUM, HEAD_OUT, STEM_OUT

VALID_MODELS = ("efficientnet-b0",)


class MBConvBlock(nn.Module):
    """Parameter container with the reference's attribute names (model.py:49-87)."""

    def __init__(self, spec):
        super().__init__()
        self.spec = spec
        c, k = spec.cexp, spec.kernel
        if spec.expand != 1:
            self._expand_conv = nn.Conv2d(spec.cin, c, 1, bias=False)
            self._bn0 = nn.BatchNorm2d(c, momentum=BN_MOMENTUM, eps=BN_EPS)
        self._depthwise_conv = nn.Conv2d(c, c, k, stride=spec.stride, groups=c, bias=False)
        self._bn1 = nn.BatchNorm2d(c, momentum=BN_MOMENTUM, eps=BN_EPS)
        self._se_reduce = nn.Conv2d(c, spec.se_squeeze, 1)
        self._se_expand = nn.Conv2d(spec.se_squeeze, c, 1)
        self._project_conv = nn.Conv2d(c, spec.cout, 1, bias=False)
        self._bn2 = nn.BatchNorm2d(spec.cout, momentum=BN_MOMENTUM, eps=BN_EPS)

    def forward(self, *a, **k):
        raise RuntimeError("MBConvBlock is a parameter container; call EfficientNet.forward")


class EfficientNet(nn.Module):
    def __init__(self, model_name: str = "efficientnet-b0", precision: str = "bf16"):
        super().__init__()
        self._check_model_name_is_valid(model_name)
        self.precision = precision
        self._conv_stem = nn.Conv2d(3, STEM_OUT, 3, stride=2, bias=False)
        self._bn0 = nn.BatchNorm2d(STEM_OUT, momentum=BN_MOMENTUM, eps=BN_EPS)
        self._blocks = nn.ModuleList([MBConvBlock(s) for s in B0_BLOCKS])
        self._conv_head = nn.Conv2d(B0_BLOCKS[-1].cout, HEAD_OUT, 1, bias=False)
        self._bn1 = nn.BatchNorm2d(HEAD_OUT, momentum=BN_MOMENTUM, eps=BN_EPS)
        self._fc = nn.Linear(HEAD_OUT, 1000)          # present in reference checkpoints; unused by forward
        self.drop_connect_rate = 0.2                  # GlobalParams.drop_connect_rate of efficientnet-b0 (utils.py:523)
        self._drop_connect_rand = None                # test hook: callable(n) -> n uniform draws (default: torch.rand on the device)
        self._packed: Optional[weights.Packed] = None
        self._packed_key = None
        self._ws = None

    # ------------------------------------------------------------------ reference constructors
    @classmethod
    def _check_model_name_is_valid(cls, model_name):
        if model_name not in VALID_MODELS:
            raise ValueError("model_name should be one of: " + ", ".join(VALID_MODELS))

    @classmethod
    def from_name(cls, model_name, in_channels=3, **override_params):
        if in_channels != 3:
            raise ValueError("the MINTIME path feeds 3-channel face crops (in_channels=3)")
        precision = override_params.pop("precision", "bf16")
        drop = override_params.pop("drop_connect_rate", None)
        if override_params.get("image_size", 224) != 224:
            raise ValueError("the B200 extractor is built for 224x224 crops (config image-size: 224)")
        model = cls(model_name, precision=precision)
        if drop is not None:
            model.drop_connect_rate = float(drop)
        return model

    @classmethod
    def from_pretrained(cls, model_name, weights_path=None, advprop=False, in_channels=3, num_classes=1000,
                        **override_params):
        if weights_path is None:
            raise RuntimeError("from_pretrained without weights_path downloads ImageNet weights in the reference "
                               "(utils.py:556-602); offline, pass weights_path=<.pth state_dict>")
        model = cls.from_name(model_name, in_channels=in_channels, **override_params)
        sd = torch.load(weights_path, map_location="cpu")
        model.load_state_dict({k: v for k, v in sd.items() if num_classes == 1000 or not k.startswith("_fc")},
                              strict=False)
        return model

    def load_matching_state_dict(self, state_dict):
        """model.py:368-378: strip an 'efficient_net.' prefix, skip unknown keys, copy the rest."""
        own = self.state_dict()
        for name, param in state_dict.items():
            if "efficient_net" in name:
                name = name.split("efficient_net.")[1]
            if name not in own:
                continue
            if isinstance(param, torch.nn.parameter.Parameter):
                param = param.data
            own[name].copy_(param)
        self._packed = None

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
        """Also accepts 'module.'-prefixed (``nn.DataParallel``) checkpoints (predict.py:375-388) on the bare module."""
        return super().load_state_dict(weights._strip(state_dict), *a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _get_packed(self, device) -> weights.Packed:
        key = (self.precision, str(device), tuple(p._version for p in self.parameters()))
        if self._packed is None or self._packed_key != key:
            self._packed = weights.pack_effnet(self.state_dict(), self.precision, device)
            self._packed_key = key
        return self._packed

    # ------------------------------------------------------------------ forward
    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        """inputs: (n,3,224,224) float32 / uint8 -- ideally the permuted NHWC view the callers build
        (train.py:341 ``rearrange(videos, 'b f h w c -> (b f) c h w')``), raw 0..255.
        Returns (n,1280,7,7) (NHWC memory)."""
        if inputs.dim() != 4 or inputs.shape[1:] != (3, 224, 224):
            raise ValueError(f"expected (n,3,224,224), got {tuple(inputs.shape)}")
        _lib.require_device(inputs.device)
        lib = _lib.load()
        n = inputs.shape[0]
        if inputs.dtype not in (torch.float32, torch.uint8):
            inputs = inputs.float()
        x = inputs.permute(0, 2, 3, 1)
        if not x.is_contiguous():
            x = x.contiguous()          # plumbing: caller gave true NCHW memory
        if self.training:
            # train.py:153-170: batch-statistics BatchNorm (+ running-stat update), drop-connect, gradients for the
            # parameters that require them -- fp32 kernels of csrc/effnet_train.cu, returns float32 features
            from . import efficientnet_train
            feats = efficientnet_train.forward_train(self, x.float() if x.dtype != torch.float32 else x,
                                                     self._drop_connect_rand)
            return feats.permute(0, 3, 1, 2)
        T = _lib.torch_dtype(self.precision)
        prec = _lib.prec_id(self.precision)
        with torch.cuda.device(inputs.device):
            pk = self._get_packed(inputs.device)
            need = lib.mt_effnet_b0_workspace_bytes(n, prec)
            if self._ws is None or self._ws.numel() < need or self._ws.device != inputs.device:
                self._ws = torch.empty(need, dtype=torch.uint8, device=inputs.device)
            feats = torch.empty((n, 7, 7, HEAD_OUT), dtype=T, device=inputs.device)
            rc = lib.mt_effnet_b0_fwd(pk.struct, x.data_ptr(), _lib.IN_U8 if x.dtype == torch.uint8 else _lib.IN_F32,
                                      feats.data_ptr(), n, prec, self._ws.data_ptr(), self._ws.numel(),
                                      _lib.stream_ptr())
            _lib.check(rc, "mt_effnet_b0_fwd")
        return feats.permute(0, 3, 1, 2)
