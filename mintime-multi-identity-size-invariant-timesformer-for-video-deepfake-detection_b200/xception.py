"""``Xception`` with the reference's module interface (models/xception.py:78-198 and the ``xception()`` factory :200-232),
running on the B200 library (csrc/xception.cu behind ``mt_xception_fwd``).

The alternative 2048-channel extractor of train.py:129-133 / predict.py:365-369 (``--extractor_model 1``; the shipped
``config/size_invariant_timesformer.yaml`` has ``channels: 2048`` for it).  The sub-modules below only hold the parameters
under the reference's ``state_dict`` names (``conv1.weight``, ``block3.rep.4.pointwise.weight``, ``block12.skipbn.running_var``
...); the arithmetic is in the library, eval mode only (BatchNorm with running statistics, folded at load).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, weights

# (in, out, reps, stride, start_with_relu, grow_first) of block1..block12   (xception.py:113-129)
XCEPTION_BLOCKS = ([(64, 128, 2, 2, False, True), (128, 256, 2, 2, True, True), (256, 728, 2, 2, True, True)]
                   + [(728, 728, 3, 1, True, True)] * 8 + [(728, 1024, 2, 2, True, False)])
XCEPTION_OUT = 2048


class _Slot(nn.Module):
    """parameter-free position of a ``rep`` Sequential (a ReLU or the MaxPool): keeps the reference's indices"""

    def forward(self, x):  # pragma: no cover - containers are never called
        raise RuntimeError("mintime_b200.Xception sub-modules are parameter containers")


class _Sep(nn.Module):
    """SeparableConv2d parameters (xception.py:17-23): depthwise 3x3 ``conv1`` + 1x1 ``pointwise``, no bias"""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cin, 3, 1, 1, groups=cin, bias=False)
        self.pointwise = nn.Conv2d(cin, cout, 1, bias=False)

    forward = _Slot.forward


def sep_channels(cin: int, cout: int, reps: int, grow_first: bool):
    """(cin, cout) of the separable convolutions of one block, in order (xception.py:43-58)"""
    if grow_first:
        return [(cin, cout)] + [(cout, cout)] * (reps - 1)
    return [(cin, cin)] * (reps - 1) + [(cin, cout)]


class _Block(nn.Module):
    """Block parameters (xception.py:29-65): ``rep`` = ([relu] sep bn)* [maxpool], optional ``skip`` / ``skipbn``"""

    def __init__(self, cin, cout, reps, stride, start_with_relu, grow_first):
        super().__init__()
        if cout != cin or stride != 1:
            self.skip = nn.Conv2d(cin, cout, 1, stride=stride, bias=False)
            self.skipbn = nn.BatchNorm2d(cout)
        else:
            self.skip = None
        rep = []
        for a, b in sep_channels(cin, cout, reps, grow_first):
            rep += [_Slot(), _Sep(a, b), nn.BatchNorm2d(b)]
        if not start_with_relu:
            rep = rep[1:]
        if stride != 1:
            rep.append(_Slot())
        self.rep = nn.Sequential(*rep)

    forward = _Slot.forward


class Xception(nn.Module):
    def __init__(self, in_channels=3, num_classes=1000, bn_group_size=1, bn_group=None, bn_sync_stats=True,
                 feature_visible=False, dropout=0, return_feature_idx=None, bypass_last_bn=False, precision: str = "bf16",
                 **kwargs):
        super().__init__()
        if in_channels != 3:
            raise ValueError("the MINTIME path feeds 3-channel face crops (in_channels=3)")
        _lib.prec_id(precision)
        self.precision = precision
        self.num_classes = num_classes
        self.conv1 = nn.Conv2d(3, 32, 3, 2, 0, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        self.conv2 = nn.Conv2d(32, 64, 3, bias=False)
        self.bn2 = nn.BatchNorm2d(64)
        for i, spec in enumerate(XCEPTION_BLOCKS):
            setattr(self, f"block{i + 1}", _Block(*spec))
        self.conv3 = _Sep(1024, 1536)
        self.bn3 = nn.BatchNorm2d(1536)
        self.conv4 = _Sep(1536, 2048)
        self.bn4 = nn.BatchNorm2d(XCEPTION_OUT)
        self.fc = nn.Linear(XCEPTION_OUT, num_classes)      # in the reference's state_dict; unused by forward (:196-198)
        for m in self.modules():                              # xception.py:144-152
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        self._packed: Optional[weights.Packed] = None
        self._packed_key = None
        self._ws = None

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
        return super().load_state_dict(weights._strip(state_dict), *a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _get_packed(self, device) -> weights.Packed:
        key = (self.precision, str(device), tuple(p._version for p in self.parameters()),
               tuple(b._version for b in self.buffers()))
        if self._packed is None or self._packed_key != key:
            self._packed = weights.pack_xception(self.state_dict(), self.precision, device)
            self._packed_key = key
        return self._packed

    # ------------------------------------------------------------------ forward
    def features(self, inputs: torch.Tensor):
        """xception.py:146-184 returns (x, [per-block features]); the per-block list is only used by the commented-out
        classifier path (:199-210) and is not materialised here."""
        return self.forward(inputs), []

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        """inputs: (n,3,H,W) float32 / uint8 -- ideally the permuted NHWC view the callers build (train.py:341), values as
        the caller feeds them (the module does not normalise).  Returns (n,2048,h,w) (NHWC memory), h = w = 7 for 224."""
        if inputs.dim() != 4 or inputs.shape[1] != 3 or inputs.shape[2] != inputs.shape[3]:
            raise ValueError(f"expected (n,3,H,H), got {tuple(inputs.shape)}")
        if self.training:
            raise _lib.MintimeError("mintime_b200.Xception: train mode (batch-statistic BatchNorm, backward) is not built; "
                                    "call .eval() (train.py --freeze_backbone) -- the EfficientNet-B0 extractor trains")
        _lib.require_device(inputs.device)
        lib = _lib.load()
        n, hw = inputs.shape[0], inputs.shape[2]
        o = lib.mt_xception_out_hw(hw)
        if o <= 0:
            raise ValueError("input too small for Xception")
        if inputs.dtype not in (torch.float32, torch.uint8):
            inputs = inputs.float()
        x = inputs.permute(0, 2, 3, 1)
        if not x.is_contiguous():
            x = x.contiguous()          # plumbing: caller gave true NCHW memory
        dev = x.device
        pk = self._get_packed(dev)
        prec = _lib.prec_id(self.precision)
        need = lib.mt_xception_workspace_bytes(n, hw, prec)
        if self._ws is None or self._ws.device != dev or self._ws.numel() < need:
            self._ws = torch.empty((need,), dtype=torch.uint8, device=dev)
        feats = torch.empty((n * o * o, XCEPTION_OUT), dtype=_lib.torch_dtype(self.precision), device=dev)
        with torch.cuda.device(dev):
            rc = lib.mt_xception_fwd(pk.struct, x.data_ptr(), 1 if x.dtype == torch.uint8 else 0, feats.data_ptr(), n, hw, prec,
                                     self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr())
        _lib.check(rc, "mt_xception_fwd")
        return feats.view(n, o, o, XCEPTION_OUT).permute(0, 3, 1, 2)


def xception(pretrain_path=None, **kwargs):
    """models/xception.py:200-232: build the model and copy every checkpoint entry whose (``module.``-stripped) name and
    shape match; entries that do not fit are reported and skipped, like the reference does."""
    model = Xception(**kwargs)
    if pretrain_path is not None:
        state_dict = torch.load(pretrain_path, map_location="cpu")
        own = model.state_dict()
        for name, param in state_dict.items():
            name = name.replace("module.", "")
            if name not in own:
                continue
            if isinstance(param, torch.nn.Parameter):
                param = param.data
            if own[name].shape != param.shape:
                print(f"While copying the parameter named {name}, whose dimensions in the model are {tuple(own[name].shape)} "
                      f"and whose dimensions in the checkpoint are {tuple(param.shape)}.")
                continue
            own[name].copy_(param)
        model._packed = None
        print("Features Extractor checkpoint loaded.")
    return model
