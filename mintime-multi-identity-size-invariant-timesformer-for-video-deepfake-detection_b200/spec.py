"""Static description of the hot path: EfficientNet-B0 stage table and TimeSformer config.

The stage table restates what the reference builds from its block strings
(models/efficientnet/efficientnet_pytorch/utils.py:502-510, decoded by utils.py:361-454 and
expanded by model.py:171-191) for width=depth=1.0, image_size=224.  TF-"SAME" padding follows
utils.py:254-269: pad = max((ceil(i/s)-1)*s + k - i, 0), split floor/ceil (left/top gets the
smaller half).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List

BN_EPS = 1e-3          # utils.py:521 batch_norm_epsilon
BN_MOMENTUM = 0.01     # 1 - 0.99, model.py:51
IMAGE_SIZE = 224
STEM_OUT = 32
HEAD_OUT = 1280
DROP_CONNECT_RATE = 0.2


@dataclass(frozen=True)
class MBConvSpec:
    index: int
    kernel: int
    stride: int
    expand: int
    cin: int
    cout: int
    hw_in: int          # spatial side at block input
    @property
    def cexp(self) -> int:
        return self.cin * self.expand
    @property
    def hw_out(self) -> int:
        return math.ceil(self.hw_in / self.stride)
    @property
    def se_squeeze(self) -> int:
        # model.py:78  max(1, int(input_filters * se_ratio)), se_ratio = 0.25
        return max(1, int(self.cin * 0.25))
    @property
    def pad_lo(self) -> int:
        """left/top zero padding of the depthwise conv (utils.py:264-269)."""
        total = max((self.hw_out - 1) * self.stride + self.kernel - self.hw_in, 0)
        return total // 2
    @property
    def has_skip(self) -> bool:
        # model.py:123  id_skip and stride == 1 and cin == cout.  (For the first block of a stage
        # the reference compares a list [1] with 1 -> False; those blocks change channels anyway.)
        return self.stride == 1 and self.cin == self.cout


# (repeats, kernel, stride, expand, cin, cout)  -- utils.py:502-510
_B0_STAGES = [
    (1, 3, 1, 1, 32, 16),
    (2, 3, 2, 6, 16, 24),
    (2, 5, 2, 6, 24, 40),
    (3, 3, 2, 6, 40, 80),
    (3, 5, 1, 6, 80, 112),
    (4, 5, 2, 6, 112, 192),
    (1, 3, 1, 6, 192, 320),
]


def b0_blocks(image_size: int = IMAGE_SIZE) -> List[MBConvSpec]:
    hw = math.ceil(image_size / 2)       # after the stride-2 stem
    out: List[MBConvSpec] = []
    for (r, k, s, e, ci, co) in _B0_STAGES:
        for j in range(r):
            spec = MBConvSpec(len(out), k, s if j == 0 else 1, e, ci if j == 0 else co, co, hw)
            out.append(spec)
            hw = spec.hw_out
    return out


B0_BLOCKS = b0_blocks()
assert len(B0_BLOCKS) == 16 and B0_BLOCKS[-1].hw_out == 7


def default_tsf_config(num_frames: int = 16, channels: int = 1280) -> dict:
    """config/size_invariant_timesformer.yaml with the EfficientNet channel count (yaml:25 comment)."""
    return {
        "training": {"lr": 0.01, "weight-decay": 0.0001, "bs": 8, "val_bs": 8, "optimizer": "SGD",
                     "scheduler": "cosinelr", "gamma": 0.1, "step-size": 5, "augmentation": "max"},
        "test": {"bs": 1},
        "model": {
            "image-size": 224, "patch-size": 1, "num-classes": 1, "num-patches": 49,
            "num-frames": num_frames, "max-identities": 2, "dim": 512, "depth": 9, "dim-head": 64,
            "channels": channels, "heads": 8, "attn-dropout": 0.0, "ff-dropout": 0.0,
            "shift-tokens": False, "enable-size-emb": True, "enable-pos-emb": True,
            "enable-identity-attention": True,
        },
    }
