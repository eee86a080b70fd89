"""CPU ORACLE for the Xception extractor -- TEST INFRASTRUCTURE ONLY (same rules as oracle/mintime_oracle.py: only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import it; the product path never does).

A flat, functional fp32 restatement (plain torch CPU ops on a ``state_dict``) of ``Xception.features`` in eval mode, reference
``models/xception.py`` (paths relative to the reference repo); the arithmetic lives in PyTorch (torch==1.11.0,
requirements.txt:111): conv2d / batch_norm / max_pool2d / relu on the same values in the same order.

Pinned against outputs of the UNMODIFIED reference module (``oracle/make_golden_xception.py`` imports it from
/root/reference -> ``tests/golden/xception_*.npz``); ``tests/test_oracle.py`` checks oracle == fixtures.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# (in, out, reps, stride, start_with_relu, grow_first) of block1..block12   (xception.py:113-129)
BLOCKS = ([(64, 128, 2, 2, False, True), (128, 256, 2, 2, True, True), (256, 728, 2, 2, True, True)]
          + [(728, 728, 3, 1, True, True)] * 8 + [(728, 1024, 2, 2, True, False)])


def _bn(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """nn.BatchNorm2d in eval mode, default eps 1e-5 (xception.py:91: ``BN = nn.BatchNorm2d``)"""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def _sep(sd: Dict[str, Tensor], p: str, x: Tensor) -> Tensor:
    """SeparableConv2d.forward (xception.py:25-28): depthwise 3x3 pad 1, then 1x1, no biases"""
    x = F.conv2d(x, sd[p + ".conv1.weight"], None, 1, 1, 1, x.shape[1])
    return F.conv2d(x, sd[p + ".pointwise.weight"])


def block_forward(sd: Dict[str, Tensor], p: str, spec, inp: Tensor) -> Tensor:
    """Block.forward (xception.py:66-77) over the ``rep`` built at :40-64.  The first ReLU of ``rep`` is NOT in place
    (:60 replaces it), so the skip path sees the block's input as it came in."""
    cin, cout, reps, stride, start_with_relu, grow_first = spec
    n_sep = reps
    x = inp
    idx = 1 if start_with_relu else 0
    for j in range(n_sep):
        if j > 0 or start_with_relu:
            x = F.relu(x)
        x = _sep(sd, f"{p}rep.{idx}", x)
        x = _bn(sd, f"{p}rep.{idx + 1}", x)
        idx += 3
    if stride != 1:
        x = F.max_pool2d(x, 3, stride, 1)
    if cout != cin or stride != 1:
        skip = F.conv2d(inp, sd[p + "skip.weight"], None, stride)
        skip = _bn(sd, p + "skipbn", skip)
    else:
        skip = inp
    return x + skip


def xception_features(sd: Dict[str, Tensor], x: Tensor, stages: bool = False):
    """Xception.features (xception.py:146-184) == Xception.forward (:196-198).  x: (n,3,H,W) float32.
    Returns the (n,2048,h,w) output of bn4 (no ReLU after it), and the per-stage tensors when asked."""
    feats: List[Tuple[str, Tensor]] = []
    x = F.relu(_bn(sd, "bn1", F.conv2d(x, sd["conv1.weight"], None, 2, 0)))           # :148-150
    feats.append(("conv1", x))
    x = F.relu(_bn(sd, "bn2", F.conv2d(x, sd["conv2.weight"], None, 1, 0)))           # :152-154
    feats.append(("conv2", x))
    for i, spec in enumerate(BLOCKS):                                                 # :157-180
        x = block_forward(sd, f"block{i + 1}.", spec, x)
        feats.append((f"block{i + 1}", x))
    x = F.relu(_bn(sd, "bn3", _sep(sd, "conv3", x)))                                  # :182-184
    feats.append(("conv3", x))
    x = _bn(sd, "bn4", _sep(sd, "conv4", x))                                          # :186-187
    feats.append(("conv4", x))
    return (x, feats) if stages else x
