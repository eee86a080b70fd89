"""Seeded synthetic weights and clips for parity tests and the benchmark.

No dataset or checkpoint is reachable offline, so every test/bench uses these.  All randomness comes
from numpy's PCG64 (stable across machines), never from torch's generator.

* Weights carry the reference's exact ``state_dict`` key names and shapes so
  the same dict loads into the reference modules (strict) and into this package's shims.
* Clips follow what ``deepfakes_dataset.py`` hands the model (reference deepfakes_dataset.py:259-341):
  uint8-valued frames, per-face size-embedding bucket, padded-slot mask, block-diagonal
  identities mask and the temporally coherent positions.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .spec import B0_BLOCKS, HEAD_OUT, STEM_OUT


# ----------------------------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------------------------
def _bn(rng: np.random.Generator, c: int, prefix: str, out: Dict[str, torch.Tensor]) -> None:
    out[prefix + ".weight"] = torch.from_numpy(rng.uniform(0.5, 1.5, c).astype(np.float32))
    out[prefix + ".bias"] = torch.from_numpy((rng.standard_normal(c) * 0.1).astype(np.float32))
    out[prefix + ".running_mean"] = torch.from_numpy((rng.standard_normal(c) * 0.1).astype(np.float32))
    out[prefix + ".running_var"] = torch.from_numpy(rng.uniform(0.5, 1.5, c).astype(np.float32))
    out[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def _conv(rng: np.random.Generator, cout: int, cin_per_group: int, k: int, scale: float = 1.0) -> torch.Tensor:
    fan_in = cin_per_group * k * k
    w = rng.standard_normal((cout, cin_per_group, k, k)) * np.sqrt(2.0 / fan_in) * scale
    return torch.from_numpy(w.astype(np.float32))


def make_effnet_state_dict(seed: int = 1234, conditioned: bool = False) -> Dict[str, torch.Tensor]:
    """EfficientNet-B0 ``state_dict`` (360 keys, names as in reference model.py:49-87,155-211).

    The default ``from_name`` init collapses activations to ~1e-9, which would let any
    kernel pass; this recipe keeps every stage O(1..50): He-normal convs, perturbed BN statistics,
    stem scaled by 1/64 because inputs are raw 0..255 pixels (no normalisation on the path).

    ``conditioned=True`` is the recipe of the bf16 end-to-end fixtures (tests/golden/cond_*.npz): the He-init net above
    is chaotic -- its residual blocks amplify rounding noise ~30x, the REFERENCE itself moves 0.29-0.42 rel-L2 on the
    features under bf16 autocast -- so, like a trained network, the residual branches are damped (``_bn2`` gamma of
    the skip blocks ~ U(0.1, 0.3)), the depthwise filters are scaled by 0.7 and the head conv by 8 to keep the features
    O(1).  Reference drift under bf16 autocast with these weights: 6e-3 on the features (oracle/measure_ref_bf16_drift.py).
    Same keys, same shapes, same seeded base values.
    """
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    sd["_conv_stem.weight"] = _conv(rng, STEM_OUT, 3, 3, scale=1.0 / 64.0)
    _bn(rng, STEM_OUT, "_bn0", sd)
    for b in B0_BLOCKS:
        p = f"_blocks.{b.index}."
        if b.expand != 1:
            sd[p + "_expand_conv.weight"] = _conv(rng, b.cexp, b.cin, 1)
            _bn(rng, b.cexp, p + "_bn0", sd)
        sd[p + "_depthwise_conv.weight"] = _conv(rng, b.cexp, 1, b.kernel)
        _bn(rng, b.cexp, p + "_bn1", sd)
        sq = b.se_squeeze
        sd[p + "_se_reduce.weight"] = _conv(rng, sq, b.cexp, 1)
        sd[p + "_se_reduce.bias"] = torch.from_numpy((rng.standard_normal(sq) * 0.1).astype(np.float32))
        sd[p + "_se_expand.weight"] = _conv(rng, b.cexp, sq, 1)
        sd[p + "_se_expand.bias"] = torch.from_numpy((rng.standard_normal(b.cexp) * 0.1).astype(np.float32))
        sd[p + "_project_conv.weight"] = _conv(rng, b.cout, b.cexp, 1)
        _bn(rng, b.cout, p + "_bn2", sd)
    sd["_conv_head.weight"] = _conv(rng, HEAD_OUT, B0_BLOCKS[-1].cout, 1)
    _bn(rng, HEAD_OUT, "_bn1", sd)
    sd["_fc.weight"] = torch.from_numpy((rng.standard_normal((1000, HEAD_OUT)) * 0.01).astype(np.float32))
    sd["_fc.bias"] = torch.zeros(1000)
    if conditioned:
        rng2 = np.random.default_rng(seed + 7)
        for b in B0_BLOCKS:
            p = f"_blocks.{b.index}."
            if b.has_skip:
                sd[p + "_bn2.weight"] = torch.from_numpy(rng2.uniform(0.1, 0.3, b.cout).astype(np.float32))
            sd[p + "_depthwise_conv.weight"] = sd[p + "_depthwise_conv.weight"] * 0.7
        sd["_conv_head.weight"] = sd["_conv_head.weight"] * 8.0
    return sd


def make_xception_state_dict(seed: int = 2468, num_classes: int = 1) -> Dict[str, torch.Tensor]:
    """Xception ``state_dict`` (names as in reference models/xception.py:93-137): He-normal convolutions, perturbed BatchNorm
    statistics, conv1 scaled by 1/64 (raw 0..255 pixels), and -- like a trained residual network -- the last BatchNorm of every
    block's main path damped (gamma ~ U(0.05, 0.15)) so that the twelve residual sums keep the activations O(1..20); bn4's
    gamma is scaled by 0.1 so that the 2048 output features are O(1..10)."""
    from .xception import XCEPTION_BLOCKS, sep_channels
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    sd["conv1.weight"] = _conv(rng, 32, 3, 3, scale=1.0 / 64.0)
    _bn(rng, 32, "bn1", sd)
    sd["conv2.weight"] = _conv(rng, 64, 32, 3)
    _bn(rng, 64, "bn2", sd)

    def sep(prefix, bn_prefix, cin, cout, damp=False):
        sd[prefix + ".conv1.weight"] = _conv(rng, cin, 1, 3)
        sd[prefix + ".pointwise.weight"] = _conv(rng, cout, cin, 1)
        _bn(rng, cout, bn_prefix, sd)
        if damp:
            sd[bn_prefix + ".weight"] = torch.from_numpy(rng.uniform(0.05, 0.15, cout).astype(np.float32))

    for bi, (cin, cout, reps, stride, relu0, grow_first) in enumerate(XCEPTION_BLOCKS):
        p = f"block{bi + 1}."
        if cout != cin or stride != 1:
            sd[p + "skip.weight"] = _conv(rng, cout, cin, 1)
            _bn(rng, cout, p + "skipbn", sd)
        idx = 1 if relu0 else 0
        chans = sep_channels(cin, cout, reps, grow_first)
        for j, (a, b) in enumerate(chans):
            sep(f"{p}rep.{idx}", f"{p}rep.{idx + 1}", a, b, damp=(j == len(chans) - 1))
            idx += 3
    sep("conv3", "bn3", 1024, 1536)
    sep("conv4", "bn4", 1536, 2048)
    sd["bn4.weight"] = sd["bn4.weight"] * 0.1
    sd["fc.weight"] = torch.from_numpy((rng.standard_normal((num_classes, 2048)) * 0.01).astype(np.float32))
    sd["fc.bias"] = torch.zeros(num_classes)
    return sd


def _trunc_normal(rng: np.random.Generator, shape, std: float = 0.02) -> torch.Tensor:
    # reference size_invariant_timesformer.py:200-214 uses trunc_normal_(std=.02) (cut at +-2, i.e.
    # 100 sigma: effectively a plain normal).
    x = rng.standard_normal(shape) * std
    return torch.from_numpy(np.clip(x, -2.0, 2.0).astype(np.float32))


def make_tsf_state_dict(config: dict, seed: int = 4321, perturb: bool = True) -> Dict[str, torch.Tensor]:
    """SizeInvariantTimeSformer ``state_dict`` (names per reference size_invariant_timesformer.py:173-198).

    ``perturb`` additionally randomises biases / LayerNorm affine (the reference initialises them to
    0 / 1, which would hide a kernel that forgets them).
    """
    m = config["model"]
    dim, depth, heads, dh = m["dim"], m["depth"], m["heads"], m["dim-head"]
    C, f = m["channels"], m["num-frames"]
    inner = heads * dh
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}

    def bias(n):
        return torch.from_numpy((rng.standard_normal(n) * 0.02).astype(np.float32)) if perturb else torch.zeros(n)

    def ln(prefix):
        if perturb:
            sd[prefix + ".weight"] = torch.from_numpy(rng.uniform(0.8, 1.2, dim).astype(np.float32))
            sd[prefix + ".bias"] = torch.from_numpy((rng.standard_normal(dim) * 0.05).astype(np.float32))
        else:
            sd[prefix + ".weight"] = torch.ones(dim)
            sd[prefix + ".bias"] = torch.zeros(dim)

    sd["cls_token"] = _trunc_normal(rng, (1, dim))
    sd["to_patch_embedding.weight"] = _trunc_normal(rng, (dim, C))
    sd["to_patch_embedding.bias"] = bias(dim)
    sd["pos_emb.weight"] = _trunc_normal(rng, (f * C + 1, dim))
    if m["enable-size-emb"]:
        sd["size_emb.weight"] = _trunc_normal(rng, (f * C + 1, dim))
    for l in range(depth):
        for j in (0, 1):
            p = f"layers.{l}.{j}."
            ln(p + "norm")
            sd[p + "fn.to_qkv.weight"] = _trunc_normal(rng, (3 * inner, dim))
            sd[p + "fn.to_out.0.weight"] = _trunc_normal(rng, (dim, inner))
            sd[p + "fn.to_out.0.bias"] = bias(dim)
        p = f"layers.{l}.2."
        ln(p + "norm")
        sd[p + "fn.net.0.weight"] = _trunc_normal(rng, (dim * 8, dim))
        sd[p + "fn.net.0.bias"] = bias(dim * 8)
        sd[p + "fn.net.3.weight"] = _trunc_normal(rng, (dim, dim * 4))
        sd[p + "fn.net.3.bias"] = bias(dim)
    ln("to_out.0")
    sd["to_out.1.weight"] = _trunc_normal(rng, (m["num-classes"], dim))
    sd["to_out.1.bias"] = bias(m["num-classes"])
    return sd


# ----------------------------------------------------------------------------------------------
# clips
# ----------------------------------------------------------------------------------------------
def identity_slots(num_frames: int, n_identities: int) -> List[int]:
    """Slots per identity for a video with enough faces for every identity.

    Restates deepfakes_dataset.py:50-53 (max_faces_per_identity) and :171-186 (unused slots are
    topped up into the last identity).
    """
    f = num_frames
    table = {1: [f], 2: [f // 2, f // 2], 3: [f // 3, f // 3, f // 4], 4: [f // 3, f // 3, f // 8, f // 8]}
    slots = list(table[n_identities])
    slots[-1] += f - sum(slots)
    return slots


def make_clip_meta(num_frames: int, n_identities: int, rng: np.random.Generator, pad_tail: bool = True,
                   num_patches: int = 49) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """(size_emb[f] i32, mask[f] bool, identities_mask[f,f] bool, positions[1+f*n] i64) for one video.

    A synthetic INPUT generator shaped like predict.py generate_masks (:254-352): size bucket 1..20, 0 on padded
    slots; mask 0 on padded slots; block-diagonal identities mask; positions = rank of the source frame among the
    video's distinct frames.  (Padded slots repeat the identity's own last frame here; the executed reference repeats
    the clip-wide maximum so far -- the builder that reproduces the reference exactly is utils.build_clip_meta, pinned
    by tests/golden/clip_meta_ref.json.  The golden model fixtures were generated from THIS generator's tensors.)
    """
    f = num_frames
    slots = identity_slots(f, n_identities)
    size_emb = np.zeros(f, np.int32)
    mask = np.ones(f, bool)
    idm = np.zeros((f, f), bool)
    frames = np.zeros(f, np.int64)
    start = 0
    for ns in slots:
        npad = int(rng.integers(0, 3)) if (pad_tail and n_identities > 1 and ns > 2) else 0
        nreal = ns - npad
        idm[start:start + ns, start:start + ns] = True
        size_emb[start:start + nreal] = rng.integers(1, 21, nreal)
        mask[start + nreal:start + ns] = False
        # every identity is seen in (a prefix of) the same ordered set of source frames
        src = np.arange(nreal) * 3 + 7
        frames[start:start + nreal] = src
        frames[start + nreal:start + ns] = src.max()
        start += ns
    rank = {v: i + 1 for i, v in enumerate(sorted(set(frames.tolist())))}
    pos = [0]
    for fr in frames:
        p = rank[int(fr)]
        pos.extend(range((p - 1) * num_patches + 1, p * num_patches + 1))
    return size_emb, mask, idm, np.asarray(pos, np.int64)


def make_batch_meta(batch: int, num_frames: int, identities: Sequence[int], seed: int = 1234,
                    pad_tail: bool = True) -> Dict[str, torch.Tensor]:
    """Batch of clip metadata; ``identities`` is cycled over the batch (config 5: 1,2,3,4)."""
    rng = np.random.default_rng(seed)
    se, mk, im, ps = [], [], [], []
    for b in range(batch):
        a, m, i, p = make_clip_meta(num_frames, identities[b % len(identities)], rng, pad_tail)
        se.append(a); mk.append(m); im.append(i); ps.append(p)
    return {
        "size_embedding": torch.from_numpy(np.stack(se)),            # (B,f) int32
        "mask": torch.from_numpy(np.stack(mk)),                      # (B,f) bool
        "identities_mask": torch.from_numpy(np.stack(im)),           # (B,f,f) bool
        "positions": torch.from_numpy(np.stack(ps)),                 # (B,1+f*49) int64
    }


def make_frames(batch: int, num_frames: int, seed: int = 1234, mask: Optional[torch.Tensor] = None,
                image_size: int = 224, dtype=torch.float32) -> torch.Tensor:
    """uint8-valued frames, NHWC ``(B,f,H,W,3)`` as ``DeepFakesDataset.__getitem__`` returns them
    (deepfakes_dataset.py:339: ``torch.tensor(sequence).float()``, raw 0..255 BGR, no mean/std).
    Padded slots are zero images (deepfakes_dataset.py:276)."""
    rng = np.random.default_rng(seed + 99)
    x = rng.integers(0, 256, (batch, num_frames, image_size, image_size, 3), dtype=np.uint8)
    t = torch.from_numpy(x)
    if mask is not None:
        t = t * mask.view(batch, num_frames, 1, 1, 1).to(torch.uint8)
    return t.to(dtype)
