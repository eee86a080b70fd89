"""Shared helpers for the parity tests (rebuild the seeded inputs behind each golden fixture)."""
import os

import numpy as np
import torch

import mintime_b200  # noqa: F401
from mintime_b200 import synth
from mintime_b200.spec import default_tsf_config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# must match oracle/make_golden.py:CASES
CASES = {
    "cfg1_b1_f8_id1": (1, 8, [1], False),
    "b2_f16_id2": (2, 16, [2], True),
    "b4_f16_mixed": (4, 16, [1, 2, 3, 4], True),
    "b2_f8_id2": (2, 8, [2, 1], True),
}


def sample_index(numel: int, k: int = 2048) -> torch.Tensor:
    if numel <= k:
        return torch.arange(numel)
    return (torch.arange(k, dtype=torch.float64) * (numel - 1) / (k - 1)).round().long()


def sample(t: torch.Tensor, k: int = 2048) -> np.ndarray:
    flat = t.detach().float().cpu().reshape(-1)
    return flat[sample_index(flat.numel(), k)].numpy()


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


_CACHE = {}


def case_inputs(name):
    """(cfg, effnet_sd, tsf_sd, meta, frames) for a golden case, rebuilt from the seeds.  'cond_<case>' is the same
    case with the conditioned extractor weights (the bf16 end-to-end fixtures, oracle/make_golden.py --conditioned)."""
    if name in _CACHE:
        return _CACHE[name]
    cond = name.startswith("cond_")
    B, f, ids, pad = CASES[name[5:] if cond else name]
    cfg = default_tsf_config(num_frames=f, channels=1280)
    ekey = "esd_cond" if cond else "esd"
    if ekey not in _CACHE:
        _CACHE[ekey] = synth.make_effnet_state_dict(1234, conditioned=cond)
    key = ("tsd", f)
    if key not in _CACHE:
        _CACHE[key] = synth.make_tsf_state_dict(cfg, 4321)
    meta = synth.make_batch_meta(B, f, ids, seed=1234, pad_tail=pad)
    frames = synth.make_frames(B, f, seed=1234, mask=meta["mask"])
    _CACHE[name] = (cfg, _CACHE[ekey], _CACHE[key], meta, frames)
    return _CACHE[name]


BENCH_CLIPS = [0, 9, 18, 31]      # oracle/make_golden.py: clips of the bench batch held in cond_bench_b32_clips.npz


def bench_batch_inputs():
    """(cfg, conditioned effnet_sd, tsf_sd, meta, uint8 frames) of bench.py's rank-0 batch (B = 32, f = 16, 1 identity)"""
    B, f = 32, 16
    cfg = default_tsf_config(num_frames=f, channels=1280)
    meta = synth.make_batch_meta(B, f, [1], seed=1234)
    frames = synth.make_frames(B, f, seed=1234, mask=meta["mask"], dtype=torch.uint8)
    return cfg, synth.make_effnet_state_dict(1234, conditioned=True), synth.make_tsf_state_dict(cfg, 4321), meta, frames


def rel_err(a, b):
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


# ---------------------------------------------------------------------------------------------------
# training-step cases (oracle/make_golden_grads.py): name -> (B, f, identities, depth)
# ---------------------------------------------------------------------------------------------------
GRAD_CASES = {
    "b2_f8_id2": (2, 8, [2, 1], 9),
    "b3_f16_mixed_d2": (3, 16, [3, 1, 2], 2),
}


def grad_case_inputs(name):
    """(cfg, tsf_sd, meta, feats (B,f,1280,7,7) bf16-representable float32, labels (B,1), pos_weight)"""
    B, f, ids, depth = GRAD_CASES[name]
    cfg = default_tsf_config(num_frames=f, channels=1280)
    cfg["model"]["depth"] = depth
    tsd = synth.make_tsf_state_dict(cfg, 777)
    meta = synth.make_batch_meta(B, f, ids, seed=f + 1, pad_tail=True)
    g = torch.Generator().manual_seed(1000 + f)
    feats = torch.nn.functional.silu(torch.randn((B, f, 1280, 7, 7), generator=g)) * 20.0
    feats = feats.bfloat16().float()
    labels = torch.tensor([[float(i % 2 == 0)] for i in range(B)])
    return cfg, tsd, meta, feats, labels, 0.8169


# ---------------------------------------------------------------------------------------------------
# extractor in train mode (oracle/make_golden_extractor_train.py)
# ---------------------------------------------------------------------------------------------------
EXTRACTOR_TRAIN_KEYS = ["_conv_stem.weight", "_bn0.weight", "_blocks.0._depthwise_conv.weight", "_blocks.1._expand_conv.weight",
                        "_blocks.2._se_reduce.weight", "_blocks.5._bn1.bias", "_blocks.10._project_conv.weight",
                        "_blocks.15._se_expand.bias", "_conv_head.weight", "_bn1.weight"]


def extractor_train_inputs():
    """(effnet state_dict, 4 faces (4,3,224,224) raw 0..255, probe (4,1280,7,7)) for the train-mode extractor fixture"""
    esd = synth.make_effnet_state_dict(1234)
    meta = synth.make_batch_meta(1, 4, [1], seed=11, pad_tail=False)
    frames = synth.make_frames(1, 4, seed=11, mask=meta["mask"])                 # (1,4,224,224,3)
    x = frames.view(4, 224, 224, 3).permute(0, 3, 1, 2).contiguous()
    probe = torch.randn((4, 1280, 7, 7), generator=torch.Generator().manual_seed(3))
    return esd, x, probe


# ---------------------------------------------------------------------------------------------------
# Xception extractor (oracle/make_golden_xception.py)
# ---------------------------------------------------------------------------------------------------
XCEPTION_STAGES = ["conv2"] + [f"block{i}" for i in range(1, 13)]


def xception_inputs():
    """(state_dict, 2 faces (2,3,224,224) raw 0..255) behind tests/golden/xception_b2.npz"""
    sd = synth.make_xception_state_dict(2468)
    meta = synth.make_batch_meta(1, 2, [1], seed=21, pad_tail=False)
    frames = synth.make_frames(1, 2, seed=21, mask=meta["mask"])
    return sd, frames.view(2, 224, 224, 3).permute(0, 3, 1, 2).contiguous()


def xception_tsf_inputs():
    """(cfg, xception sd, tsf sd, meta, frames (1,8,224,224,3)) behind tests/golden/xception_tsf_b1_f8.npz"""
    B, f = 1, 8
    cfg = default_tsf_config(num_frames=f, channels=2048)
    meta = synth.make_batch_meta(B, f, [2], seed=33, pad_tail=True)
    frames = synth.make_frames(B, f, seed=33, mask=meta["mask"])
    return cfg, synth.make_xception_state_dict(2468), synth.make_tsf_state_dict(cfg, 4321), meta, frames
