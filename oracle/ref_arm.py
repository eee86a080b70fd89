"""The UNMODIFIED reference modules as the CPU arm of bench.py (`--impl reference`, `cpu_baseline.kind: "reference"`).

TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.

`stage()` (called by __graft_entry__.build() in the build container, where /root/reference exists) copies the four
pure-Python files of the reference's hot path into the git-ignored directory oracle/_ref/ with their package layout:

    models/size_invariant_timesformer.py
    models/efficientnet/efficientnet_pytorch/{__init__,model,utils}.py

oracle/_ref/ is NOT gpurun-ignored, so it travels to the GPU box like the built .so; nothing is copied into git history.
`load()` imports the staged modules (torch + einops + cv2 only) and `forward()` runs the reference's own loop body
(train.py:341-355 / predict.py:401-406) on the host cores.
"""
from __future__ import annotations

import os
import shutil
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
STAGE = os.path.join(HERE, "_ref")
FILES = ["models/size_invariant_timesformer.py", "models/efficientnet/efficientnet_pytorch/__init__.py",
         "models/efficientnet/efficientnet_pytorch/model.py", "models/efficientnet/efficientnet_pytorch/utils.py"]


def stage(reference_root: str = "/root/reference") -> bool:
    """Copy the reference files into oracle/_ref/ (byte-identical).  Returns False when the reference is absent."""
    if not os.path.isdir(reference_root):
        return False
    for rel in FILES:
        dst = os.path.join(STAGE, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(reference_root, rel), dst)
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(STAGE, rel)) for rel in FILES)


def load():
    """(EfficientNet, SizeInvariantTimeSformer) classes of the staged, unmodified reference."""
    if not available():
        raise RuntimeError("oracle/_ref is not staged: run __graft_entry__.build() where /root/reference exists")
    if STAGE not in sys.path:
        sys.path.insert(0, STAGE)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.efficientnet.efficientnet_pytorch import EfficientNet
        from models.size_invariant_timesformer import SizeInvariantTimeSformer
    return EfficientNet, SizeInvariantTimeSformer


def build_modules(esd, tsd, cfg, require_attention=True):
    EfficientNet, SizeInvariantTimeSformer = load()
    ext = EfficientNet.from_name("efficientnet-b0")
    ext.load_state_dict(esd, strict=True)
    model = SizeInvariantTimeSformer(config=cfg, require_attention=require_attention)
    model.load_state_dict(tsd, strict=True)
    return ext.eval(), model.eval()


def forward(ext, model, frames, meta):
    """frames (B,f,H,W,3) float32 raw 0..255 -> what the reference model returns (train.py:341-355)."""
    from einops import rearrange
    b = frames.shape[0]
    videos = rearrange(frames, "b f h w c -> (b f) c h w")                 # train.py:341
    features = ext(videos)                                                 # train.py:344-348
    features = rearrange(features, "(b f) c h w -> b f c h w", b=b)        # train.py:354
    return model(features, mask=meta["mask"], size_embedding=meta["size_embedding"],
                 identities_mask=meta["identities_mask"], positions=meta["positions"])      # train.py:355
