"""Host-side mirror of the reference's ``utils.aggregate_attentions`` (utils.py:68-96), computed on
the device by mt_aggregate_attn_fwd.  No CPU fallback."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import _lib


def aggregate_attentions_batched(attentions: Sequence[torch.Tensor], heads: int, num_frames: int,
                                 scale_factor: float = 50000.0) -> torch.Tensor:
    """attentions = [space_attn, time_attn], each (B*heads, 1, N) float32 on the GPU (what
    SizeInvariantTimeSformer returns with require_attention=True).  Returns (B, 3, num_frames):
    per video the softmaxed per-frame attention for space, time and space+time.

    The reference takes the max over batch AND heads (utils.py:75; it is only ever called with B = 1);
    this version keeps videos separate, which is identical for B = 1."""
    space, time = attentions
    if space.dim() == 3:
        space, time = space.squeeze(1), time.squeeze(1)
    space = space.contiguous().float()
    time = time.contiguous().float()
    rows, n_tok = space.shape
    if rows % heads:
        raise ValueError(f"{rows} attention rows are not a multiple of heads={heads}")
    b = rows // heads
    _lib.require_device(space.device)
    out = torch.empty((b, 3, num_frames), dtype=torch.float32, device=space.device)
    with torch.cuda.device(space.device):
        rc = _lib.load().mt_aggregate_attn_fwd(space.data_ptr(), time.data_ptr(), out.data_ptr(), b, heads, num_frames,
                                               n_tok, float(scale_factor), _lib.stream_ptr())
    _lib.check(rc, "mt_aggregate_attn_fwd")
    return out


def aggregate_attentions(attentions, heads, num_frames, frames_per_identity, scale_factor=50000) -> Tuple[List, List]:
    """Same signature and return values as the reference (utils.py:68-96) for a single video:
    ([space, time, combined] as lists of num_frames floats, identity_attentions)."""
    agg = aggregate_attentions_batched(attentions, heads, num_frames, scale_factor)
    if agg.shape[0] != 1:
        raise ValueError("aggregate_attentions mirrors the reference's single-video call; use "
                         "aggregate_attentions_batched for B > 1")
    aggregated = [agg[0, i].cpu().numpy() for i in range(3)]
    identity_attentions = []
    for index, identity_frames in enumerate(frames_per_identity):      # utils.py:88-95, kept verbatim in meaning
        if index == 0:
            identity_attention = sum(aggregated[-1][:identity_frames - 1])
        else:
            previous_identity_frames = frames_per_identity[index - 1]
            identity_attention = sum(aggregated[-1][previous_identity_frames - 1:identity_frames - 1])
        identity_attentions.append(identity_attention)
    return aggregated, identity_attentions


def build_clip_meta(slots: torch.Tensor, n_real: torch.Tensor, frame_no: torch.Tensor, ratio: torch.Tensor,
                    num_patches: int = 49, source: str = "predict"):
    """mask / identities_mask / size_embedding / positions of a batch of clips, assembled on the device
    (mt_clip_meta_fwd) from the per-identity slot table -- what ``DeepFakesDataset.__getitem__``
    (deepfakes_dataset.py:259-330) and predict.py's ``generate_masks`` build per clip on the host.

    ``source`` picks which of the reference's two assemblies is reproduced: "predict" (generate_masks: padded slots
    get mask 0) or "dataset" (DeepFakesDataset as executed: the mask is all ones, see csrc/timesformer.cu).
    slots, n_real: int32 (B, max_identities); frame_no, ratio: int32 (B, f); all on the same CUDA device.
    Returns a dict with the tensors ``SizeInvariantTimeSformer.forward`` takes: mask bool (B,f), identities_mask
    bool (B,f,f), size_embedding int32 (B,f), positions int64 (B, 1+f*num_patches)."""
    if source not in ("predict", "dataset"):
        raise ValueError("source must be 'predict' or 'dataset'")
    dev = slots.device
    _lib.require_device(dev)
    for t in (slots, n_real, frame_no, ratio):
        if t.dtype != torch.int32 or not t.is_contiguous() or t.device != dev:
            raise ValueError("build_clip_meta takes contiguous int32 tensors on one device")
    b, ids = slots.shape
    f = frame_no.shape[1]
    mask = torch.empty((b, f), dtype=torch.uint8, device=dev)
    idm = torch.empty((b, f, f), dtype=torch.uint8, device=dev)
    se = torch.empty((b, f), dtype=torch.int32, device=dev)
    pos = torch.empty((b, 1 + f * num_patches), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().mt_clip_meta_fwd(slots.data_ptr(), n_real.data_ptr(), frame_no.data_ptr(), ratio.data_ptr(), ids,
                                          1 if source == "predict" else 0, mask.data_ptr(), idm.data_ptr(), se.data_ptr(),
                                          pos.data_ptr(), b, f, num_patches, _lib.stream_ptr())
    _lib.check(rc, "mt_clip_meta_fwd")
    return {"mask": mask.view(torch.bool), "identities_mask": idm.view(torch.bool), "size_embedding": se, "positions": pos}
