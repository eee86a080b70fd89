"""CPU restatement of the clip-metadata assembly of the reference (TEST INFRASTRUCTURE ONLY: imported by tests/,
never by the product path).

Follows ``DeepFakesDataset.__getitem__`` (deepfakes_dataset.py:259-330) and predict.py ``generate_masks`` (:254-352)
statement by statement, with the file reads replaced by in-memory (frame number, area ratio) pairs, and
``aggregate_attentions`` (utils.py:68-86).  PINNED against the executed reference: oracle/make_golden_clip_meta.py cuts
the reference's ``DeepFakesDataset`` class / ``aggregate_attentions`` function out of their source files, runs them
unmodified (synthetic on-disk clips; only the augmentation pipeline is replaced) and stores inputs + outputs in
tests/golden/clip_meta_ref.json / aggregate_attn_ref.npz; tests/test_oracle.py holds this file to them.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

RANGE_SIZE = 5                                                                                     # :30
SIZE_EMB_DICT = [(1 + i * RANGE_SIZE, (i + 1) * RANGE_SIZE) if i != 0 else (0, RANGE_SIZE) for i in range(20)]   # :31


def clip_meta(identities: Sequence[Tuple[int, List[Tuple[int, int]]]], num_frames: int, num_patches: int = 49,
              enable_identity_attention: bool = True, source: str = "dataset"):
    """identities: [(max_faces, [(frame_number, ratio), ...faces read for this identity]), ...] with sum(max_faces) ==
    num_frames.  Returns (size_embeddings[f] int32, mask[f] bool, identities_mask[f][f] bool, positions[1+f*n] int64).
    source = "dataset": DeepFakesDataset.__getitem__; "predict": predict.py generate_masks (masks the padded slots)."""
    mask: List[int] = []
    size_embeddings: List[float] = []
    images_frames: List[int] = []
    for max_faces, faces in identities:
        identity_size_embeddings: List[float] = []
        n_read = 0
        for frame, ratio in faces[:max_faces]:
            side_ranges = list(map(lambda a_: ratio in range(a_[0], a_[1] + 1), SIZE_EMB_DICT))    # :262
            identity_size_embeddings.append(np.where(side_ranges)[0][0] + 1)                       # :263
            images_frames.append(frame)                                                            # :266-267
            n_read += 1
        diff = 0
        if n_read < max_faces:                                                                     # :273
            diff = max_faces - len(identity_size_embeddings)
            identity_size_embeddings = list(identity_size_embeddings) + [0] * diff                 # :275
            # max over the CLIP-wide list (earlier identities included); empty list -> 0 (the except branch)
            images_frames.extend([max(images_frames) if images_frames else 0 for _ in range(diff)])   # :277-281
            n_images = max_faces                   # identity_images.extend(...) (:276): the list is full from here on
        else:
            n_images = n_read
        if source == "predict":
            masked = n_read < max_faces            # predict.py:300-306: the mask is built inside the padding branch
        else:
            # :283 re-tests len(identity_images) AFTER :276 padded it, so as executed this is never true and
            # DeepFakesDataset always returns an all-ones mask, whatever enable_identity_attention says
            masked = enable_identity_attention and n_images < max_faces
        if masked:
            mask.extend([1 if i < max_faces - diff else 0 for i in range(max_faces)])              # :284
        else:
            mask.extend([1 for _ in range(max_faces)])                                             # :286
        size_embeddings.extend(identity_size_embeddings)                                           # :289
    identities_mask = []                                                                           # :314-321
    last_range_end = 0
    for max_faces, _ in identities:
        identity_mask = [True if last_range_end <= i < last_range_end + max_faces else False for i in range(num_frames)]
        for _k in range(max_faces):
            identities_mask.append(identity_mask)
        last_range_end += max_faces
    images_frames_positions = {k: v + 1 for v, k in enumerate(sorted(set(images_frames)))}         # :324
    frame_positions = [images_frames_positions[frame] for frame in images_frames]                  # :325
    positions = [[i + 1 for i in range((fp - 1) * num_patches, num_patches * fp)] for fp in frame_positions]   # :327
    positions = sum(positions, [])
    positions.insert(0, 0)                                                                         # :329
    return (np.asarray(size_embeddings, np.int32), np.asarray(mask, bool), np.asarray(identities_mask, bool),
            np.asarray(positions, np.int64))


def aggregate_attentions(attentions, heads: int, num_frames: int, scale_factor: float = 50000.0):
    """utils.py:68-86 for one video: attentions = [space, time], each (heads, 1, N) float arrays.
    Returns (3, num_frames) float64: softmaxed per-frame attention of space, time, space + time."""
    aggregated = []
    for attention in attentions:
        a = np.asarray(attention, np.float64).reshape(-1, heads, np.asarray(attention).shape[-1])   # '(b h) t -> b h t' (:74)
        aggregated.append([a[:, :, i].max() for i in range(a.shape[2])])                             # :75
    aggregated.append(list(np.sum(np.asarray(aggregated), axis=0)))                                  # :79
    out = []
    for lst in aggregated:                                                                           # :83-85
        chunks = np.array_split(np.asarray(lst), num_frames)
        v = np.asarray([float(np.mean(c)) * scale_factor for c in chunks], np.float64)
        e = np.exp(v - v.max())                                                                      # scipy softmax
        out.append(e / e.sum())
    return np.stack(out)
