"""CPU restatement of the clip-metadata assembly of the reference (TEST INFRASTRUCTURE ONLY: imported by tests/,
never by the product path).

Follows ``DeepFakesDataset.__getitem__`` (deepfakes_dataset.py:259-330) and predict.py ``generate_masks`` (:254-352)
statement by statement, with the file reads replaced by in-memory (frame number, area ratio) pairs.  PARITY UNPINNED
against an executed reference: the dataset class needs albumentations / python-magic / image files and cannot run in
this container, and the reference ships no fixture for it; the known-answer case in tests/test_host_logic.py is
worked out by hand from the cited lines.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

RANGE_SIZE = 5                                                                                     # :30
SIZE_EMB_DICT = [(1 + i * RANGE_SIZE, (i + 1) * RANGE_SIZE) if i != 0 else (0, RANGE_SIZE) for i in range(20)]   # :31


def clip_meta(identities: Sequence[Tuple[int, List[Tuple[int, int]]]], num_frames: int, num_patches: int = 49,
              enable_identity_attention: bool = True):
    """identities: [(max_faces, [(frame_number, ratio), ...faces read for this identity]), ...] with sum(max_faces) ==
    num_frames.  Returns (size_embeddings[f] int32, mask[f] bool, identities_mask[f][f] bool, positions[1+f*n] int64)."""
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
            own = images_frames[len(images_frames) - n_read:] if n_read else []
            images_frames.extend([max(own) if own else 0 for _ in range(diff)])                    # :277-281
        if enable_identity_attention and n_read < max_faces:                                       # :283
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
