"""Golden vectors for the clip-metadata builder and the attention aggregation, produced by EXECUTING the reference's
own code (TEST INFRASTRUCTURE: run once in the build container, where /root/reference exists; the .npz/.json travel).

  * ``DeepFakesDataset.__getitem__`` (deepfakes_dataset.py:206-339) cannot be imported as a module here (albumentations,
    python-magic, the repo-local ``transforms`` package are absent), so the class definition is cut out of the source
    file by ``ast`` and executed unmodified in a namespace that provides what it needs at run time (torch, numpy, cv2,
    os, re, random, statistics.mean, torch's Dataset).  It then runs against a synthetic on-disk clip: identity folders
    of PNG faces named ``<frame>_<i>.png`` and a real .mp4 written with cv2.VideoWriter (the size-embedding ratio is
    face area / video area, :259-263).  Only the augmentation pipeline is replaced (``create_val_transform`` returns a
    plain resize): it does not take part in the tensors under test.
  * ``aggregate_attentions`` (utils.py:68-96) is cut out of utils.py the same way (the module imports matplotlib /
    pytorchvideo at the top) and executed unmodified.

Writes tests/golden/clip_meta_ref.json and tests/golden/aggregate_attn_ref.npz.
"""
from __future__ import annotations

import ast
import json
import os
import random
import re
import shutil
import sys
import tempfile
from statistics import mean

import cv2
import numpy as np
import torch
from einops import rearrange
from scipy.special import softmax
from torch.utils.data import Dataset

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def cut(path: str, names) -> str:
    """source text of the top-level definitions / assignments called `names`, in file order"""
    src = open(path).read()
    tree = ast.parse(src)
    lines = src.splitlines(keepends=True)
    out = []
    for node in tree.body:
        tgt = None
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)):
            tgt = node.name
        elif isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            tgt = node.targets[0].id
        if tgt in names:
            out.append("".join(lines[node.lineno - 1:node.end_lineno]))
    return "\n".join(out)


def load_dataset_class():
    ns = {"torch": torch, "np": np, "os": os, "cv2": cv2, "random": random, "re": re, "mean": mean, "Dataset": Dataset}
    code = cut(os.path.join(REF, "deepfakes_dataset.py"), {"MODES", "RANGE_SIZE", "SIZE_EMB_DICT", "DeepFakesDataset"})
    exec(compile(code, "deepfakes_dataset.py[cut]", "exec"), ns)
    return ns["DeepFakesDataset"]


def load_generate_masks():
    from PIL import Image  # noqa: F401
    ns = {"torch": torch, "np": np, "os": os, "cv2": cv2,
          # the augmentation pipeline is not under test: same call signature, plain resize
          "create_val_transform": lambda size, additional_targets: (
              lambda **imgs: {k: cv2.resize(v, (size, size)) for k, v in imgs.items()})}
    exec(compile(cut(os.path.join(REF, "predict.py"), {"RANGE_SIZE", "SIZE_EMB_DICT", "generate_masks"}),
                 "predict.py[cut]", "exec"), ns)
    return ns["generate_masks"]


def load_aggregate():
    ns = {"torch": torch, "np": np, "rearrange": rearrange, "softmax": softmax, "mean": mean}
    exec(compile(cut(os.path.join(REF, "utils.py"), {"aggregate_attentions"}), "utils.py[cut]", "exec"), ns)
    return ns["aggregate_attentions"]


VIDEO_W, VIDEO_H = 200, 100      # video area 20000 -> ratio = int(face_h * face_w * 100 / 20000)


def face_for_ratio(ratio: int):
    """an (h, w) whose area ratio to the 200x100 video is exactly `ratio` per cent"""
    return 10, 20 * ratio if ratio > 0 else 1        # h*w = 200*ratio -> 200*ratio*100/20000 = ratio


# name -> (num_frames, max_identities, identity attention, index, [[(frame, ratio), ...] per identity])
# identities_ordering = 1 sorts identities by their number of faces (descending): all counts below are distinct
SCENARIOS = {
    "one_identity_f8": (8, 3, True, 0, [[(3, 4), (5, 5), (9, 6), (12, 100), (14, 37), (20, 11), (21, 55), (22, 96)]]),
    "two_ids_padded_f8": (8, 3, True, 0, [[(3, 0), (5, 5), (9, 6), (12, 100)], [(5, 17), (9, 50)]]),
    "two_ids_padded_no_identity_attention_f8": (8, 3, False, 0, [[(3, 0), (5, 5), (9, 6), (12, 100)], [(5, 17), (9, 50)]]),
    "three_ids_f16": (16, 3, True, 0, [[(1, 10), (2, 12), (3, 14), (4, 16), (5, 18)], [(2, 31), (3, 35), (7, 39), (8, 41)],
                                       [(1, 70), (8, 75), (9, 80)]]),
    "four_ids_f16": (16, 4, True, 0, [[(f, 2 + f) for f in range(1, 6)], [(f, 20 + f) for f in (2, 4, 6, 8)],
                                      [(3, 44), (30, 46), (31, 48)], [(10, 91), (11, 99)]]),
    "subsampled_even_index_f8": (8, 3, True, 0, [[(f, f) for f in range(1, 14)]]),
    "subsampled_odd_index_f8": (8, 3, True, 1, [[(f, f) for f in range(1, 14)]]),
    "two_ids_one_overfull_f16": (16, 2, True, 1, [[(f, 3 * f) for f in range(1, 12)], [(2, 8), (5, 9), (6, 10)]]),
}


def build_clip(root: str, mode: str, vid: str, identities):
    clip = os.path.join(root, "faces", mode, vid)
    for i, faces in enumerate(identities):
        d = os.path.join(clip, f"identity_{i}")
        os.makedirs(d)
        for j, (frame, ratio) in enumerate(faces):
            h, w = face_for_ratio(ratio)
            assert cv2.imwrite(os.path.join(d, f"{frame}_{j}.png"), np.full((h, w, 3), 7 * i + j, np.uint8))
    vdir = os.path.join(root, "videos", mode)
    os.makedirs(vdir, exist_ok=True)
    wr = cv2.VideoWriter(os.path.join(vdir, vid + ".mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 5, (VIDEO_W, VIDEO_H))
    assert wr.isOpened()
    for _ in range(3):
        wr.write(np.zeros((VIDEO_H, VIDEO_W, 3), np.uint8))
    wr.release()


def run_scenarios():
    DS = load_dataset_class()
    root = tempfile.mkdtemp(prefix="mintime_clipmeta_")
    cases = {}
    try:
        for name, (f, max_ids, ident_attn, index, identities) in SCENARIOS.items():
            vid = name
            build_clip(root, "test", vid, identities)
            paths = ["pad_a", "pad_b"]
            paths[index] = os.path.join("test", vid)
            ds = DS(paths, [0, 1], os.path.join(root, "faces"), os.path.join(root, "videos"), image_size=8, mode="test",
                    num_frames=f, max_identities=max_ids, num_patches=49, enable_identity_attention=ident_attn,
                    identities_ordering=1)
            # the augmentation pipeline is not under test: same call signature, plain resize
            ds.create_val_transform = lambda size, additional_targets: (
                lambda **imgs: {k: cv2.resize(v, (size, size)) for k, v in imgs.items()})
            seq, size_emb, mask, idmask, positions, label = ds[index]
            assert seq.shape == (f, 8, 8, 3)
            # the slot table the reference worked with (sorted identities and their slot counts), re-derived through
            # the reference's own get_sorted_identities; the faces it read are the (sorted, subsampled) file names
            sorted_ids, discarded = ds.get_sorted_identities(os.path.join(root, "faces", "test", vid))
            table = []
            for path, _side, max_faces in sorted_ids:
                k = int(os.path.basename(path).split("_")[1])
                faces = sorted(identities[k], key=lambda t: t[0])
                if len(faces) > max_faces:                                    # deepfakes_dataset.py:238-244
                    if index % 2:
                        idx = np.round(np.linspace(0, len(faces) - 2, max_faces)).astype(int)
                    else:
                        idx = np.round(np.linspace(1, len(faces) - 1, max_faces)).astype(int)
                    faces = [faces[i] for i in idx]
                table.append([int(max_faces), [[int(a), int(b)] for a, b in faces]])
            cases[name] = {
                "source": "dataset", "num_frames": f, "num_patches": 49, "enable_identity_attention": ident_attn,
                "identities": table, "size_embedding": size_emb.tolist(), "mask": [int(v) for v in mask.tolist()],
                "identities_mask": [[int(v) for v in row] for row in idmask.tolist()], "positions": positions.tolist(),
            }
            # the same slot table through predict.py's generate_masks (:254-352): faces are (frame, PIL image) pairs
            # already in memory; its ratio is face w*h against HALF the video area (:292-293)
            if all(len(faces) <= mf for mf, faces in table):
                from PIL import Image
                gm = load_generate_masks()
                idents = []
                for i, (mf, faces) in enumerate(table):
                    imgs = []
                    for fr, ra in faces:
                        h, w = face_for_ratio(ra // 2)       # (a larger face has no size bucket there: IndexError)
                        imgs.append((fr, Image.fromarray(np.zeros((h, w, 3), np.uint8))))
                    idents.append([f"id{i}", 0, mf, imgs])
                video = os.path.join(root, "videos", "test", vid + ".mp4")
                _seq, se2, mask2, idm2, pos2, _tok = gm(video, idents, [], f, 8, 49)
                cases[name + "@predict"] = {
                    "source": "predict", "num_frames": f, "num_patches": 49, "enable_identity_attention": True,
                    # ratio as generate_masks computes it for these images: int(w*h*100 / (W*H/2))
                    "identities": [[mf, [[fr, int(face_for_ratio(ra // 2)[0] * face_for_ratio(ra // 2)[1] * 100 /
                                                  (VIDEO_W * VIDEO_H / 2))] for fr, ra in faces]] for mf, faces in table],
                    "size_embedding": se2[0].tolist(), "mask": [int(v) for v in mask2[0].tolist()],
                    "identities_mask": [[int(v) for v in row] for row in idm2[0].tolist()], "positions": pos2[0].tolist(),
                }
    finally:
        shutil.rmtree(root, ignore_errors=True)
    with open(os.path.join(OUT, "clip_meta_ref.json"), "w") as fh:
        json.dump(cases, fh, separators=(",", ":"))
    print("clip_meta_ref.json:", {k: (v["num_frames"], [t[0] for t in v["identities"]]) for k, v in cases.items()})


def run_aggregate():
    agg = load_aggregate()
    out = {}
    for name, (heads, f, n, seed, scale) in {"h8_f16": (8, 16, 49, 1, 50000), "h8_f8": (8, 8, 49, 2, 50000),
                                             "h4_f8_n16_scale100": (4, 8, 16, 3, 100)}.items():
        g = torch.Generator().manual_seed(seed)
        N = 1 + f * n
        space = torch.softmax(torch.randn((heads, 1, N), generator=g) * 2, dim=-1)      # b = 1, as predict.py calls it
        time = torch.softmax(torch.randn((heads, 1, N), generator=g) * 2, dim=-1)
        aggregated, _ = agg([space, time], heads, f, [f], scale_factor=scale)
        out[name + ".space_in"] = space.numpy()
        out[name + ".time_in"] = time.numpy()
        out[name + ".out"] = np.stack([np.asarray(a, np.float64) for a in aggregated])   # (3, f): space, time, combined
        out[name + ".meta"] = np.asarray([heads, f, N, scale], np.int64)
    np.savez_compressed(os.path.join(OUT, "aggregate_attn_ref.npz"), **out)
    print("aggregate_attn_ref.npz:", sorted(k for k in out if k.endswith(".out")))


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container)")
    run_scenarios()
    run_aggregate()
