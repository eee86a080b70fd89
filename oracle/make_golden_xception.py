"""Generate tests/golden/xception_*.npz / .json by running the UNMODIFIED reference ``models/xception.py`` from
/root/reference (build container only):

    python oracle/make_golden_xception.py

  * xception_keys.json     : the reference module's ``state_dict`` keys and shapes (num_classes=1, train.py:131)
  * xception_b2.npz        : reference ``Xception.forward`` on 2 seeded 224x224 faces with the seeded weights of
                             ``synth.make_xception_state_dict``: per-stage samples (conv1, conv2, block1..12, conv3, conv4)
  * xception_tsf_b1_f8.npz : Xception -> SizeInvariantTimeSformer(channels=2048) on one 8-frame clip: logits + attention maps
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
REF = os.environ.get("MINTIME_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import mintime_b200  # noqa: E402,F401
from mintime_b200 import synth  # noqa: E402
from mintime_b200.spec import default_tsf_config  # noqa: E402
from make_golden import sample  # noqa: E402

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    from models.xception import xception  # noqa: E402
    from models.size_invariant_timesformer import SizeInvariantTimeSformer  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
torch.manual_seed(0)
ref = xception(num_classes=1)
with open(os.path.join(GOLD, "xception_keys.json"), "w") as fh:
    json.dump({k: list(v.shape) for k, v in ref.state_dict().items()}, fh, indent=0)
sd = synth.make_xception_state_dict(2468)
ref.load_state_dict(sd)
ref.eval()

# ---- extractor alone
meta = synth.make_batch_meta(1, 2, [1], seed=21, pad_tail=False)
frames = synth.make_frames(1, 2, seed=21, mask=meta["mask"])            # (1,2,224,224,3) raw 0..255
x = frames.view(2, 224, 224, 3).permute(0, 3, 1, 2).contiguous()
with torch.no_grad():
    out, feats = ref.features(x)
names = ["conv2"] + [f"block{i}" for i in range(1, 13)]
store = {"out": sample(out, 8192), "out_mean_abs": np.float32(out.abs().mean())}
for n, t in zip(names, feats):
    store["stage." + n] = sample(t)
    store["stage." + n + ".mean_abs"] = np.float32(t.abs().mean())
np.savez_compressed(os.path.join(GOLD, "xception_b2.npz"), **store)
print("xception_b2", tuple(out.shape), float(out.abs().mean()), {n: float(t.abs().mean()) for n, t in zip(names, feats)})

# ---- extractor -> transformer (channels = 2048, the shipped yaml's value)
B, f = 1, 8
cfg = default_tsf_config(num_frames=f, channels=2048)
tsd = synth.make_tsf_state_dict(cfg, 4321)
model = SizeInvariantTimeSformer(config=cfg, require_attention=True)
model.load_state_dict(tsd)
model.eval()
meta = synth.make_batch_meta(B, f, [2], seed=33, pad_tail=True)
frames = synth.make_frames(B, f, seed=33, mask=meta["mask"])
x = frames.permute(0, 1, 4, 2, 3).reshape(B * f, 3, 224, 224)
with torch.no_grad():
    feats = ref(x)
    logits, (space, time) = model(feats.view(B, f, 2048, 7, 7), mask=meta["mask"], size_embedding=meta["size_embedding"],
                                  identities_mask=meta["identities_mask"], positions=meta["positions"])
np.savez_compressed(os.path.join(GOLD, "xception_tsf_b1_f8.npz"), features=sample(feats, 8192), logits=logits.numpy(),
                    space_attn=space.numpy(), time_attn=time.numpy())
print("xception_tsf", logits.flatten().tolist(), float(feats.abs().mean()))
