"""Training-step timing (BASELINE configs[3] shape on one GPU): frozen extractor forward + transformer forward /
backward + SGD step, per-kernel table from the mt_prof_* hooks.   python scripts/train_step_bench.py [B] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mintime_b200 as mt                                  # noqa: E402
from mintime_b200 import _lib, synth                       # noqa: E402
from mintime_b200.spec import default_tsf_config           # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = "cuda:0"
f = 16
cfg = default_tsf_config(num_frames=f)
ext = mt.EfficientNet.from_name("efficientnet-b0", precision="bf16")
ext.load_state_dict(synth.make_effnet_state_dict(1234))
ext = ext.to(dev).eval()
model = mt.SizeInvariantTimeSformer(config=cfg, precision="bf16")
model.load_state_dict(synth.make_tsf_state_dict(cfg, 4321))
model = model.to(dev).train()
opt = torch.optim.SGD(model.parameters(), lr=0.01, weight_decay=1e-4)
meta = synth.make_batch_meta(B, f, [1], seed=1234)
frames = synth.make_frames(B, f, seed=1234, mask=meta["mask"]).to(dev)
labels = (torch.rand((B, 1), generator=torch.Generator().manual_seed(0)) < 0.55).float().to(dev)
lossf = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([0.8169], device=dev))
mask, idm, pos = meta["mask"].to(dev), meta["identities_mask"].to(dev), meta["positions"].to(dev)
se = meta["size_embedding"].to(dev)


def step():
    with torch.no_grad():
        feats = ext(frames.permute(0, 1, 4, 2, 3).reshape(B * f, 3, 224, 224))
    feats = feats.view(B, f, *feats.shape[1:])
    opt.zero_grad(set_to_none=True)
    y = model(feats, mask=mask, size_embedding=se, identities_mask=idm, positions=pos)
    loss = lossf(y, labels)
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    l = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    l = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(json.dumps({"train_step_ms": ms, "videos_per_s": B / ms * 1e3, "batch": B, "loss": l.item(),
                  "max_mem_GiB": torch.cuda.max_memory_allocated() / 2**30}))
lib = _lib.load()
lib.mt_prof_reset()
lib.mt_prof_enable(1)
step()
torch.cuda.synchronize()
rows = _lib.profile_collect()
lib.mt_prof_enable(0)
tot = sum(r[1] for r in rows)
print(f"profiled kernels: {tot:.2f} ms in {sum(r[4] for r in rows)} launches")
for name, ms_, fl, by, cnt in rows[:28]:
    print(f"{name:40s} {ms_:8.3f} ms  x{cnt:4d}  {fl / ms_ / 1e9 if ms_ else 0:8.1f} TF/s  {by / ms_ / 1e6 if ms_ else 0:8.1f} GB/s")
