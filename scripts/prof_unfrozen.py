"""Kernel-time table (torch.profiler / CUPTI) of one eager training step with the extractor in train mode (B = 8 x 16 frames)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import synth
from mintime_b200.spec import default_tsf_config
dev = "cuda:0"
B, f = int(os.environ.get("B", 8)), 16
cfg = default_tsf_config(num_frames=f, channels=1280)
ext = mintime_b200.EfficientNet.from_name("efficientnet-b0", precision="bf16")
ext.load_state_dict(synth.make_effnet_state_dict(1234, conditioned=True)); ext = ext.to(dev).train()
model = mintime_b200.SizeInvariantTimeSformer(config=cfg, precision="bf16")
model.load_state_dict(synth.make_tsf_state_dict(cfg, 4321)); model = model.to(dev).train()
opt = torch.optim.SGD(list(model.parameters()) + list(ext.parameters()), lr=0.01)
meta = synth.make_batch_meta(B, f, [1], seed=1)
frames = synth.make_frames(B, f, seed=1, mask=meta["mask"], dtype=torch.uint8).to(dev)
labels = torch.ones((B, 1), device=dev)
lossf = torch.nn.BCEWithLogitsLoss()
def step():
    feats = ext(frames.view(B * f, 224, 224, 3).permute(0, 3, 1, 2)).reshape(B, f, 1280, 7, 7)
    opt.zero_grad(set_to_none=True)
    y = model(feats, mask=meta["mask"].to(dev), size_embedding=meta["size_embedding"].to(dev),
              identities_mask=meta["identities_mask"].to(dev), positions=meta["positions"].to(dev))
    loss = lossf(y, labels); loss.backward(); opt.step()
step(); step(); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=70))
