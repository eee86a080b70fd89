"""Whole-step time at batch B (device-resident inputs): eager module calls vs one CUDA-graph replay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import synth
from mintime_b200.spec import default_tsf_config
from mintime_b200.graphed import GraphedHotPath
dev = "cuda:0"
B, f = int(os.environ.get("B", "32")), 16
cfg = default_tsf_config(num_frames=f)
ext = mintime_b200.EfficientNet.from_name("efficientnet-b0"); ext.load_state_dict(synth.make_effnet_state_dict(1234)); ext = ext.to(dev).eval()
model = mintime_b200.SizeInvariantTimeSformer(config=cfg); model.load_state_dict(synth.make_tsf_state_dict(cfg, 4321)); model = model.to(dev).eval()
meta = synth.make_batch_meta(B, f, [1], seed=3)
vid = synth.make_frames(B, f, seed=3, mask=meta["mask"], dtype=torch.uint8).to(dev).float()
md = {k: v.to(dev) for k, v in meta.items()}
def eager():
    with torch.no_grad():
        x = vid.view(B * f, 224, 224, 3).permute(0, 3, 1, 2)
        return model(ext(x).reshape(B, f, 1280, 7, 7), mask=md["mask"], size_embedding=md["size_embedding"],
                     identities_mask=md["identities_mask"], positions=md["positions"])
hot = GraphedHotPath(ext, model, B, f, frame_dtype=torch.float32, device=dev)
hot.static["videos"].copy_(vid)
for k in ("mask", "identities_mask", "size_embedding", "positions"): hot.static[k].copy_(md[k])
def graphed():
    hot.graph.replay()
for name, fn in (("eager", eager), ("graph", graphed)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20:.3f} ms per step (B={B})")
