"""Single-video latency (predict.py use case: b = 1, f = 16): eager nn.Module calls vs one CUDA-graph replay."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import synth
from mintime_b200.spec import default_tsf_config
from mintime_b200.graphed import GraphedHotPath
dev = "cuda:0"
f = 16
cfg = default_tsf_config(num_frames=f)
ext = mintime_b200.EfficientNet.from_name("efficientnet-b0"); ext.load_state_dict(synth.make_effnet_state_dict(1234)); ext = ext.to(dev).eval()
model = mintime_b200.SizeInvariantTimeSformer(config=cfg, require_attention=True); model.load_state_dict(synth.make_tsf_state_dict(cfg, 4321)); model = model.to(dev).eval()
meta = synth.make_batch_meta(1, f, [2], seed=3)
vid = synth.make_frames(1, f, seed=3, mask=meta["mask"], dtype=torch.uint8).pin_memory()
md = {k: v.to(dev) for k, v in meta.items()}
def eager():
    with torch.no_grad():
        x = vid.to(dev, non_blocking=True).view(f, 224, 224, 3).permute(0, 3, 1, 2)
        out = model(ext(x).reshape(1, f, 1280, 7, 7), mask=md["mask"], size_embedding=md["size_embedding"],
                    identities_mask=md["identities_mask"], positions=md["positions"])
        return out[0].cpu()
hot = GraphedHotPath(ext, model, 1, f, device=dev)
def graphed():
    out = hot(vid, md["mask"], md["identities_mask"], md["size_embedding"], md["positions"])
    return out[0].cpu()
for name, fn in (("eager", eager), ("graph", graphed)):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): fn()
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per video (H2D of the clip + forward + logits D2H)")
