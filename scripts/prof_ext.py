"""One eval forward of the extractor at 512 images (bf16) for ncu captures:
   ncu --set full --clock-control none --import-source on -s <launches of 2 warm-up forwards> -o gpurun_out/ext python scripts/prof_ext.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import synth
n = int(os.environ.get("N", 512))
ext = mintime_b200.EfficientNet.from_name("efficientnet-b0", precision="bf16")
ext.load_state_dict(synth.make_effnet_state_dict(1234, conditioned=True))
ext = ext.to("cuda:0").eval()
x = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda:0").permute(0, 3, 1, 2)
with torch.no_grad():
    for _ in range(3):
        y = ext(x)
torch.cuda.synchronize()
print("done", float(y.float().abs().mean()))
