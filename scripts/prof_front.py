"""Runs the fused expand+depthwise kernel on the block-1..5 shapes at 512 images (for ncu / timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import ops
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
def r(*shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)
cases = [(112, 16, 96, 3, 2), (56, 24, 144, 3, 1), (56, 24, 144, 5, 2), (28, 40, 240, 5, 1), (28, 40, 240, 3, 2)]
only = os.environ.get("ONLY")
reps = int(os.environ.get("REPS", "1"))
for ci, (h, cin, cexp, k, s) in enumerate(cases):
    if only and str(ci) not in only.split(","):
        continue
    x = r(512, h, h, cin); we = r(cexp, cin, scale=cin ** -0.5); es = r(cexp, dtype=torch.float32)
    wt = r(k * k, cexp, scale=1.0 / k, dtype=torch.float32); ds = r(cexp, dtype=torch.float32)
    for _ in range(reps):
        ops.expand_dwconv(x, we, es, wt, ds, k, s)
    torch.cuda.synchronize()
    if reps > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.expand_dwconv(x, we, es, wt, ds, k, s)
        e1.record(); torch.cuda.synchronize()
        print(f"H{h} {cin}->{cexp} k{k} s{s}: {e0.elapsed_time(e1)/5*1e3:.1f} us")
print("done")
