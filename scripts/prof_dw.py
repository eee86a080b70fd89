"""Runs every depthwise layer shape of EfficientNet-B0 once at 512 images (for ncu / per-kernel timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import ops, _lib
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
def r(*shape, dtype=torch.bfloat16):
    return torch.randn(*shape, device=dev, generator=g).to(dtype)
cases = [(112, 32, 3, 1), (112, 96, 3, 2), (56, 144, 3, 1), (56, 144, 5, 2), (28, 240, 5, 1), (28, 240, 3, 2),
         (14, 480, 3, 1), (14, 480, 5, 1), (14, 672, 5, 1), (14, 672, 5, 2), (7, 1152, 5, 1), (7, 1152, 3, 1)]
n = int(os.environ.get("N_IMG", "512"))
reps = int(os.environ.get("REPS", "1"))
lib = _lib.load()
for (h, c, k, s) in cases:
    x = r(n, h, h, c); wt = r(k * k, c, dtype=torch.float32); sh = r(c, dtype=torch.float32)
    for it in range(reps):
        ops.dwconv(x, wt, sh, k, s)
    torch.cuda.synchronize()
    if reps > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(10):
            ops.dwconv(x, wt, sh, k, s)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        ho = (h + s - 1) // s
        by = n * c * (h * h + ho * ho) * 2
        print(f"H{h} C{c} k{k} s{s}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s  {2*k*k*n*ho*ho*c/ms/1e9:.2f} TFLOP/s")
print("done")
