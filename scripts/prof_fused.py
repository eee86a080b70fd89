"""Run the fused divided-attention kernel at the bench shape (B = 32, f = 16, n = 49):
   ncu --set full --clock-control none --import-source on -k regex:fused_attn -o gpurun_out/fused python scripts/prof_fused.py
   REPS=20 python scripts/prof_fused.py      # CUDA-event time per launch (library profiling hooks: GPU time, not host time)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import _lib, ops, weights
dev = "cuda:0"
B, f, n, heads = int(os.environ.get("B", 32)), int(os.environ.get("F", 16)), 49, 8
N = 1 + f * n
g = torch.Generator(device=dev).manual_seed(0)
xn = torch.randn(B, N, 512, device=dev, generator=g).bfloat16()
w = (torch.randn(1536, 512, device=dev, generator=g) * 0.04).bfloat16()
wh = weights.qkv_per_head(w, heads, 64)
mask = torch.ones(B, f, dtype=torch.uint8, device=dev); idm = torch.ones(B, f, f, dtype=torch.uint8, device=dev)
reps = int(os.environ.get("REPS", 2))
lib = _lib.load()
for mode in ("time", "space"):
    for it in range(2):
        ops.fused_attention(xn, wh, mask, idm, mode, f, n, heads, want_cls_attn=False)
    torch.cuda.synchronize()
    if reps > 2:
        lib.mt_prof_reset(); lib.mt_prof_enable(1)
        for it in range(reps):
            ops.fused_attention(xn, wh, mask, idm, mode, f, n, heads, want_cls_attn=False)
        torch.cuda.synchronize()
        lib.mt_prof_enable(0)
        for name, ms, flops, byts, cnt in _lib.profile_collect():
            print(f"{name}: {ms / cnt * 1e3:.1f} us per launch, {flops / ms / 1e9:.0f} TFLOP/s algorithmic")
if os.environ.get("UNFUSED"):
    qkv = torch.randn(B, N, 1536, device=dev, generator=g).bfloat16()
    lib.mt_prof_reset(); lib.mt_prof_enable(1)
    for it in range(reps):
        ops.pointwise(xn.view(B * N, 512), w)
        ops.divided_attention(qkv, mask, idm, "time", f, n, heads, want_cls_attn=False)
        ops.divided_attention(qkv, mask, idm, "space", f, n, heads, want_cls_attn=False)
    torch.cuda.synchronize(); lib.mt_prof_enable(0)
    for name, ms, flops, byts, cnt in _lib.profile_collect():
        print(f"{name}: {ms / cnt * 1e3:.1f} us per launch")
print("done")
