"""Times the extractor's 1x1-convolution GEMM shapes at 512 images (CUDA events, 10 reps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import ops
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
def r(*shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)
# (hw, K, N, gated, act, resid)
cases = [(112 * 112, 16, 96, 0, 1, 0), (56 * 56, 24, 144, 0, 1, 0), (28 * 28, 40, 240, 0, 1, 0), (112 * 112, 32, 16, 1, 0, 0),
         (56 * 56, 144, 24, 1, 0, 1), (14 * 14, 112, 672, 0, 1, 0), (49, 1152, 192, 1, 0, 1), (14 * 14, 672, 112, 1, 0, 1),
         (14 * 14, 480, 80, 1, 0, 1), (49, 1152, 320, 1, 0, 0), (49, 192, 1152, 0, 1, 0), (14 * 14, 80, 480, 0, 1, 0),
         (49, 1152, 192, 0, 0, 1), (49, 1152, 192, 0, 0, 0), (49, 1152, 192, 1, 0, 0), (14 * 14, 672, 112, 0, 0, 0),
         (49, 320, 1280, 0, 1, 0)]
only = os.environ.get("ONLY")
n_img = 512
for ci, (hw, K, N, gated, act, res) in enumerate(cases):
    if only and str(ci) not in only.split(","):
        continue
    M = n_img * hw
    a = r(M, K); w = r(N, K, scale=K ** -0.5); sh = r(N, dtype=torch.float32)
    gate = torch.rand(n_img, K, device=dev) if gated else None
    resid = r(M, N) if res else None
    f = lambda: ops.pointwise(a, w, sh, act=act, gate=gate, rows_per_gate=hw if gated else 0, residual=resid)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    by = (M * K + M * N * (2 if res else 1) + N * K) * 2
    print(f"M{M} K{K} N{N} gate{gated} act{act} res{res}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s")
    del a, resid
