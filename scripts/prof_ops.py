"""Run the hot kernels once each at the bench workload shapes (for ncu captures).
   ncu --set full -k regex:<pattern> ... python scripts/prof_ops.py [gemm|attn|dw|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import ops, weights
which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = "cuda:0"
B, f, n = 32, 16, 49
M = B * (1 + f * n)
bf = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
def r(*shape, scale=1.0, dtype=bf):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)
for it in range(2):      # first pass = warm-up
    if which in ("gemm", "all"):
        a = r(M, 512); w1 = r(4096, 512, scale=0.04); b1 = r(4096, dtype=torch.float32)
        ops.linear_geglu(a, w1, b1)                                       # FF1 + GEGLU
        wq = r(1536, 512, scale=0.04); ops.pointwise(a, wq)               # to_qkv
        x = r(M, 512, dtype=torch.float32); wo = r(512, 512, scale=0.04)
        ops.linear_residual_(x, a, wo, b1[:512].contiguous())             # to_out + residual
        h = r(M, 2048); w2 = r(512, 2048, scale=0.02)
        ops.linear_residual_(x, h, w2, b1[:512].contiguous())             # FF2 + residual
        a1 = r(512 * 112 * 112, 16); we = r(96, 16, scale=0.2); sh = r(96, dtype=torch.float32)
        ops.pointwise(a1, we, sh, act=1)                                  # block-1 expand
        del a1
        d = r(512 * 56 * 56, 144); wp = r(24, 144, scale=0.1); gate = torch.rand(512, 144, device=dev)
        ops.pointwise(d, wp, sh[:24].contiguous(), gate=gate, rows_per_gate=56 * 56)   # block-2 project (+SE gate)
        del d
    if which in ("attn", "all"):
        qkv = r(B, 1 + f * n, 1536, scale=0.5)
        mask = torch.ones(B, f, dtype=torch.uint8, device=dev); idm = torch.ones(B, f, f, dtype=torch.uint8, device=dev)
        ops.divided_attention(qkv, mask, idm, "time", f, n, 8)
        ops.divided_attention(qkv, mask, idm, "space", f, n, 8)
    if which in ("dw", "all"):
        x = r(512, 112, 112, 96); wt = r(9, 96, dtype=torch.float32); sh = r(96, dtype=torch.float32)
        ops.dwconv(x, wt, sh, 3, 2)
        del x
        x = r(512, 14, 14, 672); wt = r(25, 672, dtype=torch.float32); sh = r(672, dtype=torch.float32)
        ops.dwconv(x, wt, sh, 5, 1)
        xs = torch.randint(0, 256, (512, 224, 224, 3), device=dev, dtype=torch.uint8).float()
        ops.stem(xs, r(27, 32, dtype=torch.float32), r(32, dtype=torch.float32))
        del xs, x
    torch.cuda.synchronize()
print("done")
