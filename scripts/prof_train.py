"""The backward kernels of the training step at BASELINE shapes (B = 32, f = 16), a few launches each, for
`ncu --set full -k regex:'gemm_tc|attn_|grad_prep|layernorm_bwd|geglu_bwd'`.   python scripts/prof_train.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mintime_b200  # noqa: F401,E402
from mintime_b200 import ops  # noqa: E402

dev = "cuda:0"
B, f, n, heads = 32, 16, 49, 8
N = 1 + f * n
M = B * N
g = torch.Generator(device=dev).manual_seed(0)
bf = torch.bfloat16
dy = torch.randn((M, 1536), device=dev, generator=g).to(bf)
xn = torch.randn((M, 512), device=dev, generator=g).to(bf)
gres = torch.randn((M, 512), device=dev, generator=g)
x = torch.randn((M, 512), device=dev, generator=g)
gamma = torch.ones(512, device=dev)
h = torch.randn((M, 4096), device=dev, generator=g).to(bf)
dgo = torch.randn((M, 2048), device=dev, generator=g).to(bf)
qkv = (torch.randn((B, N, 1536), device=dev, generator=g) * 0.5).to(bf)
dao = torch.randn((B, N, 512), device=dev, generator=g).to(bf)
mask = torch.ones((B, f), dtype=torch.uint8, device=dev)
idm = torch.ones((B, f, f), dtype=torch.uint8, device=dev)
dw = torch.zeros((1536, 512), device=dev)
dwo = torch.zeros((512, 512), device=dev)
for _ in range(2):
    gb, _, cs = ops.grad_prep(gres, want_rm=True, want_colsum=True)
    _, dyT, _ = ops.grad_prep(dy, want_t=True)
    _, xT, _ = ops.grad_prep(xn, want_t=True)
    _, gbT, _ = ops.grad_prep(gb, want_t=True)
    ops.linear_wgrad_(dw, dyT, xT)            # to_qkv weight gradient: 1536 x 512 over 25152 tokens
    ops.linear_wgrad_(dwo, gbT, xT)           # to_out weight gradient: 512 x 512
    ops.divided_attention_bwd(qkv, dao, mask, idm, "space", f, n, heads)
    ops.divided_attention_bwd(qkv, dao, mask, idm, "time", f, n, heads)
    ops.layernorm_bwd_(gres, x, gamma, xn)
    ops.geglu_bwd(h, dgo)
torch.cuda.synchronize()
print("ok")
