"""The MN-major weight-gradient GEMM at the B = 32 shapes, for ncu:
   ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 4 -c 4 -o /tmp/wg python scripts/prof_wgrad_nt.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import ops
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
m = 25120
for rep in range(2):                                   # first pass = warm-up (skipped by -s 4)
    for n_out, k_in in ((4096, 512), (1536, 512), (512, 2048), (512, 512)):
        dy = torch.randn((m, n_out), device=dev, generator=g).bfloat16()
        x = torch.randn((m, k_in), device=dev, generator=g).bfloat16()
        dw = torch.zeros((n_out, k_in), device=dev)
        ops.linear_wgrad_nt_(dw, dy, x)
torch.cuda.synchronize()
print("done")
