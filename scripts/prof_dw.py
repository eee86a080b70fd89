import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import ops
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
def r(*shape, dtype=torch.bfloat16):
    return torch.randn(*shape, device=dev, generator=g).to(dtype)
cases = [(14, 672, 5, 1, 28), (112, 32, 3, 1, 8), (7, 1152, 5, 1, 48)]
for it in range(2):
    for (h, c, k, s, sq) in cases:
        x = r(512, h, h, c); wt = r(k * k, c, dtype=torch.float32); sh = r(c, dtype=torch.float32)
        wr = r(sq, c, dtype=torch.float32); br = r(sq, dtype=torch.float32); we = r(sq, c, dtype=torch.float32); be = r(c, dtype=torch.float32)
        ops.dwconv_se(x, wt, sh, k, s, wr, br, we, be)
        ops.dwconv(x, wt, sh, k, s)
    torch.cuda.synchronize()
print("done")
