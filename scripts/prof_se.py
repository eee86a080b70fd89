"""SE excitation kernel at the B0 layer shapes, 512 images (for ncu / timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import ops
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
def r(*shape): return torch.randn(*shape, device=dev, generator=g)
# (C, SQ, chunks, hw)
cases = [(32, 8, 28, 12544), (96, 4, 16, 3136), (144, 6, 8, 3136), (240, 10, 4, 784), (480, 20, 1, 196), (672, 28, 1, 196), (1152, 48, 1, 49)]
for (c, sq, ch, hw) in cases:
    pool = r(512, ch, c); wr = r(sq, c); br = r(sq); we = r(sq, c); be = r(c)
    for _ in range(3): ops.se_gate(pool, hw, wr, br, we, be)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.se_gate(pool, hw, wr, br, we, be)
    e1.record(); torch.cuda.synchronize()
    print(f"C{c} SQ{sq} chunks{ch}: {e0.elapsed_time(e1)/20*1e3:.1f} us")
