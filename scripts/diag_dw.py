import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import mintime_b200
from mintime_b200 import ops
dev = "cuda:0"
def run(h, c, k, n=1, seed=0):
    g = np.random.default_rng(seed)
    x = torch.from_numpy(g.standard_normal((n, h, h, c)).astype(np.float32)).bfloat16()
    w = torch.from_numpy((g.standard_normal((c, 1, k, k)) / k).astype(np.float32))
    wb = w.bfloat16().float()
    shift = torch.zeros(c)
    y = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (k // 2,) * 4), wb, None, 1, 0, 1, c)
    ref = (y * torch.sigmoid(y)).permute(0, 2, 3, 1)
    taps = w[:, 0].permute(1, 2, 0).reshape(k * k, c).contiguous()
    out, pool = ops.dwconv(x.to(dev), taps.to(dev), shift.to(dev), k, 1, precision="bf16")
    torch.cuda.synchronize()
    o = out.float().cpu()
    err = (o - ref).abs()
    rel = float((o - ref).norm() / ref.norm())
    bad = err > 0.05
    print(f"h={h} c={c} k={k} BO={os.environ.get('MINTIME_B200_DW_BO','1')}: rel {rel:.4f} bad frac {bad.float().mean():.4f} nan {int(torch.isnan(o).sum())}")
    if bad.any():
        bp = bad.any(-1)[0]          # (h, h) pixels with any bad channel
        print("  bad pixel map (rows):")
        for r in range(min(h, 16)):
            print("   ", "".join("X" if bp[r, cidx] else "." for cidx in range(min(h, 64))))
        bc = bad[0].any(0).any(0)
        print("  bad channels:", [i for i in range(c) if bc[i]][:40])
for (h, c, k) in [(7, 64, 3), (14, 64, 3), (14, 32, 5), (20, 16, 3)]:
    run(h, c, k)
