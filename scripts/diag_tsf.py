import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import mintime_oracle as orc
import mintime_b200
from mintime_b200 import synth, spec, weights, _lib
from mintime_b200.size_invariant_timesformer import SizeInvariantTimeSformer
from helpers import case_inputs
DEV = "cuda:0"
cfg0, esd, tsd0, meta0, frames0 = case_inputs("cfg1_b1_f8_id1")
f = 8
def run(tag, cfg, tsd, feats, meta, depth=None):
    if depth is not None:
        import copy
        cfg = copy.deepcopy(cfg); cfg["model"]["depth"] = depth
        tsd = {k: v for k, v in tsd.items() if not k.startswith("layers.") or int(k.split(".")[1]) < depth}
    with torch.no_grad():
        rl, (rs, rt) = orc.tsf_forward(tsd, cfg, feats, meta["mask"], meta["identities_mask"], meta["size_embedding"], meta["positions"])
    model = SizeInvariantTimeSformer(config=cfg, require_attention=True, precision="fp32"); model.load_state_dict(tsd); model = model.to(DEV).eval()
    with torch.no_grad():
        l, (s, t) = model(feats.to(DEV), mask=meta["mask"].to(DEV), size_embedding=meta["size_embedding"], identities_mask=meta["identities_mask"].to(DEV), positions=meta["positions"].to(DEV))
    print(f"{tag:34s} dlogit {(l.cpu()-rl).abs().max().item():.2e}  space rel {((s.cpu()-rs).norm()/rs.norm()).item():.2e}  time rel {((t.cpu()-rt).norm()/rt.norm()).item():.2e}")
g = torch.Generator().manual_seed(8)
rand_feats = torch.nn.functional.silu(torch.randn((1, f, 1280, 7, 7), generator=g)) * 20.0
with torch.no_grad():
    x = frames0.permute(0, 1, 4, 2, 3).reshape(f, 3, 224, 224)
    ofeats = orc.effnet_b0_forward(esd, x).view(1, f, 1280, 7, 7)
tsd99 = synth.make_tsf_state_dict(cfg0, 99)
meta2 = synth.make_batch_meta(1, f, [2], seed=8, pad_tail=False)
run("fixture tsd, oracle feats, meta0", cfg0, tsd0, ofeats, meta0)
run("fixture tsd, RANDOM feats, meta0", cfg0, tsd0, rand_feats, meta0)
run("tsd99, oracle feats, meta0", cfg0, tsd99, ofeats, meta0)
run("fixture tsd, oracle feats, meta2", cfg0, tsd0, ofeats, meta2)
run("fixture tsd, RANDOM feats, depth1", cfg0, tsd0, rand_feats, meta0, depth=1)
run("fixture tsd, RANDOM/20 feats", cfg0, tsd0, rand_feats / 20.0, meta0)
