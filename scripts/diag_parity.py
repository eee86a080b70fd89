"""Print parity errors of each stage on the GPU (diagnostic; numbers quoted in DESIGN.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import case_inputs, load_golden, rel_err, sample
from oracle import mintime_oracle as orc
import mintime_b200
from mintime_b200 import synth, spec, weights, _lib
DEV = "cuda:0"
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1_b1_f8_id1"
cfg, esd, tsd, meta, frames = case_inputs(name)
B, f = frames.shape[:2]
with torch.no_grad():
    x = frames.permute(0, 1, 4, 2, 3).reshape(B * f, 3, 224, 224)
    feats_ref = orc.effnet_b0_forward(esd, x)
    lref, (sref, tref) = orc.tsf_forward(tsd, cfg, feats_ref.view(B, f, 1280, 7, 7), meta["mask"], meta["identities_mask"], meta["size_embedding"], meta["positions"])
for prec in ("fp32", "bf16"):
    ext = mintime_b200.EfficientNet.from_name("efficientnet-b0", precision=prec); ext.load_state_dict(esd); ext = ext.to(DEV).eval()
    model = mintime_b200.SizeInvariantTimeSformer(config=cfg, require_attention=True, precision=prec); model.load_state_dict(tsd); model = model.to(DEV).eval()
    with torch.no_grad():
        vid = frames.to(DEV).view(B * f, 224, 224, 3).permute(0, 3, 1, 2)
        feats = ext(vid)
        feats2 = ext(vid)
        print(prec, "extractor rel-L2 vs oracle", rel_err(feats.float().cpu(), feats_ref), "max-abs", (feats.float().cpu() - feats_ref).abs().max().item(),
              "ref absmax", feats_ref.abs().max().item(), "run-to-run max diff", (feats.float() - feats2.float()).abs().max().item())
        kw = dict(mask=meta["mask"].to(DEV), size_embedding=meta["size_embedding"], identities_mask=meta["identities_mask"].to(DEV), positions=meta["positions"].to(DEV))
        l1, (s1, t1) = model(feats.reshape(B, f, 1280, 7, 7), **kw)
        l2, (s2, t2) = model(feats_ref.to(DEV).view(B, f, 1280, 7, 7), **kw)
    print(prec, " full path : dlogit", (l1.cpu() - lref).abs().max().item(), "space rel", rel_err(s1.cpu(), sref), "time rel", rel_err(t1.cpu(), tref))
    print(prec, " tsf only  : dlogit", (l2.cpu() - lref).abs().max().item(), "space rel", rel_err(s2.cpu(), sref), "time rel", rel_err(t2.cpu(), tref))
# patch embed fp32
prec = "fp32"
f2, B2 = 8, 3
cfg2 = spec.default_tsf_config(num_frames=f2); sd = synth.make_tsf_state_dict(cfg2, 4321); m2 = synth.make_batch_meta(B2, f2, [2, 1, 2], seed=3)
g = np.random.default_rng(9); ft = torch.from_numpy((g.standard_normal((B2, f2, 1280, 7, 7)) * 20).astype(np.float32))
ref = orc.tsf_embed(sd, cfg2, ft, m2["size_embedding"], m2["positions"])
pk = weights.pack_tsf(sd, cfg2, prec, DEV); c = weights.tsf_cfg_struct(cfg2)
tok = ft.permute(0, 1, 3, 4, 2).contiguous().to(DEV); xo = torch.empty((B2, 1 + f2 * 49, 512), dtype=torch.float32, device=DEV)
rc = _lib.load().mt_patch_embed_fwd(0, pk.struct, c, tok.data_ptr(), m2["size_embedding"].to(DEV).data_ptr(), m2["positions"].to(DEV).data_ptr(), xo.data_ptr(), B2, _lib.stream_ptr())
torch.cuda.synchronize()
d = (xo.cpu() - ref)
print("patch_embed fp32 rel", rel_err(xo.cpu(), ref), "max abs", d.abs().max().item(), "ref absmax", ref.abs().max().item(), "cls row err", d[:, 0].abs().max().item(), "argmax", np.unravel_index(d.abs().argmax().item(), d.shape))
