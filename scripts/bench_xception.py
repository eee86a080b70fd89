"""Throughput of the `--extractor_model 1` path (Xception -> SizeInvariantTimeSformer, channels = 2048) at the bench shape:
B = 32 clips x 16 frames per step, bf16, one CUDA-graph replay per step, CUDA events.  Prints one JSON line.
    python scripts/bench_xception.py [--batch 32] [--steps 10]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mintime_b200
from mintime_b200 import synth
from mintime_b200.graphed import GraphedHotPath
from mintime_b200.spec import default_tsf_config

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
dev = "cuda:0"
B, f = args.batch, args.frames
cfg = default_tsf_config(num_frames=f, channels=2048)
ext = mintime_b200.Xception(num_classes=1, precision="bf16")
ext.load_state_dict(synth.make_xception_state_dict(2468))
ext = ext.to(dev).eval()
model = mintime_b200.SizeInvariantTimeSformer(config=cfg, require_attention=False, precision="bf16")
model.load_state_dict(synth.make_tsf_state_dict(cfg, 4321))
model = model.to(dev).eval()
meta = synth.make_batch_meta(B, f, [1], seed=1234)
frames = synth.make_frames(B, f, seed=1234, mask=meta["mask"], dtype=torch.uint8)
hot = GraphedHotPath(ext, model, batch=B, num_frames=f)
hot(frames.to(dev), meta["mask"].to(dev), meta["identities_mask"].to(dev), meta["size_embedding"].to(dev), meta["positions"].to(dev))
for _ in range(args.warmup):
    hot.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    out = hot.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
print(json.dumps({"metric": "videos_per_sec_16f_224px", "value": B / ms * 1e3, "unit": "videos/s", "ms_per_step": ms, "n_gpus": 1,
                  "steps": args.steps, "warmup": args.warmup, "dtype": "bf16", "data": "synthetic",
                  "config": {"workload": f"batch={B} {f}-frame clips, Xception (eval) -> SizeInvariantTimeSformer, channels=2048",
                             "launch": "one CUDA-graph replay per step"},
                  "gpu_launches": hot.kernels_per_replay * args.steps, "logit0": float(out.flatten()[0])}))

if os.environ.get("KERNELS"):
    # per-kernel CUDA-event times of one eager extractor forward (the library's own scopes)
    from mintime_b200 import _lib
    lib = _lib.load()
    x = frames.to(dev).view(B * f, 224, 224, 3).permute(0, 3, 1, 2)
    with torch.no_grad():
        ext(x); torch.cuda.synchronize()
        lib.mt_prof_reset(); lib.mt_prof_enable(1)
        ext(x); torch.cuda.synchronize()
        lib.mt_prof_enable(0)
    rows = sorted(_lib.profile_collect(), key=lambda r: -r[1])
    tot = sum(r[1] for r in rows)
    print(f"# extractor forward, {B * f} faces: {tot:.2f} ms in {sum(r[4] for r in rows)} launches", file=sys.stderr)
    for name, ms, flops, byts, cnt in rows[:24]:
        print(f"# {name:34s} {ms:7.3f} ms /{cnt:3d}  {flops / ms / 1e9:7.1f} TF/s  {byts / ms / 1e6:7.0f} GB/s", file=sys.stderr)
