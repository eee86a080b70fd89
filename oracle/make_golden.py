"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

The reference modules are imported as they lie (``PYTHONPATH=/root/reference``); weights and
inputs come from the seeded numpy generators in ``mintime_b200.synth`` so that tests can rebuild
the exact same tensors anywhere and compare against the stored reference OUTPUTS.

Stored per case (kept small -- strided samples for the big tensors):
  * extractor: for every stage (stem, block0..15, head) mean, mean-abs and 2048 samples
  * transformer: embeddings sample, per-layer residual-stream sample, logits, both attention maps
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("MINTIME_REFERENCE", "/root/reference")

import mintime_b200 as mt                                    # noqa: E402
from mintime_b200 import synth                               # noqa: E402
from mintime_b200.spec import default_tsf_config             # noqa: E402


def sample(t: torch.Tensor, k: int = 2048) -> np.ndarray:
    flat = t.detach().reshape(-1)
    idx = sample_index(flat.numel(), k)
    return flat[idx].numpy().astype(np.float32)


def sample_index(numel: int, k: int = 2048) -> torch.Tensor:
    if numel <= k:
        return torch.arange(numel)
    return (torch.arange(k, dtype=torch.float64) * (numel - 1) / (k - 1)).round().long()


def load_reference():
    sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.efficientnet.efficientnet_pytorch import EfficientNet
        from models.size_invariant_timesformer import SizeInvariantTimeSformer
    return EfficientNet, SizeInvariantTimeSformer


CASES = {
    # name: (B, f, identities, pad_tail)
    "cfg1_b1_f8_id1": (1, 8, [1], False),          # BASELINE.json configs[0]
    "b2_f16_id2": (2, 16, [2], True),              # configs[2] shape, small batch
    "b4_f16_mixed": (4, 16, [1, 2, 3, 4], True),   # configs[4] identity mix, small batch
    "b2_f8_id2": (2, 8, [2, 1], True),
}


# bf16 end-to-end fixtures: the same cases with the CONDITIONED extractor weights (synth.make_effnet_state_dict(...,
# conditioned=True)), plus 4 clips of the benchmark batch (BASELINE.json configs[1]: B = 32, f = 16, 1 identity,
# seed 1234 = rank 0 of bench.py) so that bench.py and tests can check the bench-shape run against the reference
BENCH_CLIPS = [0, 9, 18, 31]


def bench_clips():
    B, f = 32, 16
    cfg = default_tsf_config(num_frames=f, channels=1280)
    meta = synth.make_batch_meta(B, f, [1], seed=1234)
    frames = synth.make_frames(B, f, seed=1234, mask=meta["mask"], dtype=torch.uint8)
    idx = torch.tensor(BENCH_CLIPS)
    return cfg, {k: v[idx] for k, v in meta.items()}, frames[idx].float()


def main(conditioned: bool = False):
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    EfficientNet, SizeInvariantTimeSformer = load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    esd = synth.make_effnet_state_dict(1234, conditioned=conditioned)
    ext = EfficientNet.from_name("efficientnet-b0")
    missing = ext.load_state_dict(esd, strict=True)
    ext.eval()
    print("extractor state_dict keys:", len(esd), missing, "conditioned" if conditioned else "")
    cases = dict(CASES)
    if conditioned:
        cases = {"cond_" + k: v for k, v in CASES.items()}
        cases["cond_bench_b32_clips"] = None

    for name, spec in cases.items():
        if spec is None:
            cfg, meta, frames = bench_clips()
            B, f = frames.shape[:2]
            tsd = synth.make_tsf_state_dict(cfg, 4321)
            model = SizeInvariantTimeSformer(config=cfg, require_attention=True)
            model.load_state_dict(tsd, strict=True)
            model.eval()
            with torch.no_grad():
                x = frames.permute(0, 1, 4, 2, 3).reshape(B * f, 3, 224, 224)
                feats = ext(x)
                logits, (space_attn, time_attn) = model(feats.view(B, f, *feats.shape[1:]), mask=meta["mask"],
                                                        size_embedding=meta["size_embedding"],
                                                        identities_mask=meta["identities_mask"], positions=meta["positions"])
            g = {"clips": np.asarray(BENCH_CLIPS), "tsf.logits": logits.numpy(), "tsf.space_attn": space_attn.numpy(),
                 "tsf.time_attn": time_attn.numpy(), "ext.head.sample": sample(feats),
                 "ext.head.absmean": np.float32(feats.abs().mean().item())}
            path = os.path.join(out_dir, name + ".npz")
            np.savez_compressed(path, **g)
            print(name, "logits", logits.flatten().tolist(), "->", path, os.path.getsize(path) // 1024, "KiB")
            continue
        B, f, ids, pad = spec
        cfg = default_tsf_config(num_frames=f, channels=1280)
        tsd = synth.make_tsf_state_dict(cfg, 4321)
        model = SizeInvariantTimeSformer(config=cfg, require_attention=True)
        model.load_state_dict(tsd, strict=True)
        model.eval()
        meta = synth.make_batch_meta(B, f, ids, seed=1234, pad_tail=pad)
        frames = synth.make_frames(B, f, seed=1234, mask=meta["mask"])
        g = {}
        with torch.no_grad():
            # train.py:341 / predict.py:402
            x = frames.permute(0, 1, 4, 2, 3).reshape(B * f, 3, 224, 224)
            taps = {}
            hooks = []
            hooks.append(ext._bn0.register_forward_hook(lambda m, i, o: taps.__setitem__("stem_prebn", o)))
            for i, blk in enumerate(ext._blocks):
                hooks.append(blk.register_forward_hook(lambda m, i_, o, i=i: taps.__setitem__(f"block{i}", o)))
            feats = ext(x)
            for h in hooks:
                h.remove()
            taps["stem"] = taps.pop("stem_prebn")
            taps["stem"] = taps["stem"] * torch.sigmoid(taps["stem"])
            taps["head"] = feats
            for k, v in taps.items():
                g[f"ext.{k}.sample"] = sample(v)
                g[f"ext.{k}.mean"] = np.float32(v.mean().item())
                g[f"ext.{k}.absmean"] = np.float32(v.abs().mean().item())
                g[f"ext.{k}.shape"] = np.asarray(v.shape)
            feats5 = feats.view(B, f, *feats.shape[1:])                       # train.py:354
            ltaps = {}
            hooks = []
            for l, layer in enumerate(model.layers):
                hooks.append(layer[2].register_forward_hook(
                    lambda m, i, o, l=l: ltaps.__setitem__(f"layer{l}.ff", o + i[0])))
            logits, (space_attn, time_attn) = model(
                feats5, mask=meta["mask"], size_embedding=meta["size_embedding"],
                identities_mask=meta["identities_mask"], positions=meta["positions"])
            for h in hooks:
                h.remove()
            for k, v in ltaps.items():
                g[f"tsf.{k}.sample"] = sample(v)
            g["tsf.logits"] = logits.numpy()
            g["tsf.space_attn"] = space_attn.numpy()
            g["tsf.time_attn"] = time_attn.numpy()
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **g)
        print(name, "logits", logits.flatten().tolist(), "feat std", float(feats.std()),
              "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    if "--conditioned" in sys.argv:
        main(conditioned=True)
    else:
        main()
