"""Generate tests/golden/grads_*.npz: gradients of the UNMODIFIED reference SizeInvariantTimeSformer (torch autograd on
the module imported from /root/reference) for one training step body of train.py:355-377 with a frozen extractor:

    y_pred = model(features, mask=, size_embedding=, identities_mask=, positions=)
    loss = BCEWithLogitsLoss(pos_weight)(y_pred, labels);  loss.backward()

Run in the build container only:   python oracle/make_golden_grads.py
Inputs are rebuilt from seeds by tests/helpers.py::grad_case_inputs; stored per parameter: its gradient's norm and a
strided sample (512 values; the embedding tables also a 2048-value sample of their reachable rows), plus the logits
and the loss.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("MINTIME_REFERENCE", "/root/reference")

from helpers import GRAD_CASES, grad_case_inputs, sample   # noqa: E402


def main():
    sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.size_invariant_timesformer import SizeInvariantTimeSformer
    torch.set_num_threads(os.cpu_count())
    for name in GRAD_CASES:
        cfg, tsd, meta, feats, labels, pos_weight = grad_case_inputs(name)
        model = SizeInvariantTimeSformer(config=cfg)
        model.load_state_dict(tsd, strict=True)
        model.train()                                                        # train.py:315
        y = model(feats, mask=meta["mask"], size_embedding=meta["size_embedding"],
                  identities_mask=meta["identities_mask"], positions=meta["positions"])
        loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([pos_weight]))(y, labels)   # train.py:186-187
        loss.backward()
        g = {"logits": y.detach().numpy(), "loss": np.float32(loss.item())}
        for k, p in model.named_parameters():
            g[f"grad.{k}.norm"] = np.float64(p.grad.double().norm().item())
            g[f"grad.{k}.sample"] = sample(p.grad, 512)
            if k in ("pos_emb.weight", "size_emb.weight"):
                # only the first rows of the (oversized, :173-180) tables are reachable: sample those densely
                rows = 1 + cfg["model"]["num-frames"] * cfg["model"]["num-patches"] if k.startswith("pos") else 21
                g[f"grad.{k}.head_sample"] = sample(p.grad[:rows], 2048)
        path = os.path.join(ROOT, "tests", "golden", f"grads_{name}.npz")
        np.savez_compressed(path, **g)
        print(name, "loss", loss.item(), "logits", y.flatten().tolist(), "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
