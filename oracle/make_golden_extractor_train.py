"""tests/golden/extractor_train.npz: the UNMODIFIED reference EfficientNet-B0 in train mode (train.py:155-170, unfrozen
extractor) on 4 seeded faces -- output sample, updated BatchNorm running stats, and the gradients of sum(out * probe)
w.r.t. a set of parameters -- once with drop-connect off and once with the stock rate 0.2 under torch.manual_seed(7).
Pins oracle.effnet_b0_forward_train (tests/test_train_oracle.py).   python oracle/make_golden_extractor_train.py"""
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

from helpers import EXTRACTOR_TRAIN_KEYS, extractor_train_inputs, sample   # noqa: E402


def main():
    sys.path.insert(0, REF)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.efficientnet.efficientnet_pytorch import EfficientNet
    torch.set_num_threads(os.cpu_count())
    esd, x, probe = extractor_train_inputs()
    g = {}
    for tag, rate in (("nodrop", 0.0), ("drop", 0.2)):
        ext = EfficientNet.from_name("efficientnet-b0")
        ext.load_state_dict(esd, strict=True)
        ext._global_params = ext._global_params._replace(drop_connect_rate=rate)
        ext.train()
        torch.manual_seed(7)
        out = ext(x)
        (out * probe).sum().backward()
        g[f"{tag}.out"] = sample(out)
        g[f"{tag}.out_absmean"] = np.float32(out.abs().mean().item())
        params = dict(ext.named_parameters())
        bufs = dict(ext.named_buffers())
        for k in EXTRACTOR_TRAIN_KEYS:
            g[f"{tag}.grad.{k}"] = sample(params[k].grad, 512)
            g[f"{tag}.gradnorm.{k}"] = np.float64(params[k].grad.double().norm().item())
        for k in ("_bn0", "_blocks.0._bn1", "_blocks.5._bn0", "_blocks.15._bn2", "_bn1"):
            g[f"{tag}.{k}.running_mean"] = bufs[k + ".running_mean"].numpy().copy()
            g[f"{tag}.{k}.running_var"] = bufs[k + ".running_var"].numpy().copy()
        print(tag, "out absmean", float(out.abs().mean()))
    path = os.path.join(ROOT, "tests", "golden", "extractor_train.npz")
    np.savez_compressed(path, **g)
    print("->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
