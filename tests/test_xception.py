"""Xception extractor (reference models/xception.py): oracle vs fixtures of the unmodified reference (CPU), the
module's interface (CPU), and the B200 path through the C ABI against both (GPU)."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, XCEPTION_STAGES, load_golden, rel_err, sample, xception_inputs, xception_tsf_inputs
from oracle import mintime_oracle as orc
from oracle import xception_oracle as xo

import mintime_b200
from mintime_b200 import synth
from mintime_b200.xception import XCEPTION_BLOCKS, Xception, sep_channels, xception

DEV = "cuda:0"


# ------------------------------------------------------------------------------------------ CPU
def test_xception_state_dict_matches_reference_keys_and_shapes():
    with open(os.path.join(GOLDEN_DIR, "xception_keys.json")) as fh:
        ref = json.load(fh)
    own = Xception(num_classes=1).state_dict()
    assert list(own.keys()) == list(ref.keys())                    # same names in the same order (276)
    assert all(list(own[k].shape) == ref[k] for k in ref)
    sd = synth.make_xception_state_dict(2468)
    assert set(sd) == set(ref) and all(list(sd[k].shape) == ref[k] for k in ref)
    assert len(list(sum((sep_channels(c, o, r, g) for c, o, r, _, _, g in XCEPTION_BLOCKS), []))) == 32


def test_xception_oracle_matches_reference_fixture():
    sd, x = xception_inputs()
    g = load_golden("xception_b2")
    with torch.no_grad():
        out, feats = xo.xception_features(sd, x, stages=True)
    assert tuple(out.shape) == (2, 2048, 7, 7)
    by_name = dict(feats)
    for name in XCEPTION_STAGES:
        t = by_name[name]
        scale = float(g[f"stage.{name}.mean_abs"])
        assert abs(float(t.abs().mean()) - scale) <= 1e-4 * scale
        assert np.abs(sample(t) - g["stage." + name]).max() <= 2e-4 * scale, name
    scale = float(g["out_mean_abs"])
    assert np.abs(sample(out, 8192) - g["out"]).max() <= 2e-4 * scale


def test_xception_factory_copies_matching_entries_and_strips_module_prefix(tmp_path):
    sd = synth.make_xception_state_dict(2468)
    ck = {"module." + k: v for k, v in sd.items()}
    ck["module.fc.weight"] = torch.zeros((5, 2048))               # wrong shape: reported and skipped (xception.py:222-229)
    ck["something.else"] = torch.zeros(3)
    path = os.path.join(tmp_path, "ckpt.pth")
    torch.save(ck, path)
    m = xception(num_classes=1, pretrain_path=path)
    own = m.state_dict()
    assert torch.equal(own["block7.rep.4.pointwise.weight"], sd["block7.rep.4.pointwise.weight"])
    assert torch.equal(own["bn4.running_var"], sd["bn4.running_var"])
    assert not torch.equal(own["fc.weight"], torch.zeros((1, 2048)))


def test_xception_has_no_cpu_fallback():
    m = Xception(num_classes=1).eval()
    with pytest.raises(mintime_b200.lib.MintimeError):
        m(torch.zeros((1, 3, 224, 224)))
    m.train()
    with pytest.raises(mintime_b200.lib.MintimeError):
        m(torch.zeros((1, 3, 224, 224)))


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("prec,u8", [("fp32", False), ("bf16", False), ("bf16", True)])
def test_xception_matches_reference_fixture(prec, u8):
    sd, x = xception_inputs()
    g = load_golden("xception_b2")
    m = Xception(num_classes=1, precision=prec)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    xin = x.to(DEV)
    xin = xin.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)          # NHWC memory, as the callers build it
    if u8:
        xin = xin.to(torch.uint8)
    out = m(xin)
    torch.cuda.synchronize()
    assert tuple(out.shape) == (2, 2048, 7, 7) and out.dtype == (torch.float32 if prec == "fp32" else torch.bfloat16)
    got = sample(out.float().cpu().contiguous(), 8192)
    # fp32: the reference's own result to 1e-3 of scale; bf16: 36 bf16 convolutions deep, stated bar 3e-2 rel-L2
    if prec == "fp32":
        assert np.abs(got - g["out"]).max() <= 1e-3 * float(g["out_mean_abs"])
    else:
        assert rel_err(got, g["out"]) <= 3e-2


@pytest.mark.gpu
def test_xception_true_nchw_input_and_batch_independence():
    sd, x = xception_inputs()
    m = Xception(num_classes=1, precision="bf16")
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    a = m(x.to(DEV))                                                           # true NCHW memory
    b = m(x.to(DEV).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
    c = m(x[1:].to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert torch.equal(a[1:], c)                                               # a face does not depend on its batch


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_xception_into_timesformer_matches_reference(prec):
    """the `--extractor_model 1` path: Xception features (channels = 2048) -> SizeInvariantTimeSformer"""
    cfg, xsd, tsd, meta, frames = xception_tsf_inputs()
    g = load_golden("xception_tsf_b1_f8")
    ext = Xception(num_classes=1, precision=prec)
    ext.load_state_dict(xsd)
    ext = ext.to(DEV).eval()
    model = mintime_b200.SizeInvariantTimeSformer(config=cfg, require_attention=True, precision=prec)
    model.load_state_dict(tsd)
    model = model.to(DEV).eval()
    B, f = frames.shape[:2]
    videos = frames.to(DEV)
    with torch.no_grad():
        feats = ext(videos.view(B * f, 224, 224, 3).permute(0, 3, 1, 2))
        logits, (space, time) = model(feats.view(B, f, 2048, 7, 7), mask=meta["mask"].to(DEV),
                                      size_embedding=meta["size_embedding"].to(DEV),
                                      identities_mask=meta["identities_mask"].to(DEV), positions=meta["positions"].to(DEV))
    torch.cuda.synchronize()
    if prec == "fp32":
        np.testing.assert_allclose(logits.cpu().numpy(), g["logits"], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(space.cpu().numpy(), g["space_attn"], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(time.cpu().numpy(), g["time_attn"], rtol=1e-3, atol=1e-4)
    else:
        assert np.abs(logits.float().cpu().numpy() - g["logits"]).max() <= 2e-2
        assert rel_err(space.float().cpu(), g["space_attn"]) <= 2e-2
        assert rel_err(time.float().cpu(), g["time_attn"]) <= 2e-2


def test_pack_xception_folds_batchnorm_and_orders_im2col_columns():
    """host logic (CPU): the packed conv1 rows -- BatchNorm folded, columns (ky, kx, ci), 27 padded to 32 -- reproduce
    bn1(conv1(x)) of the reference on an im2col matrix built the way xc_im2col3x3_kernel builds it; the first separable unit's
    pointwise rows reproduce bn(pointwise(.))"""
    import torch.nn.functional as F
    from mintime_b200 import weights
    sd = synth.make_xception_state_dict(2468)
    pk = weights.pack_xception(sd, "fp32", "cpu")
    w1, s1 = pk.keep[0], pk.keep[1]                                  # conv1: [32][32], shift [32]
    assert tuple(w1.shape) == (32, 32) and float(w1[:, 27:].abs().max()) == 0.0
    x = torch.rand((1, 3, 9, 9), generator=torch.Generator().manual_seed(1)) * 255.0
    ref = xo._bn(sd, "bn1", F.conv2d(x, sd["conv1.weight"], None, 2, 0))           # (1,32,4,4)
    cols = F.unfold(x, 3, stride=2)                                  # (1, ci*9, 16), rows ordered (ci, ky, kx)
    cols = cols.view(1, 3, 9, 16).permute(0, 3, 2, 1).reshape(16, 27)  # -> [pixel][(ky,kx), ci]
    cols = torch.cat([cols, torch.zeros(16, 5)], dim=1)
    got = (cols @ w1.t() + s1).t().reshape(1, 32, 4, 4)
    assert rel_err(got, ref) <= 1e-5
    # unit 0 = block1.rep.0: keep[4] depthwise taps [9][64], keep[5] folded pointwise [128][64], keep[6] shift
    dw, pw, sh = pk.keep[4], pk.keep[5], pk.keep[6]
    assert tuple(dw.shape) == (9, 64) and tuple(pw.shape) == (128, 64)
    a = torch.randn((1, 64, 5, 5), generator=torch.Generator().manual_seed(2))
    ref = xo._bn(sd, "block1.rep.1", xo._sep(sd, "block1.rep.0", a))
    mid = F.conv2d(a, dw.t().reshape(64, 1, 3, 3), None, 1, 1, 1, 64)
    got = F.conv2d(mid, pw.reshape(128, 64, 1, 1)) + sh.view(1, -1, 1, 1)
    assert rel_err(got, ref) <= 1e-5
