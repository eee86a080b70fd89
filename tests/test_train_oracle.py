"""CPU: the oracle's gradients (torch autograd through oracle/mintime_oracle.py) against the gradients of the UNMODIFIED
reference stored in tests/golden/grads_*.npz (oracle/make_golden_grads.py), and the host-side training helpers."""
import numpy as np
import pytest
import torch

import mintime_b200  # noqa: F401
from mintime_b200 import training, weights
from oracle import mintime_oracle as orc
from helpers import GRAD_CASES, grad_case_inputs, load_golden, sample


@pytest.mark.parametrize("case", list(GRAD_CASES))
def test_oracle_autograd_matches_reference_gradients(case):
    cfg, tsd, meta, feats, labels, pw = grad_case_inputs(case)
    gold = load_golden("grads_" + case)
    sd = {k: v.clone().requires_grad_(True) for k, v in tsd.items()}
    logits, _ = orc.tsf_forward(sd, cfg, feats, meta["mask"], meta["identities_mask"], meta["size_embedding"],
                                meta["positions"])
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([pw]))(logits, labels)
    loss.backward()
    assert np.allclose(logits.detach().numpy(), gold["logits"], rtol=1e-4, atol=1e-5)
    assert abs(loss.item() - float(gold["loss"])) <= 1e-6
    checked = 0
    for k, v in sd.items():
        if f"grad.{k}.sample" not in gold:
            continue
        ref = gold[f"grad.{k}.sample"]
        got = sample(v.grad, 512)
        assert np.linalg.norm(got - ref) <= 1e-4 * np.linalg.norm(ref) + 1e-10, k
        assert abs(float(v.grad.double().norm()) - float(gold[f"grad.{k}.norm"])) <= 1e-4 * float(gold[f"grad.{k}.norm"]) + 1e-10
        checked += 1
    assert checked == len([k for k in gold if k.endswith(".norm")])


def test_geglu_uninterleave_inverts_interleave():
    t = torch.arange(256 * 3, dtype=torch.float32).view(256, 3)
    assert torch.equal(training._uninterleave(weights.geglu_interleave(t)), t)
    b = torch.arange(128, dtype=torch.float32)
    assert torch.equal(training._uninterleave(weights.geglu_interleave(b)), b)


def test_layer_gradient_layout_covers_every_layer_parameter():
    from mintime_b200 import SizeInvariantTimeSformer
    from mintime_b200.spec import default_tsf_config
    cfg = default_tsf_config(num_frames=8)
    cfg["model"]["depth"] = 1
    model = SizeInvariantTimeSformer(config=cfg)
    dim, inner = 512, 512
    names = {f"layers.0.{j}.{k}": s for j, lay in ((0, training._attn_layout(dim, inner)), (1, training._attn_layout(dim, inner)),
                                                  (2, training._ff_layout(dim))) for k, s in lay}
    got = {k: tuple(p.shape) for k, p in model.named_parameters() if k.startswith("layers.0.")}
    assert names == got


def test_train_pack_layouts_on_cpu():
    """TrainPack (device-side re-packing of the live parameters, redone after every optimizer step): transposed copies
    for the dgrad GEMMs, q rows pre-scaled by dim_head^-0.5 (:114), GEGLU rows interleaved like the inference pack."""
    from mintime_b200 import SizeInvariantTimeSformer, synth
    from mintime_b200.spec import default_tsf_config
    cfg = default_tsf_config(num_frames=8)
    cfg["model"]["depth"] = 2
    model = SizeInvariantTimeSformer(config=cfg, precision="fp32")
    sd = synth.make_tsf_state_dict(cfg, 5)
    model.load_state_dict(sd)
    pk = training.TrainPack(model, "fp32", "cpu")
    ref = weights.pack_tsf(sd, cfg, "fp32", "cpu")
    assert len(pk.layers) == 2
    for l, L in enumerate(pk.layers):
        for j, name in ((0, "time"), (1, "space")):
            wq = sd[f"layers.{l}.{j}.fn.to_qkv.weight"].clone()
            wq[:512] *= 0.125
            assert torch.equal(L[name + ".wqkv"], wq) and torch.equal(L[name + ".wqkv_t"], wq.t().contiguous())
            assert L[name + ".wqkv_t"].is_contiguous() and L[name + ".wqkv_t"].shape == (512, 1536)
            assert torch.equal(L[name + ".wo_t"], sd[f"layers.{l}.{j}.fn.to_out.0.weight"].t())
        w1 = weights.geglu_interleave(sd[f"layers.{l}.2.fn.net.0.weight"])
        assert torch.equal(L["ff.w1"], w1) and torch.equal(L["ff.w1_t"], w1.t().contiguous())
        assert torch.equal(L["ff.b1"], weights.geglu_interleave(sd[f"layers.{l}.2.fn.net.0.bias"]))
        assert torch.equal(L["ff.w2_t"], sd[f"layers.{l}.2.fn.net.3.weight"].t())
    # the same bytes as the load-time pack of the inference path (so both paths read identical weights)
    packed_ref = {t.data_ptr(): t for t in ref.keep}
    assert torch.equal(pk.layers[0]["ff.w1"], packed_ref[ref.struct.ff[0].w1])
    assert torch.equal(pk.layers[1]["time.wqkv"], packed_ref[ref.struct.time_attn[1].w_qkv])
    # fp32 parameters are read in place (no copy): an optimizer step is visible without re-packing them
    assert pk.out_w.data_ptr() == model.to_out[1].weight.data_ptr()


@pytest.mark.parametrize("tag,rate", [("nodrop", 0.0), ("drop", 0.2)])
def test_oracle_extractor_train_mode_matches_reference(tag, rate):
    """oracle.effnet_b0_forward_train (BatchNorm batch statistics + running-stat update, drop-connect with the reference's
    RNG draws, autograd gradients) against the unmodified reference in .train() -- the checker for the extractor's
    backward kernels (not built in the product yet, which raises instead)."""
    from helpers import EXTRACTOR_TRAIN_KEYS, extractor_train_inputs
    gold = load_golden("extractor_train")
    esd, x, probe = extractor_train_inputs()
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in esd.items()}
    new_stats = {}
    torch.manual_seed(7)
    out = orc.effnet_b0_forward_train(sd, x, drop_connect_rate=rate, new_stats=new_stats)
    (out * probe).sum().backward()
    scale = float(gold[f"{tag}.out_absmean"])
    assert np.abs(sample(out) - gold[f"{tag}.out"]).max() <= 2e-4 * max(scale, 1.0)
    for k in EXTRACTOR_TRAIN_KEYS:
        ref = gold[f"{tag}.grad.{k}"]
        got = sample(sd[k].grad, 512)
        assert np.linalg.norm(got - ref) <= 2e-3 * np.linalg.norm(ref) + 1e-8, k
    for k in ("_bn0", "_blocks.0._bn1", "_blocks.5._bn0", "_blocks.15._bn2", "_bn1"):
        assert np.allclose(new_stats[k + ".running_mean"].numpy(), gold[f"{tag}.{k}.running_mean"], rtol=1e-4, atol=1e-5), k
        assert np.allclose(new_stats[k + ".running_var"].numpy(), gold[f"{tag}.{k}.running_var"], rtol=1e-4, atol=1e-5), k
