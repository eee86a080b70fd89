"""Backward / training-step parity on the GPU (through the C ABI): every kernel of csrc/train.cu against torch autograd
of the same op, then the whole `loss.backward()` against gradients of the UNMODIFIED reference
(tests/golden/grads_*.npz, oracle/make_golden_grads.py) and the oracle's own autograd."""
import copy

import numpy as np
import pytest
import torch

import mintime_b200  # noqa: F401
from mintime_b200 import SizeInvariantTimeSformer, ops, weights
from oracle import mintime_oracle as orc
from helpers import GRAD_CASES, grad_case_inputs, load_golden, rel_err, sample

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PRECS = ["fp32", "bf16"]


def _T(prec):
    return torch.float32 if prec == "fp32" else torch.bfloat16


def _tol(prec, f32, b16):
    return f32 if prec == "fp32" else b16


# --------------------------------------------------------------------------------------------- operand preparation
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("src_f32", [True, False])
@pytest.mark.parametrize("m,c", [(785 * 2, 512), (100, 64), (393 * 3, 1536)])
def test_grad_prep(prec, src_f32, m, c):
    g = torch.Generator().manual_seed(m + c)
    src = torch.randn((m, c), generator=g)
    T = _T(prec)
    s = src.to(DEV) if src_f32 else src.to(DEV).to(T)
    rm, tr, cs = ops.grad_prep(s, want_rm=True, want_t=True, want_colsum=True, precision=prec)
    mp = (m + 63) // 64 * 64
    assert rm.shape == (m, c) and tr.shape == (c, mp) and cs.shape == (c,)
    want = s.float().to(T)
    assert torch.equal(rm, want)
    assert torch.equal(tr[:, :m], want.t())
    assert float(tr[:, m:].float().abs().sum()) == 0.0
    assert torch.allclose(cs.cpu(), s.float().sum(0).cpu(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("prec", PRECS)
def test_grad_prep_drops_cls_rows(prec):
    B, rows, c = 3, 8 * 49, 512
    src = torch.randn((B * (rows + 1), c), generator=torch.Generator().manual_seed(5)).to(DEV)
    _, tr, cs = ops.grad_prep(src, want_t=True, want_colsum=True, rows_per_batch=rows, m=B * rows, precision=prec)
    kept = src.view(B, rows + 1, c)[:, 1:].reshape(B * rows, c)
    assert torch.equal(tr[:, :B * rows], kept.to(_T(prec)).t())
    assert torch.allclose(cs, kept.sum(0), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n_out,k_in,m", [(512, 512, 785 * 2), (1536, 512, 785 * 4), (4096, 512, 1000), (512, 2048, 785 * 32),
                                          (512, 1280, 784 * 3)])
def test_linear_wgrad_split_k(prec, n_out, k_in, m):
    """dW += dY^T X through the transposed operands (split-K over the SMs on the bf16 path), accumulating in place."""
    g = torch.Generator().manual_seed(n_out + k_in + m)
    T = _T(prec)
    dy = torch.randn((m, n_out), generator=g).to(DEV).to(T)
    x = torch.randn((m, k_in), generator=g).to(DEV).to(T)
    _, dyT, _ = ops.grad_prep(dy, want_t=True, precision=prec)
    _, xT, _ = ops.grad_prep(x, want_t=True, precision=prec)
    dw0 = torch.randn((n_out, k_in), generator=g).to(DEV)
    dw = dw0.clone()
    ops.linear_wgrad_(dw, dyT, xT, prec)
    ref = dw0.double() + dy.double().t() @ x.double()
    assert rel_err(dw, ref) <= _tol(prec, 1e-5, 1e-5)        # bf16 operands are exact products, fp32 accumulation


# --------------------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("rows,dim", [(785 * 2, 512), (37, 128), (5000, 1024)])
def test_layernorm_bwd(prec, rows, dim):
    g = torch.Generator().manual_seed(rows)
    x = (torch.randn((rows, dim), generator=g) * 3 + 0.5).to(DEV)
    gamma = (torch.rand((dim,), generator=g) + 0.5).to(DEV)
    dy = torch.randn((rows, dim), generator=g).to(DEV).to(_T(prec))
    gx0 = torch.randn((rows, dim), generator=g).to(DEV)
    xr = x.double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    br = torch.zeros(dim, dtype=torch.float64, device=DEV, requires_grad=True)
    torch.nn.functional.layer_norm(xr, (dim,), gr, br, 1e-5).backward(dy.double())
    gx = gx0.clone()
    dgm, dbt = ops.layernorm_bwd_(gx, x, gamma, dy, prec)
    assert rel_err(gx - gx0, xr.grad) <= 1e-5
    assert rel_err(dgm, gr.grad) <= 1e-5 and rel_err(dbt, br.grad) <= 1e-5


# --------------------------------------------------------------------------------------------- GEGLU
@pytest.mark.parametrize("prec", PRECS)
def test_geglu_fwd_bwd(prec):
    m, hd = 300, 2048
    g = torch.Generator().manual_seed(3)
    T = _T(prec)
    h_plain = (torch.randn((m, 2 * hd), generator=g) * 1.5).to(DEV).to(T)        # [u | g] column order of net.0
    dout = torch.randn((m, hd), generator=g).to(DEV).to(T)
    h_int = weights.geglu_interleave(h_plain.t().contiguous()).t().contiguous()    # packed (interleaved) column order
    hr = h_plain.double().requires_grad_(True)
    u, gate = hr.chunk(2, dim=-1)
    ref = u * torch.nn.functional.gelu(gate)                                       # size_invariant_timesformer.py:60-63
    ref.backward(dout.double())
    out = ops.geglu(h_int, prec)
    dh_int = ops.geglu_bwd(h_int, dout, prec)
    from mintime_b200.training import _uninterleave
    dh = _uninterleave(dh_int.t().contiguous()).t()
    assert rel_err(out, ref) <= _tol(prec, 1e-6, 3e-3)
    assert rel_err(dh, hr.grad) <= _tol(prec, 1e-6, 3e-3)


# --------------------------------------------------------------------------------------------- attention core
def attn_core_ref(qkv, mask, idm, mode, f, n, heads):
    """Attention.forward :114-141 on projected q (pre-scaled), k, v: (B,N,3*inner) -> (B,N,inner)."""
    B, N, _ = qkv.shape
    inner = heads * 64
    q, k, v = [t.view(B, N, heads, 64).permute(0, 2, 1, 3) for t in qkv.split(inner, dim=-1)]      # (B,h,N,64)
    neg = -torch.finfo(torch.float32).max
    allow_cls = torch.cat([torch.ones(B, 1, dtype=torch.bool, device=qkv.device), mask.repeat_interleave(n, 1)], 1)
    s = torch.einsum("bhd,bhjd->bhj", q[:, :, 0], k).masked_fill(~allow_cls[:, None], neg)
    o_cls = torch.einsum("bhj,bhjd->bhd", s.softmax(-1), v)
    qg, kg, vg = [t[:, :, 1:].reshape(B, heads, f, n, 64) for t in (q, k, v)]
    if mode == "time":
        qg, kg, vg = [t.transpose(2, 3) for t in (qg, kg, vg)]                                    # (B,h,n,f,64)
    G = qg.shape[2]
    kc = torch.cat([k[:, :, :1, None].expand(B, heads, G, 1, 64), kg], 3)
    vc = torch.cat([v[:, :, :1, None].expand(B, heads, G, 1, 64), vg], 3)
    s = torch.einsum("bhgid,bhgjd->bhgij", qg, kc)
    if mode == "time":
        allow = mask[:, None, :] & idm                                                              # (B,q,k)
        allow = torch.cat([torch.ones(B, f, 1, dtype=torch.bool, device=qkv.device), allow], 2)
        s = s.masked_fill(~allow[:, None, None], neg)
    o = torch.einsum("bhgij,bhgjd->bhgid", s.softmax(-1), vc)
    if mode == "time":
        o = o.transpose(2, 3)
    o = torch.cat([o_cls[:, :, None], o.reshape(B, heads, f * n, 64)], 2)
    return o.permute(0, 2, 1, 3).reshape(B, N, inner)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("mode", ["time", "space"])
@pytest.mark.parametrize("f,ids", [(8, [2, 1]), (16, [4, 1, 3]), (32, [2])])
def test_divided_attention_bwd(prec, mode, f, ids):
    from mintime_b200 import synth
    B, n, heads = len(ids), 49, 8
    N = 1 + f * n
    T = _T(prec)
    meta = synth.make_batch_meta(B, f, ids, seed=f, pad_tail=True)
    g = torch.Generator().manual_seed(f)
    qkv = torch.randn((B, N, 3 * heads * 64), generator=g)
    qkv[..., :heads * 64] *= 0.35                                   # logits with a spread of a few units
    qkv = qkv.to(DEV).to(T)
    dout = torch.randn((B, N, heads * 64), generator=g).to(DEV).to(T)
    mask, idm = meta["mask"].to(DEV), meta["identities_mask"].to(DEV)
    mask_u8, idm_u8 = mask.to(torch.uint8).contiguous(), idm.to(torch.uint8).contiguous()
    qr = qkv.double().requires_grad_(True)
    ref = attn_core_ref(qr, mask, idm, mode, f, n, heads)
    ref.backward(dout.double())
    out, _ = ops.divided_attention(qkv, mask_u8, idm_u8, mode, f, n, heads, want_cls_attn=False, precision=prec)
    assert rel_err(out, ref) <= _tol(prec, 1e-5, 6e-3)              # the restatement above == the forward kernel
    dqkv = ops.divided_attention_bwd(qkv, dout, mask_u8, idm_u8, mode, f, n, heads, precision=prec)
    torch.cuda.synchronize()
    assert rel_err(dqkv, qr.grad) <= _tol(prec, 1e-5, 1e-2)
    # the CLS token's own rows (query over all keys; key of every group)
    assert rel_err(dqkv[:, 0], qr.grad[:, 0]) <= _tol(prec, 1e-5, 1e-2)


# --------------------------------------------------------------------------------------------- embeddings, head
def test_embed_bwd():
    from mintime_b200 import synth
    B, f, n, dim, rows = 3, 8, 49, 512, 8 * 1280 + 1
    N = 1 + f * n
    meta = synth.make_batch_meta(B, f, [2, 1, 1], seed=3, pad_tail=True)
    g0 = torch.randn((B, N, dim), generator=torch.Generator().manual_seed(1)).to(DEV)
    pos, se = meta["positions"].to(DEV), meta["size_embedding"].to(DEV).int().contiguous()
    dpos, dsize, dcls = ops.embed_bwd(g0, pos, se, rows, f, n)
    rp = torch.zeros((rows, dim), dtype=torch.float64, device=DEV).index_add_(0, pos.reshape(-1), g0.double().view(-1, dim))
    sidx = torch.cat([torch.zeros(B, 1, dtype=torch.long, device=DEV), se.long().repeat_interleave(n, 1)], 1)
    rs = torch.zeros((rows, dim), dtype=torch.float64, device=DEV).index_add_(0, sidx.reshape(-1), g0.double().view(-1, dim))
    assert rel_err(dpos, rp) <= 1e-6 and rel_err(dsize, rs) <= 1e-6
    assert rel_err(dcls, g0[:, 0].double().sum(0)) <= 1e-6


@pytest.mark.parametrize("classes", [1, 3])
def test_head_bwd(classes):
    B, N, dim = 5, 393, 512
    g = torch.Generator().manual_seed(classes)
    x = torch.randn((B, N, dim), generator=g).to(DEV)
    ln_g, ln_b = (torch.rand(dim, generator=g) + 0.5).to(DEV), torch.randn(dim, generator=g).to(DEV)
    w, dl = torch.randn((classes, dim), generator=g).to(DEV), torch.randn((B, classes), generator=g).to(DEV)
    xr, gr, br, wr = [t.double().requires_grad_(True) for t in (x, ln_g, ln_b, w)]
    bias = torch.zeros(classes, dtype=torch.float64, device=DEV, requires_grad=True)
    (torch.nn.functional.layer_norm(xr[:, 0], (dim,), gr, br, 1e-5) @ wr.t() + bias).backward(dl.double())
    gx = torch.zeros_like(x)
    dW, db, dgm, dbt = ops.head_bwd_(gx, x, ln_g, ln_b, w, dl)
    assert rel_err(gx, xr.grad) <= 1e-5
    assert rel_err(dW, wr.grad) <= 1e-5 and rel_err(db, bias.grad) <= 1e-6
    assert rel_err(dgm, gr.grad) <= 1e-5 and rel_err(dbt, br.grad) <= 1e-5


# --------------------------------------------------------------------------------------------- whole training step
def _step(model, cfg, meta, feats, labels, pos_weight):
    """the body of train.py:355-377 (loss on the host like the reference: y_pred.cpu())"""
    y = model(feats, mask=meta["mask"].to(DEV), size_embedding=meta["size_embedding"],
              identities_mask=meta["identities_mask"].to(DEV), positions=meta["positions"].to(DEV))
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([pos_weight]))(y.cpu(), labels)
    return y, loss


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("case", list(GRAD_CASES))
def test_backward_matches_reference_gradients(prec, case):
    cfg, tsd, meta, feats, labels, pw = grad_case_inputs(case)
    gold = load_golden("grads_" + case)
    model = SizeInvariantTimeSformer(config=cfg, precision=prec)
    model.load_state_dict(tsd)
    model = model.to(DEV).train()
    x = feats.to(DEV) if prec == "fp32" else feats.to(DEV).bfloat16()
    y, loss = _step(model, cfg, meta, x, labels, pw)
    loss.backward()
    torch.cuda.synchronize()
    assert np.abs(y.detach().cpu().numpy() - gold["logits"]).max() <= _tol(prec, 2e-4, 1e-2)
    assert abs(loss.item() - float(gold["loss"])) <= _tol(prec, 1e-4, 5e-3)
    bad, report = [], []
    f, n = cfg["model"]["num-frames"], cfg["model"]["num-patches"]
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        ref_s, ref_n = gold[f"grad.{k}.sample"], float(gold[f"grad.{k}.norm"])
        got_s, got_n = sample(p.grad, 512), float(p.grad.double().norm())
        if f"grad.{k}.head_sample" in gold:      # embedding tables: dense sample of the reachable rows
            ref_s = gold[f"grad.{k}.head_sample"]
            got_s = sample(p.grad[:(1 + f * n) if k.startswith("pos") else 21], 2048)
        e = float(np.linalg.norm(got_s - ref_s) / (np.linalg.norm(ref_s) + 1e-30))
        cos = float(np.dot(got_s, ref_s) / (np.linalg.norm(got_s) * np.linalg.norm(ref_s) + 1e-30))
        nerr = abs(got_n - ref_n) / (ref_n + 1e-30)
        report.append((k, e, cos, nerr))
        # fp32: arithmetic parity.  bf16: activations and both GEMM operands are rounded to 8 bits of mantissa, so the
        # bar is the direction and size of every parameter's gradient ("gradient cosine vs oracle")
        ok = (e <= 2e-3 and nerr <= 2e-3) if prec == "fp32" else (cos >= 0.97 and nerr <= 0.10)
        if not ok:
            bad.append((k, e, cos, nerr))
    report.sort(key=lambda r: -r[1])
    print("largest sample errors (name, rel err, cosine, norm err):", report[:5])
    assert not bad, bad


@pytest.mark.parametrize("switch", ["enable-pos-emb", "enable-size-emb"])
def test_backward_with_an_embedding_switched_off(switch):
    """enable-pos-emb: False still adds pos_emb(arange(N)) (size_invariant_timesformer.py:237-238), so the table keeps
    training; enable-size-emb: False drops the size table.  fp32 path vs the oracle differentiated by torch autograd."""
    cfg, tsd, meta, feats, labels, pw = grad_case_inputs("b3_f16_mixed_d2")
    cfg = copy.deepcopy(cfg)
    cfg["model"][switch] = False
    tsd = {k: v for k, v in tsd.items() if cfg["model"]["enable-size-emb"] or not k.startswith("size_emb")}
    model = SizeInvariantTimeSformer(config=cfg, precision="fp32")
    model.load_state_dict(tsd)
    model = model.to(DEV).train()
    y, loss = _step(model, cfg, meta, feats.to(DEV), labels, pw)
    loss.backward()
    osd = {k: v.clone().requires_grad_(True) for k, v in tsd.items()}
    ol, _ = orc.tsf_forward(osd, cfg, feats, meta["mask"], meta["identities_mask"], meta["size_embedding"], meta["positions"])
    torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([pw]))(ol, labels).backward()
    assert (y.detach().cpu() - ol.detach()).abs().max() <= 2e-4
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        assert float(osd[k].grad.norm()) > 0, k
        assert rel_err(p.grad.cpu(), osd[k].grad) <= 2e-3, k


def test_sgd_loss_curve_matches_oracle():
    """4 SGD steps (config: lr 0.01, weight decay 1e-4, train.py:266-268) on the same batch: CUDA fp32 path vs the
    oracle differentiated by torch autograd on the host."""
    case = "b3_f16_mixed_d2"
    cfg, tsd, meta, feats, labels, pw = grad_case_inputs(case)
    lr, wd = cfg["training"]["lr"], cfg["training"]["weight-decay"]
    model = SizeInvariantTimeSformer(config=cfg, precision="fp32")
    model.load_state_dict(tsd)
    model = model.to(DEV).train()
    opt = torch.optim.SGD(model.parameters(), lr=lr, weight_decay=wd)
    osd = {k: v.clone().requires_grad_(True) for k, v in tsd.items()}
    oopt = torch.optim.SGD(list(osd.values()), lr=lr, weight_decay=wd)
    lossf = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([pw]))
    got, want = [], []
    for _ in range(4):
        opt.zero_grad()
        _, loss = _step(model, cfg, meta, feats.to(DEV), labels, pw)
        loss.backward()
        opt.step()
        got.append(loss.item())
        oopt.zero_grad()
        logits, _ = orc.tsf_forward(osd, cfg, feats, meta["mask"], meta["identities_mask"], meta["size_embedding"],
                                    meta["positions"])
        ol = lossf(logits, labels)
        ol.backward()
        oopt.step()
        want.append(ol.item())
    assert np.allclose(got, want, rtol=0, atol=2e-4), (got, want)
    assert len(set(round(v, 6) for v in want)) > 1           # the steps do move the loss
    for k, p in model.named_parameters():
        assert rel_err(p.detach().cpu(), osd[k].detach()) <= 1e-4, k


def test_eval_after_train_uses_updated_weights_and_frozen_params_get_no_grad():
    case = "b3_f16_mixed_d2"
    cfg, tsd, meta, feats, labels, pw = grad_case_inputs(case)
    model = SizeInvariantTimeSformer(config=cfg, precision="bf16")
    model.load_state_dict(tsd)
    model = model.to(DEV).train()
    model.pos_emb.weight.requires_grad_(False)
    x = feats.to(DEV).bfloat16()
    y0, loss = _step(model, cfg, meta, x, labels, pw)
    loss.backward()
    assert model.pos_emb.weight.grad is None and model.cls_token.grad is not None
    with torch.no_grad():
        e0 = model(x, mask=meta["mask"].to(DEV), size_embedding=meta["size_embedding"],
                   identities_mask=meta["identities_mask"].to(DEV), positions=meta["positions"].to(DEV))
    assert (e0 - y0.detach()).abs().max() <= 2e-2          # fused-GEGLU inference path vs stored-h training path
    torch.optim.SGD(model.parameters(), lr=0.5).step()
    with torch.no_grad():
        e1 = model(x, mask=meta["mask"].to(DEV), size_embedding=meta["size_embedding"],
                   identities_mask=meta["identities_mask"].to(DEV), positions=meta["positions"].to(DEV))
    assert (e1 - e0).abs().max() > 1e-3                      # the packed copies were rebuilt from the new parameters
    with pytest.raises(NotImplementedError):
        model(x.clone().float().requires_grad_(True).bfloat16(), mask=meta["mask"].to(DEV),
              size_embedding=meta["size_embedding"], identities_mask=meta["identities_mask"].to(DEV),
              positions=meta["positions"].to(DEV))


def test_graphed_train_step_matches_eager_steps():
    """GraphedTrainStep (one CUDA-graph replay per train.py step) follows the eager loop: same losses and parameters
    after the same number of SGD steps on the same batch (fp32 atomics in the embedding tables: allclose, not equal)."""
    from mintime_b200 import EfficientNet, synth
    from mintime_b200.graphed import GraphedTrainStep
    case = "b3_f16_mixed_d2"
    cfg, tsd, meta, _, labels, pw = grad_case_inputs(case)
    B, f = 3, 16
    ext = EfficientNet.from_name("efficientnet-b0", precision="bf16")
    ext.load_state_dict(synth.make_effnet_state_dict(1234))
    ext = ext.to(DEV).eval()
    frames = synth.make_frames(B, f, seed=7, mask=meta["mask"], dtype=torch.uint8).to(DEV)
    lossf = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([pw], device=DEV))

    def make():
        m = SizeInvariantTimeSformer(config=cfg, precision="bf16")
        m.load_state_dict(tsd)
        m = m.to(DEV).train()
        return m, torch.optim.SGD(m.parameters(), lr=0.01, weight_decay=1e-4)

    # eager: 3 warm-up steps + 3 more
    model, opt = make()
    eager = []
    for _ in range(6):
        with torch.no_grad():
            feats = ext(frames.view(B * f, 224, 224, 3).permute(0, 3, 1, 2)).reshape(B, f, 1280, 7, 7)
        opt.zero_grad(set_to_none=True)
        y = model(feats, mask=meta["mask"].to(DEV), size_embedding=meta["size_embedding"],
                  identities_mask=meta["identities_mask"].to(DEV), positions=meta["positions"].to(DEV))
        loss = lossf(y, labels.to(DEV))
        loss.backward()
        opt.step()
        eager.append(loss.item())
    # graphed: 3 eager warm-up steps inside capture(), the capture itself does not execute, then 3 replays
    model2, opt2 = make()
    gs = GraphedTrainStep(ext, model2, opt2, lossf, B, f, frame_dtype=torch.uint8, device=DEV, warmup=3)
    gs.static["videos"].copy_(frames)
    for k in ("mask", "identities_mask", "size_embedding", "positions"):
        gs.static[k].copy_(meta[k])
    gs.static["labels"].copy_(labels)
    gs.capture()
    assert gs.kernels_per_replay > 150                       # extractor + 2-layer forward and backward, all in the graph
    got = [gs.replay().item() for _ in range(3)]
    assert np.allclose(got, eager[3:], rtol=0, atol=2e-3), (got, eager)
    for (k, p), q in zip(model.named_parameters(), model2.parameters()):
        assert rel_err(q, p) <= 2e-3, k


def test_training_forward_returns_attention_maps_too():
    """require_attention=True in a training step (the reference returns (logits, [space, time]) in any mode, :271-276):
    the maps equal the inference path's, carry no gradient, and loss.backward() still fills every parameter."""
    case = "b3_f16_mixed_d2"
    cfg, tsd, meta, feats, labels, pw = grad_case_inputs(case)
    model = SizeInvariantTimeSformer(config=cfg, require_attention=True, precision="bf16")
    model.load_state_dict(tsd)
    model = model.to(DEV).train()
    x = feats.to(DEV).bfloat16()
    kw = dict(mask=meta["mask"].to(DEV), size_embedding=meta["size_embedding"],
              identities_mask=meta["identities_mask"].to(DEV), positions=meta["positions"].to(DEV))
    y, (sa, ta) = model(x, **kw)
    assert y.requires_grad and not sa.requires_grad and not ta.requires_grad
    assert sa.shape == (3 * 8, 1, 785) and ta.shape == (3 * 8, 1, 785)
    torch.nn.functional.binary_cross_entropy_with_logits(y, labels.to(DEV)).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    with torch.no_grad():
        y2, (sa2, ta2) = model(x, **kw)
    assert (y2 - y.detach()).abs().max() <= 2e-2
    assert rel_err(sa, sa2) <= 1e-2 and rel_err(ta, ta2) <= 1e-2
    assert torch.allclose(sa.sum(-1), torch.ones_like(sa.sum(-1)), atol=1e-3)
