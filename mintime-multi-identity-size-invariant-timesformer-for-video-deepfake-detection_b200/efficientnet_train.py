"""Train-mode forward + backward of the EfficientNet-B0 extractor (reference train.py:153-170: the extractor is in
``.train()`` and receives gradients unless ``--freeze_backbone``; ``--extractor_unfreeze_blocks k`` leaves only the last
k MBConv blocks trainable, :157-167).

What ``.train()`` changes on the path: BatchNorm normalises with BATCH statistics and updates its running
statistics (utils.py:520-521: momentum 0.01, eps 1e-3), and every skip block applies drop-connect with rate
``0.2 * idx / 16`` (model.py:125-127, 279-282; utils.py:129-154).  So nothing can be folded into the convolution weights:
the forward keeps every raw convolution output, its batch statistics and the activated tensors, and the backward walks the
layers in reverse through the kernels of csrc/effnet_train.cu (BatchNorm / swish / depthwise / squeeze-excite backward)
and the library's GEMMs (1x1 convolutions: data gradient by mt_pointwise_fwd on W^T, weight gradient by mt_conv1x1_wgrad;
with ``precision="bf16"`` these GEMMs run on the tensor cores with bf16 operands and fp32 accumulation -- see ``pw``).

Exposed as ONE ``torch.autograd.Function`` so the reference loop (``loss.backward(); optimizer.step()``) works unchanged.
Arithmetic is fp32 (the exact path); there is no PyTorch fallback.  Gradients stop at the first layer that still has a
trainable parameter, so a partially unfrozen extractor only pays for the blocks it trains.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from . import _lib, ops
from .spec import B0_BLOCKS, BN_EPS, BN_MOMENTUM

DROP_CONNECT_RATE = 0.2          # utils.py:523 (GlobalParams.drop_connect_rate of efficientnet-b0)
f32 = torch.float32


class _Ws:
    """scratch for the reductions (BatchNorm partial sums, depthwise / stem weight gradients, SE backward)"""

    def __init__(self):
        self.buf = None

    def get(self, rows: int, c: int, dev):
        need = int(_lib.load().mt_extractor_train_workspace_bytes(int(rows), int(c)))
        if self.buf is None or self.buf.numel() < need or self.buf.device != dev:
            self.buf = torch.empty((need,), dtype=torch.uint8, device=dev)
        return self.buf


def _call(name, *args):
    rc = getattr(_lib.load(), name)(*args)
    _lib.check(rc, name)


def _st():
    return _lib.stream_ptr()


def bn_train(x2d: torch.Tensor, bn, ws: _Ws, update_running: bool = True):
    """-> (mean, var) of the rows of x2d [rows, C]; moves bn.running_* like nn.BatchNorm2d in train mode"""
    rows, c = x2d.shape
    mean = torch.empty((c,), dtype=f32, device=x2d.device)
    var = torch.empty((c,), dtype=f32, device=x2d.device)
    w = ws.get(rows, c, x2d.device)
    rm = bn.running_mean if update_running else None
    rv = bn.running_var if update_running else None
    _call("mt_bn_stats", x2d.data_ptr(), mean.data_ptr(), var.data_ptr(), _lib.ptr(rm), _lib.ptr(rv), float(bn.momentum), rows, c,
          w.data_ptr(), w.numel(), _st())
    if update_running:
        bn.num_batches_tracked += 1
    return mean, var


def bn_act(x2d, mean, var, gamma, beta, act: int):
    out = torch.empty_like(x2d)
    _call("mt_bn_act_fwd", x2d.data_ptr(), mean.data_ptr(), var.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act, float(BN_EPS),
          out.data_ptr(), x2d.shape[0], x2d.shape[1], _st())
    return out


def bn_act_bwd(dy2d, x2d, mean, var, gamma, beta, act: int, ws: _Ws):
    rows, c = x2d.shape
    dx = torch.empty_like(x2d)
    dg = torch.empty((c,), dtype=f32, device=x2d.device)
    db = torch.empty((c,), dtype=f32, device=x2d.device)
    w = ws.get(rows, c, x2d.device)
    _call("mt_bn_act_bwd", dy2d.data_ptr(), x2d.data_ptr(), mean.data_ptr(), var.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act,
          float(BN_EPS), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), rows, c, w.data_ptr(), w.numel(), _st())
    return dx, dg, db


def pw(precision, x2d, w2d, **kw):
    """1x1 convolution / its data gradient on fp32 [rows, cin] activations.  precision "fp32": the exact FFMA GEMM (what the
    reference-fixture tests run); "bf16": operands cast to bf16, tcgen05 GEMM with fp32 accumulation, fp32 result -- the
    mixed-precision mode of the bf16 extractor (the batch statistics, the normalisation and every other step stay fp32)."""
    if precision == "bf16" and x2d.shape[1] % 8 == 0 and w2d.shape[0] % 8 == 0:
        return ops.pointwise(x2d.bfloat16(), w2d.bfloat16(), precision="bf16", **kw).float()
    return ops.pointwise(x2d, w2d, precision="fp32", **kw)


def conv1x1_wgrad(dy2d, a2d, precision="fp32"):
    """dW [cout, cin] = dy^T a  (both [rows, *] fp32); bf16 mode: the MN-major tensor-core weight gradient on bf16 copies"""
    rows, co = dy2d.shape
    ci = a2d.shape[1]
    if precision == "bf16" and co % 8 == 0 and ci % 8 == 0:
        dw = torch.zeros((co, ci), dtype=f32, device=dy2d.device)
        return ops.linear_wgrad_nt_(dw, dy2d.bfloat16(), a2d.bfloat16())
    lib = _lib.load()
    ws = torch.empty((int(lib.mt_conv1x1_wgrad_workspace_bytes(rows, co, ci)),), dtype=torch.uint8, device=dy2d.device)
    dw = torch.empty((co, ci), dtype=f32, device=dy2d.device)
    _call("mt_conv1x1_wgrad", dy2d.data_ptr(), a2d.data_ptr(), dw.data_ptr(), rows, co, ci, ws.data_ptr(), ws.numel(), _st())
    return dw


def _taps(w):       # (c,1,k,k) -> [k*k][c]
    k = w.shape[-1]
    return w[:, 0].permute(1, 2, 0).reshape(k * k, w.shape[0]).contiguous()


def _untaps(dw, k):  # [k*k][c] -> (c,1,k,k)
    return dw.view(k, k, -1).permute(2, 0, 1).unsqueeze(1).contiguous()


class EffnetTrainFunction(torch.autograd.Function):
    """forward(ext, x_nhwc float32 (n,224,224,3), rand_fn, *parameters) -> features (n,7,7,1280) float32"""

    @staticmethod
    def forward(ctx, ext, x, rand_fn, *params):
        dev = x.device
        n = x.shape[0]
        ws = _Ws()
        P = dict(ext.named_parameters())
        saved: List[dict] = []
        # ---- stem (model.py:276): conv 3x3 s2 -> BN -> swish
        w_stem = ext._conv_stem.weight.detach().permute(2, 3, 1, 0).reshape(27, 32).contiguous()
        r = torch.empty((n, 112, 112, 32), dtype=f32, device=dev)
        _call("mt_stem_raw_fwd", x.data_ptr(), w_stem.data_ptr(), r.data_ptr(), n, 224, 224, _st())
        r2 = r.view(-1, 32)
        m, v = bn_train(r2, ext._bn0, ws)
        a = bn_act(r2, m, v, ext._bn0.weight.detach(), ext._bn0.bias.detach(), 1)
        stem = dict(x=x, r=r2, m=m, v=v)
        cur = a.view(n, 112, 112, 32)
        # ---- 16 MBConv blocks (model.py:89-128)
        for b, blk in zip(B0_BLOCKS, ext._blocks):
            S: Dict[str, object] = {"inp": cur}
            h, cin, cexp, ho = b.hw_in, b.cin, b.cexp, (b.hw_in + b.stride - 1) // b.stride
            x2 = cur.reshape(-1, cin)
            if b.expand != 1:
                r0 = pw(ext.precision, x2, blk._expand_conv.weight.detach().flatten(1).contiguous())
                m0, v0 = bn_train(r0, blk._bn0, ws)
                a0 = bn_act(r0, m0, v0, blk._bn0.weight.detach(), blk._bn0.bias.detach(), 1)
                S.update(r0=r0, m0=m0, v0=v0)
            else:
                a0 = x2
            S["a0"] = a0
            taps = _taps(blk._depthwise_conv.weight.detach())
            r1 = torch.empty((n * ho * ho, cexp), dtype=f32, device=dev)
            _call("mt_dwconv_raw_fwd", a0.data_ptr(), taps.data_ptr(), r1.data_ptr(), n, h, cexp, b.kernel, b.stride, _st())
            m1, v1 = bn_train(r1, blk._bn1, ws)
            a1 = bn_act(r1, m1, v1, blk._bn1.weight.detach(), blk._bn1.bias.detach(), 1)
            sq = b.se_squeeze
            pm = torch.empty((n, cexp), dtype=f32, device=dev)
            _call("mt_group_mean", a1.data_ptr(), pm.data_ptr(), n, ho * ho, cexp, _st())
            gate = torch.empty((n, cexp), dtype=f32, device=dev)
            s_pre = torch.empty((n, sq), dtype=f32, device=dev)
            wr = blk._se_reduce.weight.detach().flatten(1).contiguous()
            we = blk._se_expand.weight.detach().flatten(1).contiguous()
            _call("mt_se_fc_fwd", pm.data_ptr(), wr.data_ptr(), blk._se_reduce.bias.detach().data_ptr(), we.data_ptr(),
                  blk._se_expand.bias.detach().data_ptr(), gate.data_ptr(), s_pre.data_ptr(), n, cexp, sq, _st())
            r2_ = pw(ext.precision, a1, blk._project_conv.weight.detach().flatten(1).contiguous(), gate=gate, rows_per_gate=ho * ho)
            m2, v2 = bn_train(r2_, blk._bn2, ws)
            y = bn_act(r2_, m2, v2, blk._bn2.weight.detach(), blk._bn2.bias.detach(), 0)
            scale = None
            if b.has_skip:
                rate = ext.drop_connect_rate * float(b.index) / len(B0_BLOCKS) if ext.drop_connect_rate else 0.0   # model.py:279-282
                if rate:
                    keep = 1.0 - rate
                    u = rand_fn(n).to(device=dev, dtype=f32).reshape(n)                  # utils.py:146-150
                    scale = (torch.floor(keep + u) / keep).contiguous()
                out = torch.empty_like(y)
                _call("mt_scale_add", y.data_ptr(), _lib.ptr(scale), cur.reshape(-1, b.cout).data_ptr(), out.data_ptr(), n,
                      ho * ho * b.cout, _st())
                y = out
            S.update(r1=r1, m1=m1, v1=v1, a1=a1, pm=pm, gate=gate, s_pre=s_pre, r2=r2_, m2=m2, v2=v2, scale=scale)
            saved.append(S)
            cur = y.view(n, ho, ho, b.cout)
        # ---- head (model.py:286): conv 1x1 320 -> 1280 -> BN -> swish
        x2 = cur.reshape(-1, 320)
        rh = pw(ext.precision, x2, ext._conv_head.weight.detach().flatten(1).contiguous())
        mh, vh = bn_train(rh, ext._bn1, ws)
        feats = bn_act(rh, mh, vh, ext._bn1.weight.detach(), ext._bn1.bias.detach(), 1)
        ctx.ext, ctx.stem, ctx.saved, ctx.head = ext, stem, saved, dict(x=x2, r=rh, m=mh, v=vh)
        ctx.names = [k for k, _ in ext.named_parameters()]
        ctx.n = n
        return feats.view(n, 7, 7, 1280)

    @staticmethod
    def backward(ctx, dfeats):
        ext, n = ctx.ext, ctx.n
        ws = _Ws()
        G: Dict[str, torch.Tensor] = {}
        req = {k: p.requires_grad for k, p in ext.named_parameters()}

        def needs(prefix: str) -> bool:
            return any(v for k, v in req.items() if k.startswith(prefix))

        # The earliest layer with a trainable parameter: nothing before it needs a gradient (train.py:157-167).
        # order: 0 stem conv, 1 stem BN, 2..17 blocks 0..15, 18 head conv, 19 head BN
        order = ["_conv_stem", "_bn0."] + [f"_blocks.{i}." for i in range(16)] + ["_conv_head", "_bn1."]
        first = next((i for i, pfx in enumerate(order) if needs(pfx)), len(order))
        need_stem = first < 2
        fb = 0 if need_stem else first - 2                              # first block that still needs gradients

        def wT(w):
            return w.detach().flatten(1).t().contiguous()

        # ---- head
        hd = ctx.head
        dy = dfeats.contiguous().view(-1, 1280).to(f32)
        d_r, dg, db = bn_act_bwd(dy, hd["r"], hd["m"], hd["v"], ext._bn1.weight.detach(), ext._bn1.bias.detach(), 1, ws)
        G["_bn1.weight"], G["_bn1.bias"] = dg, db
        if req["_conv_head.weight"]:
            G["_conv_head.weight"] = conv1x1_wgrad(d_r, hd["x"], ext.precision).view(1280, 320, 1, 1)
        if first >= 18:
            return EffnetTrainFunction._pack(ctx, G)
        dcur = pw(ext.precision, d_r, wT(ext._conv_head.weight))          # [n*49, 320]
        del d_r
        # ---- blocks, last to first
        for b, blk, S in zip(reversed(B0_BLOCKS), reversed(list(ext._blocks)), reversed(ctx.saved)):
            p = f"_blocks.{b.index}."
            h, cin, cexp, ho, k = b.hw_in, b.cin, b.cexp, (b.hw_in + b.stride - 1) // b.stride, b.kernel
            rows_o = ho * ho
            cont = b.index > fb or need_stem                            # the gradient flows on to the previous layer
            dskip = None
            if b.has_skip:
                dskip = dcur
                if S["scale"] is not None:
                    t = torch.empty_like(dcur)
                    _call("mt_scale_add", dcur.data_ptr(), S["scale"].data_ptr(), None, t.data_ptr(), n, rows_o * b.cout, _st())
                    dcur = t
            d_r2, dg, db = bn_act_bwd(dcur, S["r2"], S["m2"], S["v2"], blk._bn2.weight.detach(), blk._bn2.bias.detach(), 0, ws)
            G[p + "_bn2.weight"], G[p + "_bn2.bias"] = dg, db
            if req[p + "_project_conv.weight"]:
                xg = torch.empty_like(S["a1"])
                _call("mt_gate_mul", S["a1"].data_ptr(), S["gate"].data_ptr(), xg.data_ptr(), n, rows_o, cexp, _st())
                G[p + "_project_conv.weight"] = conv1x1_wgrad(d_r2, xg, ext.precision).view(b.cout, cexp, 1, 1)
                del xg
            dxg = pw(ext.precision, d_r2, wT(blk._project_conv.weight))   # [n*ho*ho, cexp]
            del d_r2
            dev = dxg.device
            dgate = torch.empty((n, cexp), dtype=f32, device=dev)
            _call("mt_gate_bwd", dxg.data_ptr(), S["a1"].data_ptr(), None, None, dgate.data_ptr(), None, n, rows_o, cexp, 0, _st())
            sq = b.se_squeeze
            wr = blk._se_reduce.weight.detach().flatten(1).contiguous()
            we = blk._se_expand.weight.detach().flatten(1).contiguous()
            dpm = torch.empty((n, cexp), dtype=f32, device=dev)
            dwr = torch.empty((sq, cexp), dtype=f32, device=dev)
            dbr = torch.empty((sq,), dtype=f32, device=dev)
            dwe = torch.empty((cexp, sq), dtype=f32, device=dev)
            dbe = torch.empty((cexp,), dtype=f32, device=dev)
            w = ws.get(n * rows_o, cexp, dev)
            _call("mt_se_fc_bwd", dgate.data_ptr(), S["gate"].data_ptr(), S["s_pre"].data_ptr(), S["pm"].data_ptr(), wr.data_ptr(),
                  we.data_ptr(), dpm.data_ptr(), dwr.data_ptr(), dbr.data_ptr(), dwe.data_ptr(), dbe.data_ptr(), n, cexp, sq,
                  w.data_ptr(), w.numel(), _st())
            G[p + "_se_reduce.weight"], G[p + "_se_reduce.bias"] = dwr.view(sq, cexp, 1, 1), dbr
            G[p + "_se_expand.weight"], G[p + "_se_expand.bias"] = dwe.view(cexp, sq, 1, 1), dbe
            da1 = torch.empty_like(dxg)
            _call("mt_gate_bwd", dxg.data_ptr(), None, S["gate"].data_ptr(), dpm.data_ptr(), None, da1.data_ptr(), n, rows_o, cexp,
                  1, _st())
            del dxg
            d_r1, dg, db = bn_act_bwd(da1, S["r1"], S["m1"], S["v1"], blk._bn1.weight.detach(), blk._bn1.bias.detach(), 1, ws)
            G[p + "_bn1.weight"], G[p + "_bn1.bias"] = dg, db
            del da1
            if req[p + "_depthwise_conv.weight"]:
                dwd = torch.empty((k * k, cexp), dtype=f32, device=dev)
                w = ws.get(n * rows_o, cexp, dev)
                _call("mt_dwconv_wgrad", S["a0"].data_ptr(), d_r1.data_ptr(), dwd.data_ptr(), n, h, cexp, k, b.stride, w.data_ptr(),
                      w.numel(), _st())
                G[p + "_depthwise_conv.weight"] = _untaps(dwd, k)
            if b.expand == 1 and not cont:
                break
            da0 = torch.empty((n * h * h, cexp), dtype=f32, device=dev)
            taps = _taps(blk._depthwise_conv.weight.detach())
            _call("mt_dwconv_dgrad", d_r1.data_ptr(), taps.data_ptr(), da0.data_ptr(), n, h, cexp, k, b.stride, _st())
            del d_r1
            if b.expand != 1:
                d_r0, dg, db = bn_act_bwd(da0, S["r0"], S["m0"], S["v0"], blk._bn0.weight.detach(), blk._bn0.bias.detach(), 1, ws)
                G[p + "_bn0.weight"], G[p + "_bn0.bias"] = dg, db
                del da0
                if req[p + "_expand_conv.weight"]:
                    G[p + "_expand_conv.weight"] = conv1x1_wgrad(d_r0, S["inp"].reshape(-1, cin), ext.precision).view(cexp, cin, 1, 1)
                if not cont:
                    break
                dcur = pw(ext.precision, d_r0, wT(blk._expand_conv.weight))
                del d_r0
            else:
                dcur = da0
            if dskip is not None:
                t = torch.empty_like(dcur)
                _call("mt_scale_add", dcur.data_ptr(), None, dskip.data_ptr(), t.data_ptr(), n, h * h * cin, _st())
                dcur = t
        if need_stem:
            st = ctx.stem
            d_r, dg, db = bn_act_bwd(dcur, st["r"], st["m"], st["v"], ext._bn0.weight.detach(), ext._bn0.bias.detach(), 1, ws)
            G["_bn0.weight"], G["_bn0.bias"] = dg, db
            if req["_conv_stem.weight"]:
                dw = torch.empty((27, 32), dtype=f32, device=d_r.device)
                w = ws.get(n * 112 * 112, 32, d_r.device)
                _call("mt_stem_wgrad", st["x"].data_ptr(), d_r.data_ptr(), dw.data_ptr(), n, 224, 224, w.data_ptr(), w.numel(), _st())
                G["_conv_stem.weight"] = dw.view(3, 3, 3, 32).permute(3, 2, 0, 1).contiguous()
        return EffnetTrainFunction._pack(ctx, G)

    @staticmethod
    def _pack(ctx, G):
        out = [None, None, None]
        live = [(name, p) for name, p in zip(ctx.names, ctx.ext.parameters()) if p.requires_grad and G.get(name) is not None]
        sync = getattr(ctx.ext, "_grad_sync", None)
        if sync is not None and live:
            # data parallel (training.attach_grad_sync(extractor)): the extractor's gradients travel as ONE flat fp32 bucket,
            # averaged over the ranks before autograd sees them; the views handed back alias the bucket
            flat = torch.empty((sum(G[name].numel() for name, _ in live),), dtype=f32, device=G[live[0][0]].device)
            off = 0
            for name, _ in live:
                k = G[name].numel()
                flat[off:off + k].copy_(G[name].reshape(-1))
                G[name] = flat[off:off + k]
                off += k
            sync.launch(flat)
            sync.finish()
        for name, p in zip(ctx.names, ctx.ext.parameters()):
            g = G.get(name)
            out.append(g.view_as(p) if (g is not None and p.requires_grad) else None)
        ctx.saved = ctx.stem = ctx.head = None
        return tuple(out)


def forward_train(ext, x_nhwc: torch.Tensor, rand_fn: Optional[Callable[[int], torch.Tensor]] = None) -> torch.Tensor:
    """rand_fn(n) -> n uniform draws for one skip block's drop-connect; default = the reference's own call (utils.py:146):
    torch.rand([batch, 1, 1, 1]) from torch's global generator of the input's device, one draw per skip block in order."""
    if rand_fn is None:
        dev = x_nhwc.device
        rand_fn = lambda n: torch.rand([n, 1, 1, 1], dtype=f32, device=dev)       # noqa: E731
    return EffnetTrainFunction.apply(ext, x_nhwc, rand_fn, *ext.parameters())
