"""Training path of ``SizeInvariantTimeSformer``: what ``loss.backward()`` runs in the reference's train.py:376-378.

Scope: the transformer's forward / backward.  With a frozen extractor (``--freeze_backbone``: train.py:153-154, 344-346)
the features arrive without a graph; with a trainable one (train.py:155-170, efficientnet_train.py) the backward also
returns d loss / d features (the patch embedding's data gradient) and autograd carries it into the extractor.  The forward is the schedule of size_invariant_timesformer.py:224-276 with the activations the backward needs
kept; the backward is hand-scheduled over the kernels of csrc/train.cu and the forward GEMMs (dgrad / wgrad as GEMMs on
transposed operands).  It is exposed as ONE ``torch.autograd.Function`` so that the reference's training loop
(``optimizer.zero_grad(); loss.backward(); optimizer.step()``) works unchanged: the gradients land in ``param.grad``
(float32, the parameters stay the float32 master copy; compute is bf16 or fp32 as the model's ``precision`` says).

Data parallel training (BASELINE configs[3]): ``attach_grad_sync(model, group)`` makes the backward all-reduce each
layer's flat gradient buffer (NCCL over NVLink on the GPU box, gloo in the CPU tests) as soon as that layer is done,
so the exchange of layer l overlaps the backward of layers l-1..0; the buffers are averaged over the ranks before they
are handed to autograd.  There is no PyTorch fallback for any kernel.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib, ops, weights


def _uninterleave(t: torch.Tensor) -> torch.Tensor:
    """Inverse of weights.geglu_interleave along dim 0: blocks of 64 rows (32 value + 32 gate) -> [values | gates]."""
    h = t.shape[0] // 2
    b = t.reshape(h // 32, 2, 32, *t.shape[1:])
    return torch.cat([b[:, 0].reshape(h, *t.shape[1:]), b[:, 1].reshape(h, *t.shape[1:])], dim=0)


class TrainPack:
    """Device-side packing of the live parameters for forward + backward (no host round trip: it is redone after every
    optimizer step).  Holds the ctypes weight struct of the forward kernels plus the transposed copies the dgrad GEMMs
    read."""

    def __init__(self, model, precision: str, device):
        m = model.config["model"]
        T = _lib.torch_dtype(precision)
        f32 = torch.float32
        inner = m["heads"] * m["dim-head"]
        scale = m["dim-head"] ** -0.5
        self.keep: List[torch.Tensor] = []
        W = _lib.TsfWeights()
        sd = {k: v.detach() for k, v in model.named_parameters()}

        def dev(t, dt):
            t = t.to(device=device, dtype=dt).contiguous()
            self.keep.append(t)
            return t

        self.w_patch = dev(sd["to_patch_embedding.weight"], T)
        self.w_patch_t = dev(self.w_patch.t(), T)          # [C][dim]: data gradient w.r.t. the extractor's features
        W.w_patch = self.w_patch.data_ptr()
        W.b_patch = dev(sd["to_patch_embedding.bias"], f32).data_ptr()
        W.cls_token = dev(sd["cls_token"].reshape(-1), f32).data_ptr()
        W.pos_emb = dev(sd["pos_emb.weight"], f32).data_ptr()
        W.size_emb = dev(sd["size_emb.weight"], f32).data_ptr() if m["enable-size-emb"] else None
        self.layers: List[Dict[str, torch.Tensor]] = []
        for l in range(m["depth"]):
            L: Dict[str, torch.Tensor] = {}
            for j, name, arr in ((0, "time", W.time_attn), (1, "space", W.space_attn)):
                p = f"layers.{l}.{j}."
                wq = sd[p + "fn.to_qkv.weight"].clone()
                wq[:inner] *= scale
                L[name + ".ln_g"] = dev(sd[p + "norm.weight"], f32)
                L[name + ".ln_b"] = dev(sd[p + "norm.bias"], f32)
                L[name + ".wqkv"] = dev(wq, T)
                L[name + ".wqkv_t"] = dev(L[name + ".wqkv"].t(), T)
                L[name + ".wo"] = dev(sd[p + "fn.to_out.0.weight"], T)
                L[name + ".wo_t"] = dev(L[name + ".wo"].t(), T)
                L[name + ".bo"] = dev(sd[p + "fn.to_out.0.bias"], f32)
                a = arr[l]
                a.ln_g, a.ln_b = L[name + ".ln_g"].data_ptr(), L[name + ".ln_b"].data_ptr()
                a.w_qkv, a.w_out, a.b_out = L[name + ".wqkv"].data_ptr(), L[name + ".wo"].data_ptr(), L[name + ".bo"].data_ptr()
            p = f"layers.{l}.2."
            L["ff.ln_g"] = dev(sd[p + "norm.weight"], f32)
            L["ff.ln_b"] = dev(sd[p + "norm.bias"], f32)
            L["ff.w1"] = dev(weights.geglu_interleave(sd[p + "fn.net.0.weight"]), T)
            L["ff.w1_t"] = dev(L["ff.w1"].t(), T)
            L["ff.b1"] = dev(weights.geglu_interleave(sd[p + "fn.net.0.bias"]), f32)
            L["ff.w2"] = dev(sd[p + "fn.net.3.weight"], T)
            L["ff.w2_t"] = dev(L["ff.w2"].t(), T)
            L["ff.b2"] = dev(sd[p + "fn.net.3.bias"], f32)
            ff = W.ff[l]
            ff.ln_g, ff.ln_b = L["ff.ln_g"].data_ptr(), L["ff.ln_b"].data_ptr()
            ff.w1, ff.b1, ff.w2, ff.b2 = (L["ff.w1"].data_ptr(), L["ff.b1"].data_ptr(), L["ff.w2"].data_ptr(),
                                          L["ff.b2"].data_ptr())
            self.layers.append(L)
        self.out_ln_g = dev(sd["to_out.0.weight"], f32)
        self.out_ln_b = dev(sd["to_out.0.bias"], f32)
        self.out_w = dev(sd["to_out.1.weight"], f32)
        self.out_b = dev(sd["to_out.1.bias"], f32)
        W.out_ln_g, W.out_ln_b = self.out_ln_g.data_ptr(), self.out_ln_b.data_ptr()
        W.out_w, W.out_b = self.out_w.data_ptr(), self.out_b.data_ptr()
        self.struct = W


class GradSync:
    """Bucketed (one flat fp32 bucket per layer + one for the small tensors) gradient all-reduce issued from inside the
    backward; the buckets are averaged when the backward ends.  The collectives are plain ``dist.all_reduce`` calls
    (NCCL on the GPU box, gloo in the CPU tests): under ``GraphedTrainStep`` they are captured with the backward."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.pending = []
        # deferred = True (set by GraphedTrainStep): the backward only RECORDS its buckets -- they are static tensors of
        # the captured graph -- and exchange() all-reduces them after the replay, outside the capture
        self.deferred = False
        self.buckets = []

    def launch(self, flat: torch.Tensor):
        if not flat.is_contiguous():
            raise ValueError("GradSync.launch: gradient buckets must be contiguous (a copy would never be handed to autograd)")
        if self.deferred:
            self.buckets.append(flat)
        elif self.world > 1:
            self.pending.append((self.dist.all_reduce(flat, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True), flat))

    def finish(self):
        for work, flat in self.pending:
            work.wait()
            flat.mul_(1.0 / self.world)
        self.pending = []

    def exchange(self):
        """all-reduce + average the buckets a deferred backward recorded (call after the graph replay)"""
        if self.world > 1:
            works = [self.dist.all_reduce(b, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True) for b in self.buckets]
            for w, b in zip(works, self.buckets):
                w.wait()
                b.mul_(1.0 / self.world)


def attach_grad_sync(model, group=None, broadcast_parameters: bool = True):
    """Average the gradients over the ranks of `group` inside every backward of `model` (data parallel training): the
    ``SizeInvariantTimeSformer`` (per-layer buckets) or an ``EfficientNet`` in train mode (one bucket; BatchNorm statistics
    stay per replica, as under the reference's ``nn.DataParallel``).  Like DistributedDataParallel, rank 0's parameters and
    buffers are broadcast first so the replicas start identical whatever each rank seeded."""
    model._grad_sync = GradSync(group)
    if broadcast_parameters and model._grad_sync.world > 1:
        import torch.distributed as dist
        src = dist.get_global_rank(group, 0) if group is not None else 0
        with torch.no_grad():
            for p in model.parameters():
                dist.broadcast(p.data, src=src, group=group)
            for b in model.buffers():            # (the extractor's BatchNorm running statistics; DDP broadcasts buffers too)
                dist.broadcast(b.data, src=src, group=group)
        model._train_pack = None
        model._packed = None
    return model


def _attn_layout(dim, inner):
    return [("norm.weight", (dim,)), ("norm.bias", (dim,)), ("fn.to_qkv.weight", (3 * inner, dim)),
            ("fn.to_out.0.weight", (dim, inner)), ("fn.to_out.0.bias", (dim,))]


def _ff_layout(dim):
    return [("norm.weight", (dim,)), ("norm.bias", (dim,)), ("fn.net.0.weight", (8 * dim, dim)), ("fn.net.0.bias", (8 * dim,)),
            ("fn.net.3.weight", (dim, 4 * dim)), ("fn.net.3.bias", (dim,))]


def _carve(flat: torch.Tensor, off: int, shape):
    n = 1
    for s in shape:
        n *= s
    return flat[off:off + n].view(*shape), off + n


class TsfTrainFunction(torch.autograd.Function):
    """forward(model, tokens, mask_u8, idmask_u8, size_embedding, positions, *parameters) -> logits [, space, time maps]"""

    @staticmethod
    def forward(ctx, model, tok, mask_u8, idm_u8, se, pos, *params):
        precision = model.precision
        T = _lib.torch_dtype(precision)
        prec = _lib.prec_id(precision)
        lib = _lib.load()
        dev = tok.device
        B, fn, C = tok.shape
        f, n, dim, heads, dh = model.num_frames, model.num_patches, model.dim, model.heads, model.dim_head
        N = 1 + f * n
        M = B * N
        pk: TrainPack = model._get_train_pack(dev)
        x = torch.empty((M, dim), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.mt_patch_embed_fwd(prec, pk.struct, model._cfg_struct, tok.data_ptr(), _lib.ptr(se), _lib.ptr(pos),
                                        x.data_ptr(), B, _lib.stream_ptr())
        _lib.check(rc, "mt_patch_embed_fwd")
        saved = []
        maps = {}
        for l, L in enumerate(pk.layers):
            S = {}
            for name in ("time", "space"):
                want_maps = model.require_attention and l == len(pk.layers) - 1
                xn, x_in = ops.layernorm_copy(x, L[name + ".ln_g"], L[name + ".ln_b"], precision)   # (x is updated in place below)
                qkv = ops.pointwise(xn, L[name + ".wqkv"], precision=precision)
                ao, cls = ops.divided_attention(qkv.view(B, N, -1), mask_u8, idm_u8, name, f, n, heads, dh,
                                                want_cls_attn=want_maps, precision=precision)
                ao = ao.view(M, -1)
                ops.linear_residual_(x, ao, L[name + ".wo"], L[name + ".bo"], precision)
                S[name] = (x_in, xn, qkv, ao)
                if want_maps:
                    maps[name] = cls.view(B * heads, 1, N)
            xn, x_in = ops.layernorm_copy(x, L["ff.ln_g"], L["ff.ln_b"], precision)
            h = ops.pointwise(xn, L["ff.w1"], shift=L["ff.b1"], precision=precision)
            go = ops.geglu(h, precision)
            ops.linear_residual_(x, go, L["ff.w2"], L["ff.b2"], precision)
            S["ff"] = (x_in, xn, h, go)
            saved.append(S)
        x3 = x.view(B, N, dim)
        logits = ops.head(x3, pk.out_ln_g, pk.out_ln_b, pk.out_w, pk.out_b)
        ctx.model, ctx.pk, ctx.saved, ctx.x_final = model, pk, saved, x3
        ctx.inputs = (tok, mask_u8, idm_u8, se, pos)
        ctx.param_names = [k for k, _ in model.named_parameters()]
        ctx.n_extra = 6
        if model.require_attention:
            ctx.mark_non_differentiable(maps["space"], maps["time"])
            return logits, maps["space"], maps["time"]
        return logits

    @staticmethod
    def backward(ctx, dlogits, *_unused):
        if ctx.saved is None:
            raise RuntimeError("mintime_b200: the activations of this forward were released by its first backward "
                               "(retain_graph=True is not supported: call forward again)")
        model, pk = ctx.model, ctx.pk
        precision = model.precision
        tok, mask_u8, idm_u8, se, pos = ctx.inputs
        dev = tok.device
        B = tok.shape[0]
        f, n, dim, heads, dh = model.num_frames, model.num_patches, model.dim, model.heads, model.dim_head
        inner = heads * dh
        N = 1 + f * n
        M = B * N
        scale = dh ** -0.5
        sync: Optional[GradSync] = getattr(model, "_grad_sync", None)
        grads: Dict[str, torch.Tensor] = {}
        f32 = torch.float32

        g = torch.zeros((B, N, dim), dtype=f32, device=dev)       # gradient of the fp32 residual stream
        dW, db, dgam, dbet = ops.head_bwd_(g, ctx.x_final, pk.out_ln_g, pk.out_ln_b, pk.out_w,
                                           dlogits.to(device=dev, dtype=f32).contiguous())
        grads["to_out.1.weight"], grads["to_out.1.bias"] = dW, db
        grads["to_out.0.weight"], grads["to_out.0.bias"] = dgam, dbet
        g = g.view(M, dim)

        # The weight-gradient GEMMs have few output tiles (512 x 512 .. 4096 x 512 over a 25 k-long contraction) and
        # nothing on the critical path reads them: they run on side streams, next to the main stream's kernels, and
        # their post-processing (q-row scale, GEGLU row order) follows them on the same side stream.
        main = torch.cuda.current_stream(dev)
        # (model._serial_wgrad = True keeps them on the main stream: clean per-kernel timings for bench.py's table)
        side = [main] if getattr(model, "_serial_wgrad", False) else _side_streams(model, dev)
        turn = [0]

        nt_ok = precision == "bf16" and not getattr(model, "_wgrad_transposed", False)   # (the flag keeps the old route: A/B runs)

        def wgrad(dst, dy, x, after=None, **dy_kw):
            """dst += dy^T x on a side stream: the split-K GEMM (fp32 path / CLS-skipping operand: after both
            transposes by mt_grad_prep) and `after`."""
            s_ = side[turn[0] % len(side)]
            turn[0] += 1
            if s_ is not main:
                s_.wait_stream(main)
            with torch.cuda.stream(s_):
                if nt_ok and not dy_kw and dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16:
                    # bf16 path: the operands enter the tensor cores as they lie in memory (MN-major descriptors)
                    ops.linear_wgrad_nt_(dst, dy, x)
                else:
                    _, dyT, _ = ops.grad_prep(dy, want_t=True, precision=precision, **dy_kw)
                    _, xT, _ = ops.grad_prep(x, want_t=True, precision=precision)
                    ops.linear_wgrad_(dst, dyT, xT, precision)
                if after is not None:
                    after()
            if s_ is not main:
                for t in (dst, dy, x):
                    t.record_stream(s_)

        def side_done():
            evs = []
            for s_ in side:
                e = torch.cuda.Event()
                e.record(s_)
                evs.append(e)
            return evs

        prev = None          # (flat buffer, side-stream events) of the layer processed before the current one

        attn_lay, ff_lay = _attn_layout(dim, inner), _ff_layout(dim)
        layer_numel = 2 * sum(_numel(s) for _, s in attn_lay) + sum(_numel(s) for _, s in ff_lay)
        for l in range(len(pk.layers) - 1, -1, -1):
            L, S = pk.layers[l], ctx.saved[l]
            flat = torch.zeros((layer_numel,), dtype=f32, device=dev)
            off = 0
            G: Dict[str, torch.Tensor] = {}
            for j, lay in ((0, attn_lay), (1, attn_lay), (2, ff_lay)):
                for k, shape in lay:
                    G[f"{j}.{k}"], off = _carve(flat, off, shape)
            # ---- feed-forward sub-block  x3 = x2 + W2 geglu(W1 LN(x2) + b1) + b2   (:65-76, :268)
            x_in, xn, h, go = S["ff"]
            # (gb: the T copy of g the GEMMs read; g itself keeps changing, so the side stream transposes gb)
            gb, _, cs = ops.grad_prep(g, want_rm=True, want_colsum=True, precision=precision)
            G["2.fn.net.3.bias"].copy_(cs)
            dgo = ops.pointwise(gb, L["ff.w2_t"], precision=precision)
            wgrad(G["2.fn.net.3.weight"], gb, go)
            dh_, cs = ops.geglu_bwd_colsum(h, dgo, precision)      # (+ the bias gradient of net.0 in the same pass)
            del dgo, gb
            dw1 = torch.zeros((8 * dim, dim), dtype=f32, device=dev)      # interleaved row order of the packed weight
            wgrad(dw1, dh_, xn, after=lambda dw1=dw1, dst=G["2.fn.net.0.weight"]: dst.copy_(_uninterleave(dw1)))
            G["2.fn.net.0.bias"].copy_(_uninterleave(cs))
            dxn = ops.pointwise(dh_, L["ff.w1_t"], precision=precision)
            del dh_, dw1
            dgm, dbt = ops.layernorm_bwd_(g, x_in, L["ff.ln_g"], dxn, precision)
            G["2.norm.weight"].copy_(dgm)
            G["2.norm.bias"].copy_(dbt)
            # ---- attention sub-blocks, space then time   x' = x + Wo attn(Wqkv LN(x)) + bo   (:109-144, :265-267)
            for j, name in ((1, "space"), (0, "time")):
                x_in, xn, qkv, ao = S[name]
                gb, _, cs = ops.grad_prep(g, want_rm=True, want_colsum=True, precision=precision)
                G[f"{j}.fn.to_out.0.bias"].copy_(cs)
                dao = ops.pointwise(gb, L[name + ".wo_t"], precision=precision)
                wgrad(G[f"{j}.fn.to_out.0.weight"], gb, ao)
                dqkv = ops.divided_attention_bwd(qkv.view(B, N, -1), dao.view(B, N, -1), mask_u8, idm_u8, name, f, n,
                                                 heads, dh, precision).view(M, -1)
                del dao, gb
                wq = G[f"{j}.fn.to_qkv.weight"]
                # (the packed q rows carry dim_head^-0.5, :114)
                wgrad(wq, dqkv, xn, after=lambda wq=wq: wq[:inner].mul_(scale))
                dxn = ops.pointwise(dqkv, L[name + ".wqkv_t"], precision=precision)
                del dqkv
                dgm, dbt = ops.layernorm_bwd_(g, x_in, L[name + ".ln_g"], dxn, precision)
                G[f"{j}.norm.weight"].copy_(dgm)
                G[f"{j}.norm.bias"].copy_(dbt)
            ctx.saved[l] = None                      # release this layer's activations
            if sync is not None:
                # the exchange of the PREVIOUS layer's bucket is issued now: its side-stream GEMMs have long finished
                if prev is not None:
                    for e in prev[1]:
                        main.wait_event(e)
                    sync.launch(prev[0])
                prev = (flat, side_done())
            for k, v in G.items():
                grads[f"layers.{l}.{k}"] = v
        # ---- token build (:225-248)
        g3 = g.view(B, N, dim)
        rows = model.pos_emb.weight.shape[0]
        # pos_emb trains in BOTH modes: with enable-pos-emb off the reference still adds pos_emb(arange(N))
        # (size_invariant_timesformer.py:237-238); `pos` is None then and the kernel scatters to rows 0..N-1
        dpos, dsize, dcls = ops.embed_bwd(g3, pos, se, rows, f, n, want_pos=True, want_size=bool(model.enable_size_emb))
        Mt = B * f * n
        want_dtok = ctx.needs_input_grad[1]            # an unfrozen extractor (train.py:155-170) wants d loss / d features
        gtok, _, cs = ops.grad_prep(g, want_rm=want_dtok, want_colsum=True, rows_per_batch=f * n, m=Mt, precision=precision)
        dtok = ops.pointwise(gtok, pk.w_patch_t, precision=precision).view(B, f * n, -1) if want_dtok else None
        dwp = torch.zeros((dim, model.channels), dtype=f32, device=dev)
        wgrad(dwp, g, tok.view(Mt, -1), rows_per_batch=f * n, m=Mt)     # (g is final here: nothing writes it any more)
        for s_ in side:
            if s_ is not main:
                main.wait_stream(s_)
        if sync is not None and prev is not None:
            sync.launch(prev[0])
        grads["to_patch_embedding.weight"], grads["to_patch_embedding.bias"] = dwp, cs
        grads["cls_token"] = dcls.view(1, dim)
        grads["pos_emb.weight"] = dpos
        if model.enable_size_emb:
            grads["size_emb.weight"] = dsize
        if sync is not None:
            # the tensors outside the layers travel as ONE flat bucket; the views handed to autograd alias it
            names = ("to_out.1.weight", "to_out.1.bias", "to_out.0.weight", "to_out.0.bias", "to_patch_embedding.weight",
                     "to_patch_embedding.bias", "cls_token", "pos_emb.weight") + (("size_emb.weight",) if model.enable_size_emb else ())
            flat = torch.empty((sum(grads[k].numel() for k in names),), dtype=f32, device=dev)
            off = 0
            for k in names:
                view, off = _carve(flat, off, tuple(grads[k].shape))
                view.copy_(grads[k])
                grads[k] = view
            sync.launch(flat)
            sync.finish()
        ctx.saved = None
        out = [None] * ctx.n_extra
        out[1] = dtok
        for name, p in zip(ctx.param_names, model.parameters()):
            gr = grads.get(name)
            out.append(gr.view_as(p) if (gr is not None and p.requires_grad) else None)
        return tuple(out)


def _side_streams(model, dev, count: int = 4):
    st = getattr(model, "_side", None)
    if st is None or st[0] != str(dev):
        st = (str(dev), [torch.cuda.Stream(device=dev) for _ in range(count)])
        model._side = st
    return st[1]


def _numel(shape) -> int:
    n = 1
    for s in shape:
        n *= s
    return n
