"""Tensor-level wrappers over the building-block entry points of include/mintime_b200.h.

Each function takes CUDA torch tensors (torch only provides memory + the current stream), calls the
C ABI and returns the output tensor.  Used by the per-kernel parity tests and by callers that want
to schedule the kernels themselves.  No fallbacks: everything raises without the library / a B200.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def _prep(t: torch.Tensor, dtype=None) -> torch.Tensor:
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t


def _T(precision: str):
    return _lib.torch_dtype(precision)


def pointwise(a, w, shift=None, gate=None, rows_per_gate=0, residual=None, act=0, precision="bf16"):
    """out[m,n] = act(sum_k a[m,k]*gate[m//rows_per_gate,k]*w[n,k] + shift[n]) + residual[m,n]"""
    T = _T(precision)
    _prep(a, T); _prep(w, T)
    m, k = a.shape
    n = w.shape[0]
    _lib.require_device(a.device)
    out = torch.empty((m, n), dtype=T, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.load().mt_pointwise_fwd(_lib.prec_id(precision), a.data_ptr(), w.data_ptr(), _lib.ptr(shift),
                                          _lib.ptr(gate), rows_per_gate, _lib.ptr(residual), act, out.data_ptr(),
                                          m, n, k, _lib.stream_ptr())
    _lib.check(rc, "mt_pointwise_fwd")
    return out


def linear_residual_(x, a, w, bias=None, precision="bf16"):
    """x[m,n] += a @ w.T + bias   (x float32, in place)"""
    T = _T(precision)
    _prep(a, T); _prep(w, T); _prep(x, torch.float32)
    m, k = a.shape
    n = w.shape[0]
    _lib.require_device(a.device)
    with torch.cuda.device(a.device):
        rc = _lib.load().mt_linear_residual_fwd(_lib.prec_id(precision), a.data_ptr(), w.data_ptr(), _lib.ptr(bias),
                                                x.data_ptr(), m, n, k, _lib.stream_ptr())
    _lib.check(rc, "mt_linear_residual_fwd")
    return x


def linear_geglu(a, w_interleaved, bias_interleaved=None, precision="bf16"):
    T = _T(precision)
    _prep(a, T); _prep(w_interleaved, T)
    m, k = a.shape
    n = w_interleaved.shape[0]
    _lib.require_device(a.device)
    out = torch.empty((m, n // 2), dtype=T, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.load().mt_linear_geglu_fwd(_lib.prec_id(precision), a.data_ptr(), w_interleaved.data_ptr(),
                                             _lib.ptr(bias_interleaved), out.data_ptr(), m, n, k, _lib.stream_ptr())
    _lib.check(rc, "mt_linear_geglu_fwd")
    return out


def layernorm(x, gamma, beta, precision="bf16"):
    _prep(x, torch.float32)
    rows, dim = x.shape
    _lib.require_device(x.device)
    out = torch.empty((rows, dim), dtype=_T(precision), device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().mt_layernorm_fwd(_lib.prec_id(precision), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                          out.data_ptr(), rows, dim, _lib.stream_ptr())
    _lib.check(rc, "mt_layernorm_fwd")
    return out


def layernorm_copy(x, gamma, beta, precision="bf16"):
    """(LayerNorm(x), a float32 copy of x) in one pass: the training forward keeps the sub-block input this way"""
    _prep(x, torch.float32)
    rows, dim = x.shape
    _lib.require_device(x.device)
    out = torch.empty((rows, dim), dtype=_T(precision), device=x.device)
    x_copy = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.load().mt_layernorm_copy_fwd(_lib.prec_id(precision), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                               out.data_ptr(), x_copy.data_ptr(), rows, dim, _lib.stream_ptr())
    _lib.check(rc, "mt_layernorm_copy_fwd")
    return out, x_copy


def divided_attention(qkv, mask_u8, idmask_u8, mode: str, f: int, n: int, heads: int, dim_head: int = 64,
                      want_cls_attn: bool = True, precision="bf16"):
    """qkv (B, 1+f*n, 3*heads*dim_head) with q pre-scaled -> (out (B,N,heads*dim_head), cls_attn (B*heads,N))"""
    T = _T(precision)
    _prep(qkv, T)
    B, N, _ = qkv.shape
    _lib.require_device(qkv.device)
    out = torch.empty((B, N, heads * dim_head), dtype=T, device=qkv.device)
    cls = torch.empty((B * heads, N), dtype=torch.float32, device=qkv.device) if want_cls_attn else None
    lib = _lib.load()
    ws_bytes = int(lib.mt_divided_attn_workspace_bytes(B, f, n, heads))
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=qkv.device)
    with torch.cuda.device(qkv.device):
        rc = lib.mt_divided_attn_fwd(_lib.prec_id(precision), qkv.data_ptr(), mask_u8.data_ptr(),
                                     _lib.ptr(idmask_u8), _lib.ATTN_TIME if mode == "time" else _lib.ATTN_SPACE,
                                     out.data_ptr(), _lib.ptr(cls), B, f, n, heads, dim_head, ws.data_ptr(), ws_bytes,
                                     _lib.stream_ptr())
    _lib.check(rc, "mt_divided_attn_fwd")
    return out, cls


def fused_attention(xn, w_qkv_heads, mask_u8, idmask_u8, mode: str, f: int, n: int, heads: int, dim_head: int = 64,
                    want_cls_attn: bool = True):
    """LayerNorm'd tokens xn (B, 1+f*n, dim) bf16 + per-head weights (heads, 3*dim_head, dim) bf16 ->
    (out (B,N,heads*dim_head) bf16, cls_attn (B*heads,N) float32): projection + divided attention in one kernel."""
    _prep(xn, torch.bfloat16)
    _prep(w_qkv_heads, torch.bfloat16)
    B, N, dim = xn.shape
    _lib.require_device(xn.device)
    out = torch.empty((B, N, heads * dim_head), dtype=torch.bfloat16, device=xn.device)
    cls = torch.empty((B * heads, N), dtype=torch.float32, device=xn.device) if want_cls_attn else None
    lib = _lib.load()
    ws_bytes = int(lib.mt_fused_attn_workspace_bytes(B, f, n, heads))
    ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=xn.device)
    with torch.cuda.device(xn.device):
        rc = lib.mt_fused_attn_fwd(xn.data_ptr(), w_qkv_heads.data_ptr(), mask_u8.data_ptr(), _lib.ptr(idmask_u8),
                                   _lib.ATTN_TIME if mode == "time" else _lib.ATTN_SPACE, out.data_ptr(), _lib.ptr(cls),
                                   B, f, n, heads, dim_head, dim, ws.data_ptr(), ws_bytes, _lib.stream_ptr())
    _lib.check(rc, "mt_fused_attn_fwd")
    return out, cls


def stem(x_nhwc, w27x32, shift, precision="bf16"):
    n, h, w, c = x_nhwc.shape
    assert c == 3
    _prep(x_nhwc)
    _lib.require_device(x_nhwc.device)
    out = torch.empty((n, (h + 1) // 2, (w + 1) // 2, 32), dtype=_T(precision), device=x_nhwc.device)
    with torch.cuda.device(x_nhwc.device):
        rc = _lib.load().mt_stem_fwd(_lib.prec_id(precision), x_nhwc.data_ptr(),
                                     _lib.IN_U8 if x_nhwc.dtype == torch.uint8 else _lib.IN_F32, w27x32.data_ptr(),
                                     shift.data_ptr(), out.data_ptr(), n, h, w, _lib.stream_ptr())
    _lib.check(rc, "mt_stem_fwd")
    return out


def dwconv(x_nhwc, w_taps, shift, k: int, s: int, precision="bf16"):
    """-> (out NHWC, pool_part (n, n_chunks, c) float32: per-chunk sums of out over output pixels)"""
    T = _T(precision)
    _prep(x_nhwc, T)
    n, h, w, c = x_nhwc.shape
    _lib.require_device(x_nhwc.device)
    lib = _lib.load()
    out = torch.empty((n, (h + s - 1) // s, (w + s - 1) // s, c), dtype=T, device=x_nhwc.device)
    pool = torch.full((n, lib.mt_dwconv_chunks(_lib.prec_id(precision), h, w, c, k, s), c), float("nan"), dtype=torch.float32,
                      device=x_nhwc.device)
    with torch.cuda.device(x_nhwc.device):
        rc = lib.mt_dwconv_fwd(_lib.prec_id(precision), x_nhwc.data_ptr(), w_taps.data_ptr(), shift.data_ptr(),
                               out.data_ptr(), pool.data_ptr(), n, h, w, c, k, s, _lib.stream_ptr())
    _lib.check(rc, "mt_dwconv_fwd")
    return out, pool


def expand_dwconv(x_nhwc, w_exp, exp_shift, w_taps, dw_shift, k: int, s: int):
    """Fused 1x1 expand + BN + swish + depthwise kxk + BN + swish (bf16).
    -> (out NHWC bf16, pool_part (n, n_chunks, cexp) float32)"""
    _prep(x_nhwc, torch.bfloat16); _prep(w_exp, torch.bfloat16)
    n, h, w, cin = x_nhwc.shape
    cexp = w_exp.shape[0]
    assert h == w
    _lib.require_device(x_nhwc.device)
    lib = _lib.load()
    chunks = lib.mt_expand_dwconv_chunks(h, cin, cexp, k, s)
    if chunks <= 0:
        raise RuntimeError(f"no fused expand+depthwise schedule for h={h} cin={cin} cexp={cexp} k={k} s={s}")
    ho = (h + s - 1) // s
    out = torch.empty((n, ho, ho, cexp), dtype=torch.bfloat16, device=x_nhwc.device)
    pool = torch.full((n, chunks, cexp), float("nan"), dtype=torch.float32, device=x_nhwc.device)
    with torch.cuda.device(x_nhwc.device):
        rc = lib.mt_expand_dwconv_fwd(x_nhwc.data_ptr(), w_exp.data_ptr(), exp_shift.data_ptr(), w_taps.data_ptr(),
                                      dw_shift.data_ptr(), out.data_ptr(), pool.data_ptr(), n, h, cin, cexp, k, s,
                                      _lib.stream_ptr())
    _lib.check(rc, "mt_expand_dwconv_fwd")
    return out, pool


def se_gate(pool_part, hw: int, wr, br, we, be):
    n, chunks, c = pool_part.shape
    sq = wr.shape[0]
    _lib.require_device(pool_part.device)
    gate = torch.empty((n, c), dtype=torch.float32, device=pool_part.device)
    with torch.cuda.device(pool_part.device):
        rc = _lib.load().mt_se_gate_fwd(pool_part.data_ptr(), chunks, hw, wr.data_ptr(), br.data_ptr(), we.data_ptr(),
                                        be.data_ptr(), gate.data_ptr(), n, c, sq, _lib.stream_ptr())
    _lib.check(rc, "mt_se_gate_fwd")
    return gate


def mbconv(x_nhwc, block_index: int, packed_effnet, precision="bf16"):
    """One B0 MBConv block (eval) on NHWC input, using the block's weights from weights.pack_effnet()."""
    T = _T(precision)
    _prep(x_nhwc, T)
    lib = _lib.load()
    spec = _lib.MBConvSpec()
    _lib.check(lib.mt_effnet_b0_block_spec(block_index, spec), "mt_effnet_b0_block_spec")
    n, h, w, c = x_nhwc.shape
    if (h, w, c) != (spec.hw_in, spec.hw_in, spec.cin):
        raise ValueError(f"block {block_index} expects (n,{spec.hw_in},{spec.hw_in},{spec.cin}), got {tuple(x_nhwc.shape)}")
    _lib.require_device(x_nhwc.device)
    ho = (spec.hw_in + spec.stride - 1) // spec.stride
    out = torch.empty((n, ho, ho, spec.cout), dtype=T, device=x_nhwc.device)
    prec = _lib.prec_id(precision)
    ws = torch.empty(lib.mt_mbconv_workspace_bytes(spec, n, prec), dtype=torch.uint8, device=x_nhwc.device)
    with torch.cuda.device(x_nhwc.device):
        rc = lib.mt_mbconv_fwd(prec, spec, packed_effnet.struct.blocks[block_index], x_nhwc.data_ptr(), out.data_ptr(), n,
                               ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "mt_mbconv_fwd")
    return out


def head(x, ln_g, ln_b, w, bias):
    B, tokens, dim = x.shape
    classes = w.shape[0]
    _lib.require_device(x.device)
    logits = torch.empty((B, classes), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().mt_head_fwd(x.data_ptr(), ln_g.data_ptr(), ln_b.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                     logits.data_ptr(), B, tokens, dim, classes, _lib.stream_ptr())
    _lib.check(rc, "mt_head_fwd")
    return logits


# ---------------------------------------------------------------------------------------------------
# Backward building blocks (csrc/train.cu; schedule in training.py)
# ---------------------------------------------------------------------------------------------------
def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def grad_prep(src, want_rm=False, want_t=False, want_colsum=False, rows_per_batch=0, m=None, precision="bf16"):
    """One pass over src [rows, C] (float32 or T): -> (cast copy T [m, C] | None, transpose T [C, mp] | None with the
    columns m..mp-1 zero, column sums float32 [C] | None).  rows_per_batch > 0 drops the CLS row of every video
    (source row = i + i // rows_per_batch + 1; pass m = the number of kept rows)."""
    T = _T(precision)
    _prep(src)
    if src.dtype not in (torch.float32, T):
        raise TypeError(f"grad_prep: src must be float32 or {T}, got {src.dtype}")
    rows, c = src.shape
    m = rows if m is None else m
    mp = round_up(m, 64)
    dev = src.device
    _lib.require_device(dev)
    lib = _lib.load()
    out_rm = torch.empty((m, c), dtype=T, device=dev) if want_rm else None
    out_t = torch.empty((c, mp), dtype=T, device=dev) if want_t else None
    cs = torch.empty((c,), dtype=torch.float32, device=dev) if want_colsum else None
    ws_bytes = int(lib.mt_grad_prep_workspace_bytes(mp, c)) if want_colsum else 0
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mt_grad_prep(_lib.prec_id(precision), src.data_ptr(), int(src.dtype == torch.float32), _lib.ptr(out_rm),
                              _lib.ptr(out_t), _lib.ptr(cs), m, c, mp, rows_per_batch, ws.data_ptr(), ws_bytes,
                              _lib.stream_ptr())
    _lib.check(rc, "mt_grad_prep")
    return out_rm, out_t, cs


def layernorm_bwd_(gx, x, gamma, dy, precision="bf16"):
    """gx (float32 [rows, dim], in place) += LayerNorm backward of dy; returns (dgamma, dbeta)."""
    _prep(x, torch.float32); _prep(gx, torch.float32); _prep(dy, _T(precision))
    rows, dim = x.shape
    dev = x.device
    _lib.require_device(dev)
    lib = _lib.load()
    ws_bytes = int(lib.mt_layernorm_bwd_workspace_bytes(rows, dim))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    dgb = torch.empty((2 * dim,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mt_layernorm_bwd(_lib.prec_id(precision), x.data_ptr(), gamma.data_ptr(), dy.data_ptr(), gx.data_ptr(),
                                  dgb.data_ptr(), rows, dim, ws.data_ptr(), ws_bytes, _lib.stream_ptr())
    _lib.check(rc, "mt_layernorm_bwd")
    return dgb[:dim], dgb[dim:]


def geglu(h, precision="bf16"):
    """h [m, 2*n_out] in the interleaved column order of the packed net.0 weight -> u * gelu(g)  [m, n_out]"""
    T = _T(precision)
    _prep(h, T)
    m, n2 = h.shape
    _lib.require_device(h.device)
    out = torch.empty((m, n2 // 2), dtype=T, device=h.device)
    with torch.cuda.device(h.device):
        rc = _lib.load().mt_geglu_fwd(_lib.prec_id(precision), h.data_ptr(), out.data_ptr(), m, n2 // 2, _lib.stream_ptr())
    _lib.check(rc, "mt_geglu_fwd")
    return out


def geglu_bwd(h, dout, precision="bf16"):
    T = _T(precision)
    _prep(h, T); _prep(dout, T)
    m, n2 = h.shape
    _lib.require_device(h.device)
    dh = torch.empty_like(h)
    with torch.cuda.device(h.device):
        rc = _lib.load().mt_geglu_bwd(_lib.prec_id(precision), h.data_ptr(), dout.data_ptr(), dh.data_ptr(), m, n2 // 2,
                                      _lib.stream_ptr())
    _lib.check(rc, "mt_geglu_bwd")
    return dh


def geglu_bwd_colsum(h, dout, precision="bf16"):
    """(dh, column sums of dh as float32 [2*n_out]) in one pass -- the data and the bias gradient of net.0"""
    T = _T(precision)
    _prep(h, T); _prep(dout, T)
    m, n2 = h.shape
    _lib.require_device(h.device)
    lib = _lib.load()
    dh = torch.empty_like(h)
    cs = torch.empty((n2,), dtype=torch.float32, device=h.device)
    ws = torch.empty((lib.mt_geglu_bwd_colsum_workspace_bytes(m, n2 // 2),), dtype=torch.uint8, device=h.device)
    with torch.cuda.device(h.device):
        rc = lib.mt_geglu_bwd_colsum(_lib.prec_id(precision), h.data_ptr(), dout.data_ptr(), dh.data_ptr(), cs.data_ptr(), m,
                                     n2 // 2, ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "mt_geglu_bwd_colsum")
    return dh, cs


def divided_attention_bwd(qkv, dout, mask_u8, idmask_u8, mode: str, f: int, n: int, heads: int, dim_head: int = 64,
                          precision="bf16"):
    """Backward of divided_attention: (qkv, d out) -> d qkv, same shape as qkv."""
    T = _T(precision)
    _prep(qkv, T); _prep(dout, T)
    B, N, _ = qkv.shape
    dev = qkv.device
    _lib.require_device(dev)
    lib = _lib.load()
    dqkv = torch.empty_like(qkv)
    ws_bytes = int(lib.mt_divided_attn_bwd_workspace_bytes(B, f, n, heads))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mt_divided_attn_bwd(_lib.prec_id(precision), qkv.data_ptr(), dout.data_ptr(), mask_u8.data_ptr(),
                                     _lib.ptr(idmask_u8), _lib.ATTN_TIME if mode == "time" else _lib.ATTN_SPACE,
                                     dqkv.data_ptr(), B, f, n, heads, dim_head, ws.data_ptr(), ws_bytes,
                                     _lib.stream_ptr())
    _lib.check(rc, "mt_divided_attn_bwd")
    return dqkv


def embed_bwd(g0, positions, size_embedding, table_rows: int, f: int, n: int, want_pos=True, want_size=True):
    """g0 float32 [B, 1+f*n, dim] -> (d pos_emb.weight | None, d size_emb.weight | None, d cls_token [dim])"""
    _prep(g0, torch.float32)
    B, N, dim = g0.shape
    dev = g0.device
    _lib.require_device(dev)
    dpos = torch.zeros((table_rows, dim), dtype=torch.float32, device=dev) if want_pos else None
    dsize = torch.zeros((table_rows, dim), dtype=torch.float32, device=dev) if want_size else None
    dcls = torch.zeros((dim,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().mt_embed_bwd(g0.data_ptr(), _lib.ptr(positions), _lib.ptr(size_embedding), _lib.ptr(dpos),
                                      _lib.ptr(dsize), dcls.data_ptr(), B, f, n, dim, table_rows, _lib.stream_ptr())
    _lib.check(rc, "mt_embed_bwd")
    return dpos, dsize, dcls


def head_bwd_(gx, x, ln_g, ln_b, w, dlogits):
    """gx[:, 0, :] = d x[:, 0] (assigned); returns (dW [classes, dim], dbias [classes], dgamma [dim], dbeta [dim])"""
    _prep(x, torch.float32); _prep(gx, torch.float32); _prep(dlogits, torch.float32)
    B, tokens, dim = x.shape
    classes = w.shape[0]
    dev = x.device
    _lib.require_device(dev)
    lib = _lib.load()
    ws_bytes = int(lib.mt_head_bwd_workspace_bytes(B, dim, classes))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    grads = torch.empty((classes * dim + classes + 2 * dim,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mt_head_bwd(x.data_ptr(), ln_g.data_ptr(), ln_b.data_ptr(), w.data_ptr(), dlogits.data_ptr(),
                             gx.data_ptr(), grads.data_ptr(), B, tokens, dim, classes, ws.data_ptr(), ws_bytes,
                             _lib.stream_ptr())
    _lib.check(rc, "mt_head_bwd")
    o = classes * dim
    return grads[:o].view(classes, dim), grads[o:o + classes], grads[o + classes:o + classes + dim], grads[o + classes + dim:]


def linear_wgrad_(dw, dy_t, x_t, precision="bf16"):
    """dw (float32 [n_out, k_in], in place) += dy_t [n_out, mp] @ x_t [k_in, mp].T   (split-K tensor-core GEMM)"""
    T = _T(precision)
    _prep(dy_t, T); _prep(x_t, T); _prep(dw, torch.float32)
    n_out, mp = dy_t.shape
    k_in = x_t.shape[0]
    if x_t.shape[1] != mp or tuple(dw.shape) != (n_out, k_in):
        raise ValueError("linear_wgrad_: shape mismatch")
    _lib.require_device(dw.device)
    with torch.cuda.device(dw.device):
        rc = _lib.load().mt_linear_wgrad(_lib.prec_id(precision), dy_t.data_ptr(), x_t.data_ptr(), dw.data_ptr(), n_out, k_in,
                                         mp, _lib.stream_ptr())
    _lib.check(rc, "mt_linear_wgrad")
    return dw


def linear_wgrad_nt_(dw, dy, x):
    """dw (float32 [n_out, k_in], in place) += dy [m, n_out].T @ x [m, k_in], bf16 row-major operands as they are
    (MN-major tcgen05 operands: no transposed copies); n_out and k_in multiples of 8"""
    _prep(dy, torch.bfloat16); _prep(x, torch.bfloat16); _prep(dw, torch.float32)
    m, n_out = dy.shape
    k_in = x.shape[1]
    if x.shape[0] != m or tuple(dw.shape) != (n_out, k_in):
        raise ValueError("linear_wgrad_nt_: shape mismatch")
    _lib.require_device(dw.device)
    with torch.cuda.device(dw.device):
        rc = _lib.load().mt_linear_wgrad_nt(_lib.prec_id("bf16"), dy.data_ptr(), x.data_ptr(), dw.data_ptr(), n_out, k_in, m,
                                            _lib.stream_ptr())
    _lib.check(rc, "mt_linear_wgrad_nt")
    return dw
