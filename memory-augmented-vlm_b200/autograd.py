"""torch.autograd.Function wrappers: forward AND backward of every op run in libmavlm.so.

The reference trains this path with plain PyTorch autograd over ATen ops (train.py:1694-1728: the
recurrent memory transformer, the fuser and token_type_embedding are trainable; frame features are
detached, llava_arch.py:302,481).  Here autograd only provides the graph / BPTT bookkeeping
(save_for_backward, gradient accumulation across chunks); the arithmetic of each backward is a C-ABI
call (gemm_ex dgrad/wgrad, colsum, layernorm_bwd, act_bwd, xattn_bwd).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_RELU
from ._device import on_tensor_device

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}   # fp16: inference-only GEMM layouts


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _raw():
    from . import ops
    return ops


@on_tensor_device
def gemm_ex(a: torch.Tensor, trans_a: bool, b: torch.Tensor, trans_b: bool, m: int, n: int, k: int, *,
            out: Optional[torch.Tensor] = None, alpha: float = 1.0, accumulate: bool = False,
            out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """C[M,N] (+)= alpha * op(A) op(B) on 2-D row-major operands (see include/mavlm.h: mavlm_gemm_ex)."""
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype or a.dtype, device=a.device)
    st = _lib.load().mavlm_gemm_ex(_p(a), a.stride(0), int(trans_a), _p(b), b.stride(0), int(trans_b), _p(out),
                                   out.stride(0), m, n, k, float(alpha), int(accumulate), 1, 1, None, _DT[a.dtype],
                                   _DT[out.dtype], _s())
    _lib.check(st, "gemm_ex")
    return out


@on_tensor_device
def colsum(x2: torch.Tensor) -> torch.Tensor:
    """fp32 column sums of a 2-D row-major tensor."""
    out = torch.empty(x2.shape[1], dtype=torch.float32, device=x2.device)
    st = _lib.load().mavlm_colsum(_p(x2), x2.stride(0), _p(out), x2.shape[0], x2.shape[1], 0, _DT[x2.dtype], _s())
    _lib.check(st, "colsum")
    return out


@on_tensor_device
def _act_bwd(dy: torch.Tensor, ref: torch.Tensor, act: int) -> torch.Tensor:
    dx = torch.empty_like(dy)
    st = _lib.load().mavlm_act_bwd(_p(dy), _p(ref), _p(dx), dy.numel(), act, _DT[dy.dtype], _s())
    _lib.check(st, "act_bwd")
    return dx


def _to(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return t if t.dtype == dtype else _raw().cast(t.contiguous(), dtype)


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) (+ resid) (+ addvec); act in {none, relu} (GELU is kept unfused when training)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, x, w, b, resid, addvec, act, out_dtype):
        ops = _raw()
        y = ops._linear_raw(x, w, b, act=act, resid=resid, addvec=addvec, out_dtype=out_dtype)
        if act == ACT_RELU and (resid is not None or addvec is not None):
            raise RuntimeError("mavlm: fused ReLU with residual is not differentiable here (mask needs the bare output)")
        ctx.save_for_backward(x, w, y if act == ACT_RELU else None)
        ctx.act = act
        ctx.has = (b is not None, resid is not None, addvec is not None)
        ctx.resid_dtype = None if resid is None else resid.dtype
        ctx.resid_shape = None if resid is None else resid.shape
        return y

    @staticmethod
    @on_tensor_device
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        n, k = w.shape
        cdt = x.dtype
        dy2 = _to(dy.reshape(-1, n).contiguous(), cdt)
        if ctx.act == ACT_RELU:
            dy2 = _act_bwd(dy2, y.reshape(-1, n), ACT_RELU)
        m = dy2.shape[0]
        x2 = x.reshape(-1, k)
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        dx = dw = db = dres = dav = None
        if ctx.needs_input_grad[0]:
            dx = gemm_ex(dy2, False, w, False, m, k, n).reshape(x.shape)          # dX = dY W
        if ctx.needs_input_grad[1]:
            dw = gemm_ex(dy2, True, x2, False, n, k, m)                           # dW = dY^T X
        has_b, has_r, has_a = ctx.has
        if (has_b and ctx.needs_input_grad[2]) or (has_a and ctx.needs_input_grad[4]):
            cs = colsum(dy2)
            if has_b and ctx.needs_input_grad[2]:
                db = _to(cs, cdt)
            if has_a and ctx.needs_input_grad[4]:
                dav = _to(cs, cdt)
        if has_r and ctx.needs_input_grad[3]:
            dres = _to(dy.reshape(-1, n).contiguous(), ctx.resid_dtype).reshape(ctx.resid_shape)
        return dx, dw, db, dres, dav, None, None


class ActFn(torch.autograd.Function):
    """Unfused activation (training keeps the GELU pre-activation)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, x, act):
        xc = x.contiguous()
        y = torch.empty_like(xc)
        st = _lib.load().mavlm_act_fwd(_p(xc), _p(y), xc.numel(), act, _DT[xc.dtype], _s())
        _lib.check(st, "act_fwd")
        ctx.save_for_backward(xc if act == ACT_GELU_ERF else y)
        ctx.act = act
        return y

    @staticmethod
    @on_tensor_device
    def backward(ctx, dy):
        (ref,) = ctx.saved_tensors
        return _act_bwd(dy.contiguous(), ref, ctx.act), None


class LayerNormFn(torch.autograd.Function):
    """LayerNorm of the fp32 pre-LN sum."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, pre, gamma, beta, eps, out_dtype):
        y = _raw()._layernorm_raw(pre, gamma, beta, eps, out_dtype)
        ctx.save_for_backward(pre, gamma)
        ctx.eps = eps
        return y

    @staticmethod
    @on_tensor_device
    def backward(ctx, dy):
        pre, gamma = ctx.saved_tensors
        d = pre.shape[-1]
        pre2 = _to(pre.reshape(-1, d).contiguous(), torch.float32)
        dy2 = _to(dy.reshape(-1, d).contiguous(), gamma.dtype)
        dpre = torch.empty_like(pre2)
        dg = torch.zeros(d, dtype=torch.float32, device=pre.device)
        db = torch.zeros(d, dtype=torch.float32, device=pre.device)
        st = _lib.load().mavlm_layernorm_bwd(_p(pre2), _p(gamma), _p(dy2), _p(dpre), _p(dg), _p(db), pre2.shape[0], d,
                                             float(ctx.eps), _DT[gamma.dtype], _s())
        _lib.check(st, "layernorm_bwd")
        return _to(dpre, pre.dtype).reshape(pre.shape), _to(dg, gamma.dtype), _to(db, gamma.dtype), None, None


class XAttnFn(torch.autograd.Function):
    """softmax(q k^T * scale) v per head, backward from the saved log-sum-exp."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, q, k, v, heads, head_dim, scale):
        o, lse, _ = _raw()._xattn_raw(q, k, v, heads, head_dim=head_dim, scale=scale, want_lse=True)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.cfg = (heads, head_dim, scale)
        return o

    @staticmethod
    @on_tensor_device
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        heads, dh, scale = ctx.cfg
        b, lq, _ = q.shape
        lk = k.shape[1]
        do = do.contiguous()
        dq = torch.empty(q.shape, dtype=q.dtype, device=q.device)
        dk = torch.empty(k.shape, dtype=k.dtype, device=k.device)
        dv = torch.empty(v.shape, dtype=v.dtype, device=v.device)
        lib = _lib.load()
        code = _DT[q.dtype]
        nbytes = lib.mavlm_xattn_bwd_workspace_bytes(b, heads, lq, lk, dh, code)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
        st = lib.mavlm_xattn_bwd(_p(q), q.stride(1), q.stride(0), _p(k), k.stride(1), k.stride(0), _p(v), v.stride(1),
                                 v.stride(0), _p(o), o.stride(1), o.stride(0), _p(do), do.stride(1), do.stride(0),
                                 _p(lse), _p(dq), dq.stride(1), dq.stride(0), _p(dk), dk.stride(1), dk.stride(0), _p(dv),
                                 dv.stride(1), dv.stride(0), b, heads, lq, lk, dh, float(scale), code, _p(ws), nbytes,
                                 _s())
        _lib.check(st, "xattn_bwd")
        return dq, dk, dv, None, None, None


class XAttnKVFn(torch.autograd.Function):
    """XAttnFn on a fused (k | v) projection buffer [B, Lk, 2*H*dh]: the backward writes dK and dV straight into
    the two column halves of ONE gradient buffer (column-slice views of `kv` would each get a zero-filled full-size
    gradient from autograd's SliceBackward: 720 MB of fills and copies per layer at OV-7B, batch 8)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, q, kv, heads, head_dim, scale):
        hd = heads * head_dim
        o, lse, _ = _raw()._xattn_raw(q, kv[..., :hd], kv[..., hd:], heads, head_dim=head_dim, scale=scale, want_lse=True)
        ctx.save_for_backward(q, kv, o, lse)
        ctx.cfg = (heads, head_dim, scale)
        return o

    @staticmethod
    @on_tensor_device
    def backward(ctx, do):
        q, kv, o, lse = ctx.saved_tensors
        heads, dh, scale = ctx.cfg
        hd = heads * dh
        k, v = kv[..., :hd], kv[..., hd:]
        b, lq, _ = q.shape
        lk = kv.shape[1]
        do = do.contiguous()
        dq = torch.empty(q.shape, dtype=q.dtype, device=q.device)
        dkv = torch.empty(kv.shape, dtype=kv.dtype, device=kv.device)
        dk, dv = dkv[..., :hd], dkv[..., hd:]
        lib = _lib.load()
        code = _DT[q.dtype]
        nbytes = lib.mavlm_xattn_bwd_workspace_bytes(b, heads, lq, lk, dh, code)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
        st = lib.mavlm_xattn_bwd(_p(q), q.stride(1), q.stride(0), _p(k), k.stride(1), k.stride(0), _p(v), v.stride(1),
                                 v.stride(0), _p(o), o.stride(1), o.stride(0), _p(do), do.stride(1), do.stride(0),
                                 _p(lse), _p(dq), dq.stride(1), dq.stride(0), _p(dk), dk.stride(1), dk.stride(0), _p(dv),
                                 dv.stride(1), dv.stride(0), b, heads, lq, lk, dh, float(scale), code, _p(ws), nbytes,
                                 _s())
        _lib.check(st, "xattn_bwd")
        return dq, dkv, None, None, None


class AddRowsFn(torch.autograd.Function):
    """y[t, n, :] = x[t, n, :] + table[t, :]   (initial_memory + memory_pos_embed; x + type embedding with T = 1)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, x, table):
        idx = torch.arange(x.shape[0], device=x.device)
        y = _raw()._add_pe_raw(x.contiguous(), table.float().contiguous(), idx)
        ctx.tdtype = table.dtype
        return y

    @staticmethod
    @on_tensor_device
    def backward(ctx, dy):
        dy = dy.contiguous()
        dt = None
        if ctx.needs_input_grad[1]:
            dt = torch.stack([colsum(dy[t]) for t in range(dy.shape[0])], dim=0).to(ctx.tdtype)
        return (dy if ctx.needs_input_grad[0] else None), dt


class AssembleFn(torch.autograd.Function):
    """Token assembly of memory_forward_train (llava_arch.py:548-554, 620-629, 708-731) for a batch of videos with the
    assemble kernel instead of torch.cat: prompt rows | memory tokens | newline | prompt rows | fine frames + type
    embedding 1 | newline.  `memtok` already carries type embedding 0 (added by the fuser GEMM).  Gradients: slices
    of d(seq) for the memory tokens, column sums for the type embedding / newline, batch sums for the prompt rows
    (pm_emb / pf_emb are the embedding lookups; their values are re-gathered from the table by the kernel)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, memtok, z, fine_idx, emb_w, newline, embed_table, pm_ids, pf_ids, pm_emb, pf_emb, drop_frames):
        ops = _raw()
        b, n_mem, d = memtok.shape
        p = z.shape[2]
        nf = fine_idx.numel()
        n_pm, n_pf = pm_ids.numel(), pf_ids.numel()
        length = n_pm + n_mem + 1 + (0 if drop_frames else n_pf + nf * p + 1)
        dt = memtok.dtype
        seq = torch.empty((b, length, d), dtype=dt, device=memtok.device)
        tab = torch.zeros((2, d), dtype=dt, device=memtok.device)
        tab[1].copy_(emb_w[1])
        nl = newline.detach().to(dt).contiguous()
        table = embed_table.detach()
        table = table if table.dtype == dt and table.is_contiguous() else table.to(dt).contiguous()
        mt, zz = memtok.detach().contiguous(), z.detach().contiguous()
        for i in range(b):
            ops.assemble(seq[i], mt[i], n_mem, zz[i], fine_idx, p, tab, nl, table, pm_ids, pf_ids, drop_frames=drop_frames)
        ctx.meta = (n_pm, n_mem, n_pf, nf * p, bool(drop_frames), emb_w.dtype, newline.dtype, pm_emb.dtype)
        return seq

    @staticmethod
    @on_tensor_device
    def backward(ctx, dseq):
        n_pm, n_mem, n_pf, n_fine, drop, emb_dt, nl_dt, pe_dt = ctx.meta
        b, _, d = dseq.shape
        r_nl1 = n_pm + n_mem
        r_pf = r_nl1 + 1
        r_fine = r_pf + n_pf
        need = ctx.needs_input_grad
        d_mem = dseq[:, n_pm:r_nl1] if need[0] else None
        d_emb = d_nl = d_pm = d_pf = None
        if need[3]:
            d_emb = torch.zeros((2, d), dtype=torch.float32, device=dseq.device)
            if not drop:
                for i in range(b):
                    d_emb[1] += colsum(dseq[i, r_fine:r_fine + n_fine])
            d_emb = d_emb.to(emb_dt)
        if need[4]:
            d_nl = dseq[:, r_nl1].float().sum(0)
            if not drop:
                d_nl = d_nl + dseq[:, r_fine + n_fine].float().sum(0)
            d_nl = d_nl.to(nl_dt)
        if need[8]:
            d_pm = dseq[:, :n_pm].float().sum(0).to(pe_dt)
        if need[9]:
            d_pf = (dseq[:, r_pf:r_fine].float().sum(0) if not drop else torch.zeros((n_pf, d), device=dseq.device)).to(pe_dt)
        return d_mem, None, None, d_emb, d_nl, None, None, None, d_pm, d_pf, None


def linear(x, w, b, act, resid, addvec, out_dtype):
    if act == ACT_GELU_ERF:                      # keep the pre-activation: GEMM, then the unfused activation
        if resid is not None or addvec is not None:
            raise RuntimeError("mavlm: GELU with residual is not used on this path")
        pre = LinearFn.apply(x, w, b, None, None, ACT_NONE, None)
        y = ActFn.apply(pre, ACT_GELU_ERF)
        return y if out_dtype in (None, y.dtype) else _raw().cast(y, out_dtype)
    return LinearFn.apply(x, w, b, resid, addvec, act, out_dtype)
