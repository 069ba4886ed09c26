"""Tensor-level wrappers over the C ABI: validate, allocate outputs with torch (PyTorch owns all
storage), pass raw pointers + the current CUDA stream.  PyTorch is plumbing here (device memory,
streams); every computation below runs in libmavlm.so.  No fallback: CPU tensors are an error.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_RELU, BF16, F16, F32
from ._device import on_tensor_device

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}
_device_checked = set()


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f"mavlm supports float32, bfloat16 and float16 tensors, got {t.dtype}") from None


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mavlm: tensor is not on a CUDA device; this path has no CPU fallback")
        dev = t.device.index
        if dev not in _device_checked:
            _lib.check(_lib.load().mavlm_check_device(dev), "check_device")
            _device_checked.add(dev)


def _rowmajor2d(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"mavlm: {name} must be 2-D with unit inner stride, got shape {tuple(t.shape)} "
                           f"strides {t.stride()}")
    return t


def _grad_needed(*ts: Optional[torch.Tensor]) -> bool:
    need = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)
    if need and any(t is not None and t.dtype == torch.float16 for t in ts):
        raise RuntimeError("mavlm: float16 is an inference dtype on this path (the reference trains in bfloat16, "
                           "finetune_short.sh:71); use torch.no_grad() or bfloat16 / float32 parameters")
    return need


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = ACT_NONE,
           resid: Optional[torch.Tensor] = None, addvec: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """y = act(x @ weight.T + bias) (+ resid) (+ addvec);  x [..., K], weight [N, K] (nn.Linear layout).
    Differentiable (autograd.py) when an operand requires grad."""
    if _grad_needed(x, weight, bias, resid, addvec):
        if out is not None:
            raise RuntimeError("mavlm.linear: out= cannot be combined with autograd")
        from . import autograd as ag
        return ag.linear(x, weight, bias, act, resid, addvec, out_dtype)
    return _linear_raw(x, weight, bias, act=act, resid=resid, addvec=addvec, out=out, out_dtype=out_dtype)


@on_tensor_device
def _linear_raw(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = ACT_NONE,
                resid: Optional[torch.Tensor] = None, addvec: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    _need_cuda(x, weight, bias, resid, addvec, out)
    lead = x.shape[:-1]
    k = x.shape[-1]
    x2 = x.reshape(-1, k)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    n = weight.shape[0]
    if weight.shape[1] != k:
        raise RuntimeError(f"mavlm.linear: weight {tuple(weight.shape)} does not match input features {k}")
    _rowmajor2d(weight, "weight")
    m = x2.shape[0]
    odt = out_dtype or x.dtype
    r2 = None
    if resid is not None:
        r2 = resid.reshape(-1, n)
        if r2.shape[0] != m or r2.stride(1) != 1:
            raise RuntimeError("mavlm.linear: resid shape mismatch")
    if out is None:
        out2 = torch.empty((m, n), dtype=odt, device=x.device)
    else:
        out2 = out.reshape(-1, n) if out.dim() != 2 else out
        if out2.shape[0] != m or out2.stride(1) != 1 or out2.data_ptr() != out.data_ptr():
            raise RuntimeError("mavlm.linear: out must be a row-major [M, N] view")
        odt = out2.dtype
    for t in (weight, bias, resid, addvec):
        if t is not None and t.dtype != x.dtype:
            raise TypeError(f"mavlm.linear: operand dtype {t.dtype} != input dtype {x.dtype}")
    lib = _lib.load()
    st = lib.mavlm_gemm_bias_act_fwd(_ptr(x2), x2.stride(0), _ptr(weight), weight.stride(0), _ptr(bias), _ptr(r2),
                                     0 if r2 is None else r2.stride(0), _ptr(addvec), _ptr(out2), out2.stride(0), m, n,
                                     k, act, dtype_code(x), _DTYPES[odt], _stream())
    _lib.check(st, "gemm_bias_act_fwd")
    if out is not None:
        return out
    return out2.reshape(*lead, n)


class GemmWork:
    """One nn.Linear call y = act(x W^T + b) (+ resid) (+ addvec) as a TILED problem (256 x 256 output tiles in a fixed
    order) that can be computed in pieces: as the FILLER of critical-path GEMMs (`linear_fill`) and / or by `run()`
    for whatever is left before the first consumer reads y.  bf16 / fp16 tier.  Holds references to its operands."""

    def __init__(self, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, *,
                 act: int = ACT_NONE, resid: Optional[torch.Tensor] = None, addvec: Optional[torch.Tensor] = None,
                 after: Optional["GemmWork"] = None):
        _need_cuda(x, weight, bias, resid, addvec, out)
        if x.dtype not in (torch.bfloat16, torch.float16):
            raise TypeError("mavlm.GemmWork: tensor-core tier only (bfloat16 / float16)")
        for t in (weight, bias, resid, addvec):
            if t is not None and t.dtype != x.dtype:
                raise TypeError(f"mavlm.GemmWork: operand dtype {t.dtype} != input dtype {x.dtype}")
        _rowmajor2d(x, "x")
        _rowmajor2d(weight, "weight")
        _rowmajor2d(out, "out")
        m, k = x.shape
        n = weight.shape[0]
        if weight.shape[1] != k or tuple(out.shape) != (m, n) or out.dtype not in (x.dtype, torch.float32):
            raise RuntimeError("mavlm.GemmWork: shape / dtype mismatch")
        if resid is not None:
            _rowmajor2d(resid, "resid")
            if tuple(resid.shape) != (m, n):
                raise RuntimeError("mavlm.GemmWork: resid shape mismatch")
        self.keep = (x, weight, bias, resid, addvec, out)
        self.code = dtype_code(x)
        d = _lib.GemmDesc()
        d.A, d.lda, d.W, d.ldw = x.data_ptr(), x.stride(0), weight.data_ptr(), weight.stride(0)
        d.bias = _ptr(bias)
        d.resid, d.ldr = _ptr(resid), (0 if resid is None else resid.stride(0))
        d.addvec = _ptr(addvec)
        d.pe_table, d.frame_idx, d.tokens_per_frame = None, None, 0
        d.C, d.ldc = out.data_ptr(), out.stride(0)
        d.M, d.N, d.K, d.act, d.out_dtype = m, n, k, act, _DTYPES[out.dtype]
        self.desc = d
        self.out = out
        self.device = x.device
        self.after = after                       # a work whose output this one reads: it must be complete first
        self.total = _lib.load().mavlm_gemm_num_tiles(ctypes.byref(d))
        if self.total <= 0:
            raise _lib.MavlmError("mavlm.GemmWork: " + _lib.load().mavlm_last_error_string().decode())
        self.cursor = 0

    @property
    def done(self) -> bool:
        return self.cursor >= self.total

    @property
    def ready(self) -> bool:
        return self.after is None or self.after.done

    def run(self, fillers=()) -> torch.Tensor:
        """Compute every tile that is still missing and return the output: a plain launch over the missing range, or --
        when nothing of this problem has been computed yet -- a launch that takes a ready filler along."""
        if self.after is not None and not self.after.done:
            self.after.run()
        if self.done:
            return self.out
        lib = _lib.load()
        with torch.cuda.device(self.device):
            other = _pick_filler(fillers, self.desc.K, exclude=self) if self.cursor == 0 else None
            if other is not None:
                done_end = ctypes.c_int(other.cursor)
                st = lib.mavlm_gemm_fill_fwd(ctypes.byref(self.desc), ctypes.byref(other.desc), other.cursor, other.total,
                                             ctypes.byref(done_end), self.code, _stream())
                _lib.check(st, "gemm_fill_fwd")
                other.cursor = done_end.value
            else:
                st = lib.mavlm_gemm_tiles_fwd(ctypes.byref(self.desc), self.cursor, self.total, self.code, _stream())
                _lib.check(st, "gemm_tiles_fwd")
        self.cursor = self.total
        return self.out


def _pick_filler(fillers, k_primary: int, exclude=None):
    """First filler that is unfinished, whose input is complete, and whose tiles are not longer than the primary's
    (a longer tile in the last wave would stretch the critical path): K_fill <= 1.25 K_primary."""
    for w_ in fillers:
        if w_ is exclude or w_.done or not w_.ready:
            continue
        if w_.desc.K * 4 > k_primary * 5:
            continue
        return w_
    return None


def linear_fill(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = ACT_NONE,
                resid: Optional[torch.Tensor] = None, addvec: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, out_dtype: Optional[torch.dtype] = None,
                fillers=()) -> torch.Tensor:
    """`linear` on the critical path with TAIL FILL: the tile slots its last wave would leave idle compute tiles of the
    first unfinished, ready GemmWork in `fillers` (include/mavlm.h: mavlm_gemm_fill_fwd).  Without a usable filler (or
    in the fp32 tier) this is `linear`."""
    work = _pick_filler(fillers, x.shape[-1]) if x.dtype in (torch.bfloat16, torch.float16) else None
    if work is None:
        return linear(x, weight, bias, act=act, resid=resid, addvec=addvec, out=out, out_dtype=out_dtype)
    lead = x.shape[:-1]
    k = x.shape[-1]
    x2 = x.reshape(-1, k)
    m, n = x2.shape[0], weight.shape[0]
    if out is None:
        out2 = torch.empty((m, n), dtype=out_dtype or x.dtype, device=x.device)
    else:
        out2 = out.reshape(-1, n) if out.dim() != 2 else out
        if out2.data_ptr() != out.data_ptr():
            raise RuntimeError("mavlm.linear_fill: out must be a row-major [M, N] view")
    prim = GemmWork(x2, weight, bias, out2, act=act, resid=None if resid is None else resid.reshape(-1, n), addvec=addvec)
    done_end = ctypes.c_int(work.cursor)
    with torch.cuda.device(x.device):
        st = _lib.load().mavlm_gemm_fill_fwd(ctypes.byref(prim.desc), ctypes.byref(work.desc), work.cursor, work.total,
                                             ctypes.byref(done_end), prim.code, _stream())
    _lib.check(st, "gemm_fill_fwd")
    work.cursor = done_end.value
    if out is not None:
        return out
    return out2.reshape(*lead, n)


@on_tensor_device
def linear_pe(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, pe_table: torch.Tensor,
              frame_idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T, N, K] -> x @ weight.T + bias + pe_table[frame_idx][:, None, :] in one launch (bf16 / fp16 tier):
    the projector's second layer with the temporal PE added in the GEMM epilogue.  `out`: a contiguous [T, N, C]
    destination (e.g. this rank's slot of an all-gather buffer: no staging copy before the collective)."""
    _need_cuda(x, weight, bias, pe_table, frame_idx, out)
    t, n_tok, k = x.shape
    n = weight.shape[0]
    if out is not None and (tuple(out.shape) != (t, n_tok, n) or not out.is_contiguous() or out.dtype != x.dtype):
        raise RuntimeError("mavlm.linear_pe: out must be a contiguous [T, N, C] tensor of the input dtype")
    if x.dtype == torch.float32:
        return add_pe(linear(x, weight, bias), pe_table, frame_idx, out=out)
    x2 = x.reshape(t * n_tok, k)
    if x2.stride(1) != 1 or not x2.is_contiguous():
        x2 = x2.contiguous()
    _rowmajor2d(weight, "weight")
    if pe_table.dtype != torch.float32 or not pe_table.is_contiguous() or pe_table.shape[1] != n:
        raise RuntimeError("mavlm.linear_pe: pe_table must be contiguous fp32 [max_frames, N]")
    frame_idx = frame_idx.to(device=x.device, dtype=torch.int64).contiguous()
    if out is None:
        out = torch.empty((t, n_tok, n), dtype=x.dtype, device=x.device)
    st = _lib.load().mavlm_gemm_bias_pe_fwd(_ptr(x2), x2.stride(0), _ptr(weight), weight.stride(0), _ptr(bias),
                                            _ptr(pe_table), _ptr(frame_idx), n_tok, _ptr(out), n, t * n_tok, n, k,
                                            dtype_code(x), _stream())
    _lib.check(st, "gemm_bias_pe_fwd")
    return out


@on_tensor_device
def cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """x.to(dtype) for fp32 <-> bf16 / fp16 (the hand-off between the tensor-core tier and the fp32 tier)."""
    if x.dtype == dtype:
        return x
    _need_cuda(x)
    x = x.contiguous()
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    st = _lib.load().mavlm_cast_fwd(_ptr(x), _ptr(y), x.numel(), dtype_code(x), _DTYPES[dtype], _stream())
    _lib.check(st, "cast_fwd")
    return y


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if _grad_needed(x, gamma, beta):
        if out is not None:
            raise RuntimeError("mavlm.layernorm: out= cannot be combined with autograd")
        from . import autograd as ag
        return ag.LayerNormFn.apply(x, gamma, beta, eps, out_dtype)
    return _layernorm_raw(x, gamma, beta, eps, out_dtype, out)


@on_tensor_device
def _layernorm_raw(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
                   out_dtype: Optional[torch.dtype] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`out`: a contiguous tensor with x's element count (e.g. a ring-buffer slot) that receives the result."""
    _need_cuda(x, gamma, beta, out)
    d = x.shape[-1]
    x2 = x.reshape(-1, d)
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    odt = out_dtype or gamma.dtype
    if out is None:
        y = torch.empty(x2.shape, dtype=odt, device=x.device)
    else:
        if not out.is_contiguous() or out.numel() != x2.numel() or out.shape[-1] != d:
            raise RuntimeError("mavlm.layernorm: out must be contiguous [..., D] with the input's element count")
        y, odt = out, out.dtype
    st = _lib.load().mavlm_layernorm_fwd(_ptr(x2), _ptr(gamma), _ptr(beta), _ptr(y), x2.shape[0], d, float(eps),
                                         dtype_code(x2), _DTYPES[odt], _stream())
    _lib.check(st, "layernorm_fwd")
    return y if out is not None else y.reshape(x.shape)


def xattn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, *, head_dim: Optional[int] = None,
          scale: Optional[float] = None, want_lse: bool = False,
          want_col_scores: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """softmax(q k^T * scale) v per head.  q [B, Lq, H*dh], k/v [B, Lk, H*dh] (last dim contiguous; row /
    batch strides free, so k and v may be column slices of one fused projection buffer)."""
    if _grad_needed(q, k, v):
        if want_lse or want_col_scores:
            raise RuntimeError("mavlm.xattn: lse / col_scores outputs are not available under autograd")
        from . import autograd as ag
        dh = head_dim or q.shape[-1] // heads
        sc = scale if scale is not None else 1.0 / math.sqrt(dh)
        return ag.XAttnFn.apply(q, k, v, heads, dh, sc), None, None
    return _xattn_raw(q, k, v, heads, head_dim=head_dim, scale=scale, want_lse=want_lse,
                      want_col_scores=want_col_scores)


def xattn_kv(q: torch.Tensor, kv: torch.Tensor, heads: int, *, head_dim: int, scale: Optional[float] = None) -> torch.Tensor:
    """xattn on a fused projection buffer kv = [B, Lk, 2*H*dh] (k | v column halves).  Same forward as
    xattn(q, kv[..., :hd], kv[..., hd:]); under autograd the gradient comes back as one dkv buffer."""
    hd = heads * head_dim
    if kv.stride(2) != 1 or kv.shape[-1] != 2 * hd:
        raise RuntimeError("mavlm.xattn_kv: kv must be [B, Lk, 2*H*dh] with a contiguous last dim")
    sc = scale if scale is not None else 1.0 / math.sqrt(head_dim)
    if _grad_needed(q, kv):
        from . import autograd as ag
        return ag.XAttnKVFn.apply(q, kv, heads, head_dim, sc)
    return _xattn_raw(q, kv[..., :hd], kv[..., hd:], heads, head_dim=head_dim, scale=sc)[0]


@on_tensor_device
def _xattn_raw(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, *, head_dim: Optional[int] = None,
               scale: Optional[float] = None, want_lse: bool = False,
               want_col_scores: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    _need_cuda(q, k, v)
    if q.dim() != 3 or k.dim() != 3 or v.dim() != 3:
        raise RuntimeError("mavlm.xattn: q, k, v must be [B, L, H*dh]")
    for t in (q, k, v):
        if t.stride(2) != 1:
            raise RuntimeError("mavlm.xattn: last dim must be contiguous")
    b, lq, hd = q.shape
    lk = k.shape[1]
    dh = head_dim or hd // heads
    if scale is None:
        scale = 1.0 / math.sqrt(dh)
    o = torch.empty((b, lq, hd), dtype=q.dtype, device=q.device)
    lse = torch.empty((b, heads, lq), dtype=torch.float32, device=q.device) if want_lse else None
    cs = torch.empty((b, lk), dtype=torch.float32, device=q.device) if want_col_scores else None
    lib = _lib.load()
    code = dtype_code(q)
    ws_bytes = lib.mavlm_xattn_workspace_bytes(b, heads, lq, lk, dh, code)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=q.device) if ws_bytes else None
    st = lib.mavlm_xattn_fwd(_ptr(q), q.stride(1), q.stride(0), _ptr(k), k.stride(1), k.stride(0), _ptr(v),
                             v.stride(1), v.stride(0), _ptr(o), o.stride(1), o.stride(0), _ptr(lse), _ptr(cs), b,
                             heads, lq, lk, dh, float(scale), code, _ptr(ws), ws_bytes, _stream())
    _lib.check(st, "xattn_fwd")
    return o, lse, cs


@on_tensor_device
def xattn_colsum(q: torch.Tensor, k: torch.Tensor, lse: torch.Tensor, heads: int, *, head_dim: Optional[int] = None,
                 scale: Optional[float] = None) -> torch.Tensor:
    """Column sums of the normalised attention probabilities over heads and queries -> fp32 [B, Lk]
    (MemoryController.py:135), tensor-core tier: second pass over K with the LSE of `xattn(..., want_lse=True)`."""
    _need_cuda(q, k, lse)
    if q.dim() != 3 or k.dim() != 3 or q.stride(2) != 1 or k.stride(2) != 1:
        raise RuntimeError("mavlm.xattn_colsum: q, k must be [B, L, H*dh] with a contiguous last dim")
    b, lq, hd = q.shape
    lk = k.shape[1]
    dh = head_dim or hd // heads
    if scale is None:
        scale = 1.0 / math.sqrt(dh)
    if lse.dtype != torch.float32 or tuple(lse.shape) != (b, heads, lq) or not lse.is_contiguous():
        raise RuntimeError("mavlm.xattn_colsum: lse must be contiguous fp32 [B, H, Lq]")
    out = torch.empty((b, lk), dtype=torch.float32, device=q.device)
    st = _lib.load().mavlm_xattn_colsum(_ptr(q), q.stride(1), q.stride(0), _ptr(k), k.stride(1), k.stride(0), _ptr(lse),
                                        _ptr(out), b, heads, lq, lk, dh, float(scale), dtype_code(q), _stream())
    _lib.check(st, "xattn_colsum")
    return out


@on_tensor_device
def pool_pe(x: torch.Tensor, *, side: int, stride: int = 2, mode: str = "bilinear",
            pe_table: Optional[torch.Tensor] = None, frame_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[F, side*side, D] -> [F, out*out, D] (+ pe_table[frame_idx] when given)."""
    modes = {"bilinear": _lib.POOL_BILINEAR, "average": _lib.POOL_AVERAGE, "max": _lib.POOL_MAX}
    if mode not in modes:
        raise ValueError(f"Unexpected mm_spatial_pool_mode: {mode}")          # llava_arch.py:294
    _need_cuda(x, pe_table, frame_idx)
    f, n, d = x.shape
    if n != side * side:
        raise RuntimeError(f"mavlm.pool_pe: {n} tokens is not {side}x{side}")
    out_side = math.ceil(side / stride) if mode == "bilinear" else side // stride
    x = x.contiguous()
    y = torch.empty((f, out_side * out_side, d), dtype=x.dtype, device=x.device)
    if pe_table is not None:
        if pe_table.dtype != torch.float32 or not pe_table.is_contiguous() or pe_table.shape[1] != d:
            raise RuntimeError("mavlm.pool_pe: pe_table must be contiguous fp32 [max_frames, D]")
        frame_idx = frame_idx.to(device=x.device, dtype=torch.int64).contiguous()
    st = _lib.load().mavlm_pool_pe_fwd(_ptr(x), _ptr(y), _ptr(pe_table), _ptr(frame_idx), f, side, out_side, stride, d,
                                       modes[mode], dtype_code(x), _stream())
    _lib.check(st, "pool_pe_fwd")
    return y


def add_rows(x: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    """y[t, n, :] = x[t, n, :] + table[t, :]  (differentiable in x and table)."""
    if _grad_needed(x, table):
        from . import autograd as ag
        return ag.AddRowsFn.apply(x, table)
    return _add_pe_raw(x.contiguous(), table.detach().float().contiguous(), torch.arange(x.shape[0], device=x.device))


def add_pe(x: torch.Tensor, pe_table: torch.Tensor, frame_idx: torch.Tensor,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [T, N, C] + pe_table[frame_idx][:, None, :]  (out may alias x: the kernel is elementwise)."""
    return _add_pe_raw(x, pe_table, frame_idx, out)


@on_tensor_device
def _add_pe_raw(x: torch.Tensor, pe_table: torch.Tensor, frame_idx: torch.Tensor,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, pe_table, frame_idx)
    t, n, c = x.shape
    if not x.is_contiguous():
        if out is x:
            raise RuntimeError("mavlm.add_pe: in-place needs a contiguous tensor")
        x = x.contiguous()
    y = torch.empty_like(x) if out is None else out
    frame_idx = frame_idx.to(device=x.device, dtype=torch.int64).contiguous()
    st = _lib.load().mavlm_add_pe_fwd(_ptr(x), _ptr(y), _ptr(pe_table), _ptr(frame_idx), t, n, c, dtype_code(x),
                                      _stream())
    _lib.check(st, "add_pe_fwd")
    return y


@on_tensor_device
def assemble(seq: torch.Tensor, mem: Optional[torch.Tensor], n_mem_rows: int, frames: torch.Tensor,
             fine_idx: torch.Tensor, tokens: int, type_emb: torch.Tensor, newline: torch.Tensor,
             embed_table: torch.Tensor, prompt_mem_ids: torch.Tensor, prompt_frm_ids: torch.Tensor,
             drop_frames: bool = False) -> torch.Tensor:
    _need_cuda(seq, mem, frames, fine_idx, type_emb, newline, embed_table, prompt_mem_ids, prompt_frm_ids)
    d = seq.shape[-1]
    for t in (mem, frames, type_emb, newline, embed_table):
        if t is not None and (t.dtype != seq.dtype or not t.is_contiguous()):
            raise RuntimeError("mavlm.assemble: operands must be contiguous and of the sequence dtype")
    st = _lib.load().mavlm_assemble_fwd(_ptr(seq), _ptr(mem), n_mem_rows, _ptr(frames), _ptr(fine_idx),
                                        fine_idx.numel(), tokens, _ptr(type_emb), _ptr(newline), _ptr(embed_table),
                                        _ptr(prompt_mem_ids), prompt_mem_ids.numel(), _ptr(prompt_frm_ids),
                                        prompt_frm_ids.numel(), d, int(drop_frames), dtype_code(seq), _stream())
    _lib.check(st, "assemble_fwd")
    return seq


@on_tensor_device
def softmax_rows(scores: torch.Tensor, n: int, scale: float, dtype: torch.dtype, ratio: float = 1.0) -> torch.Tensor:
    """w[r, :n] = ratio * softmax(scale * scores[r, :n]) in `dtype`, columns [n, ld) zero-filled so that w can be the
    A operand of a tensor-core GEMM whose K is padded to a multiple of 8.  scores: fp32 [R, ld] (mavlm_ntm_softmax_fwd)."""
    _need_cuda(scores)
    if scores.dtype != torch.float32 or scores.dim() != 2 or scores.stride(1) != 1:
        raise RuntimeError("mavlm.softmax_rows: scores must be fp32 [rows, ld] with unit inner stride")
    rows, ld = scores.shape
    w = torch.empty((rows, ld), dtype=dtype, device=scores.device)
    st = _lib.load().mavlm_ntm_softmax_fwd(scores.data_ptr(), scores.stride(0), rows, int(n), float(scale), float(ratio),
                                           w.data_ptr(), ld, ld, None, None, 0, _DTYPES[dtype], _stream())
    _lib.check(st, "softmax_rows")
    return w
