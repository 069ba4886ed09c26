"""Host-side mirror of the reference's memory modules: same class names, constructor arguments,
attributes, call signatures and state_dict keys as

    llava/model/memory_module/MemoryController.py   (Config, Residual, Attention, TransformerLayer,
                                                      TransformerProjector)
    llava/model/memory_module/position_encoding.py  (TemporalPositionalEncoding)
    llava/model/multimodal_projector/builder.py     (build_vision_projector: mlp2x_gelu)
    llava/model/llava_arch.py:132-136, 277-297      (memory_fuser MLP, get_2dPool)

so that a reference checkpoint loads unchanged and `patch_llava` can swap them in.  The modules own
parameters (PyTorch storage) and nothing else: every forward runs through the C ABI in
libmavlm.so (ops.py), and under autograd every backward does too (autograd.py); hyper-parameters that the
reference hard-codes (chunk 32, cache depth 10, 8 slots x 196 tokens, 8 heads, depth 2, PE table 600)
are constructor arguments whose defaults equal the reference.
"""
from __future__ import annotations

import math
import re
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import ops
from ._device import on_tensor_device
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_RELU

_ACTS = {"relu": ACT_RELU, "gelu": ACT_GELU_ERF, "none": ACT_NONE}


class Config:
    """Same fields / defaults as MemoryController.py:7-18 (overridden at llava_arch.py:118-129)."""
    mm_hidden_size = 896
    mm_hidden_act = "relu"
    mm_num_attention_heads = 8
    patch_size = 196
    mm_attention_probs_dropout_prob = 0.1   # never applied by the reference (no nn.Dropout in the RMT)
    mm_layer_norm_eps = 1e-12
    mm_hidden_dropout_prob = 0.1
    mm_intermediate_size = 4 * mm_hidden_size
    num_memory_tokens = 8
    depth = 1
    mm_dtype = torch.float16
    # additions (reference hard-codes these): MemoryController.py:153-154
    cache_size = 10
    frame_scores = False


def _param_dtype(config) -> torch.dtype:
    # the reference creates fp16 parameters (MemoryController.py:17) and relies on from_pretrained(torch_dtype=...)
    # to cast; fp16 runs the tensor-core tier like bf16 (inference only), fp32 the exact SIMT tier.
    return getattr(config, "mm_dtype", torch.float32)


class Residual(nn.Module):
    """LayerNorm(dense(h) + x)   (MemoryController.py:20-29)."""

    def __init__(self, input_size, output_size, config):
        super().__init__()
        dt = _param_dtype(config)
        self.dense = nn.Linear(input_size, output_size, dtype=dt)
        self.layernorm = nn.LayerNorm(output_size, eps=config.mm_layer_norm_eps, dtype=dt)

    def forward(self, hidden_states: torch.Tensor, input_tensor: torch.Tensor, *, weight=None, out=None,
                fillers=()) -> torch.Tensor:
        w = self.dense.weight if weight is None else weight
        if fillers:     # tail fill (ops.linear_fill): off-critical-path GEMM tiles ride in this launch's idle tile slots
            pre = ops.linear_fill(hidden_states, w, self.dense.bias, resid=input_tensor, out_dtype=torch.float32,
                                  fillers=fillers)
        else:
            pre = ops.linear(hidden_states, w, self.dense.bias, resid=input_tensor, out_dtype=torch.float32)
        return ops.layernorm(pre, self.layernorm.weight, self.layernorm.bias, self.layernorm.eps,
                             out_dtype=input_tensor.dtype, out=out)


def _padded_head_dim(dh: int, dtype: torch.dtype) -> int:
    """The tcgen05 attention kernel handles head_dim 128 and 448; smaller heads (0.5B: 112) are
    zero-padded to 128 by packing the projection weights.  fp32 needs no padding."""
    if dtype == torch.float32 or dh in (128, 448):
        return dh
    if dh < 128:
        return 128
    raise RuntimeError(f"mavlm: head_dim {dh} is not supported by the bf16 tensor-core tier (<=128 or 448)")


class Attention(nn.Module):
    """8-head cross attention + Residual   (MemoryController.py:31-57).

    forward returns (output, probs) like the reference; probs is None here because the fused kernel
    never materialises them (the reference's [1,8,Lq,Lk] tensor is 157 MB per layer at 7B).  Column
    sums of the probabilities (all the reference ever derives from them, :135) are available through
    `col_scores=True` (fp32 tier: fused into the attention; bf16 / fp16 tier: a second tensor-core pass with the LSE)."""

    def __init__(self, config):
        super().__init__()
        dt = _param_dtype(config)
        self.hidden_size = config.mm_hidden_size
        self.num_attention_heads = config.mm_num_attention_heads
        self.attention_head_size = self.hidden_size // self.num_attention_heads
        self.k_proj = nn.Linear(self.hidden_size, self.hidden_size, dtype=dt)
        self.v_proj = nn.Linear(self.hidden_size, self.hidden_size, dtype=dt)
        self.q_proj = nn.Linear(self.hidden_size, self.hidden_size, dtype=dt)
        self.residual = Residual(self.hidden_size, self.hidden_size, config)
        self._pack = None
        self._pack_key = None
        self.last_col_scores: Optional[torch.Tensor] = None

    # ---- weight packing: [Wk;Wv] fused, heads padded for the tensor-core tier --------------------
    def _params(self):
        return (self.q_proj.weight, self.q_proj.bias, self.k_proj.weight, self.k_proj.bias, self.v_proj.weight,
                self.v_proj.bias, self.residual.dense.weight)

    def _build_pack(self):
        """[Wk;Wv] fused and heads padded.  Built from differentiable torch ops, so the same code serves the
        cached inference pack (under no_grad) and the per-forward training pack (gradients flow to q/k/v_proj)."""
        h, dh, d = self.num_attention_heads, self.attention_head_size, self.hidden_size
        dhp = _padded_head_dim(dh, self.q_proj.weight.dtype)

        def pad_rows(w, b):
            if dhp == dh:
                return w, b
            z = w.new_zeros(h, dhp - dh, d)
            zb = b.new_zeros(h, dhp - dh)
            return (torch.cat([w.view(h, dh, d), z], dim=1).reshape(h * dhp, d),
                    torch.cat([b.view(h, dh), zb], dim=1).reshape(h * dhp))

        wq, bq = pad_rows(self.q_proj.weight, self.q_proj.bias)
        wk, bk = pad_rows(self.k_proj.weight, self.k_proj.bias)
        wv, bv = pad_rows(self.v_proj.weight, self.v_proj.bias)
        wkv = torch.cat([wk, wv], dim=0)
        bkv = torch.cat([bk, bv], dim=0)
        wo = self.residual.dense.weight
        if dhp != dh:
            wo = torch.cat([wo.view(d, h, dh), wo.new_zeros(d, h, dhp - dh)], dim=2).reshape(d, h * dhp)
        return {"wq": wq.contiguous(), "bq": bq.contiguous(), "wkv": wkv.contiguous(), "bkv": bkv.contiguous(),
                "wo": wo.contiguous(), "dhp": dhp}

    def packed(self):
        ps = self._params()
        if torch.is_grad_enabled() and any(p.requires_grad for p in ps):
            return self._build_pack()                                   # training: part of the autograd graph
        key = tuple((p.data_ptr(), p._version, p.dtype, p.device) for p in ps)
        if self._pack_key == key:
            return self._pack
        with torch.no_grad():
            self._pack = self._build_pack()
        self._pack_key = key
        return self._pack

    def project_kv(self, kv_hidden_states: torch.Tensor) -> torch.Tensor:
        """[.., Lk, D] -> fused [.., Lk, 2*H*dhp] buffer (k | v): one GEMM, reusable across calls."""
        p = self.packed()
        return ops.linear(kv_hidden_states, p["wkv"], p["bkv"])

    def forward(self, hidden_states, kv_hidden_states=None, output_attentions=True, *, kv_projected=None,
                col_scores=False):
        p = self.packed()
        h, dhp = self.num_attention_heads, p["dhp"]
        x = hidden_states if hidden_states.dim() == 3 else hidden_states.unsqueeze(0)
        if kv_projected is None:
            src = x if kv_hidden_states is None else kv_hidden_states
            src = src if src.dim() == 3 else src.unsqueeze(0)
            kv_projected = self.project_kv(src)
        hd = h * dhp
        scale = 1.0 / math.sqrt(self.attention_head_size)
        q = ops.linear(x, p["wq"], p["bq"])
        fp32 = q.dtype == torch.float32
        k_, v_ = kv_projected[..., :hd], kv_projected[..., hd:]
        grad = torch.is_grad_enabled() and (q.requires_grad or kv_projected.requires_grad)
        fused_scores = col_scores and fp32 and not grad
        tc_scores = col_scores and not fp32 and not grad
        ctx, lse, cs = ops.xattn(q, k_, v_, h, head_dim=dhp, scale=scale, want_col_scores=fused_scores, want_lse=tc_scores)
        if tc_scores:
            # frame scores (MemoryController.py:135, detached at :157) need normalised probabilities, which the fused
            # kernel never forms: a second tensor-core pass over K with the saved LSE (half the attention's MMA work)
            cs = ops.xattn_colsum(q, k_, lse, h, head_dim=dhp, scale=scale)
        elif col_scores and not fused_scores:
            # under autograd (diagnostics only, detached): the same second pass on detached operands
            with torch.no_grad():
                qd, kd, vd = q.detach(), k_.detach(), v_.detach()
                if fp32:
                    _, _, cs = ops.xattn(qd, kd, vd, h, head_dim=dhp, scale=scale, want_col_scores=True)
                else:
                    _, lse_d, _ = ops.xattn(qd, kd, vd, h, head_dim=dhp, scale=scale, want_lse=True)
                    cs = ops.xattn_colsum(qd, kd, lse_d, h, head_dim=dhp, scale=scale)
        self.last_col_scores = cs
        out = self.residual(ctx, x, weight=p["wo"])
        return out.reshape(hidden_states.shape), None


class _ReLU(nn.Module):  # placeholder so that the MLP keeps the reference's `mlp.0.*` state_dict keys
    def forward(self, x):  # pragma: no cover - never called, the activation is fused into the GEMM epilogue
        raise RuntimeError("fused into the GEMM epilogue")


class TransformerLayer(nn.Module):
    """Cross-attention block + ReLU MLP with residual LayerNorm   (MemoryController.py:59-72)."""

    def __init__(self, config):
        super().__init__()
        dt = _param_dtype(config)
        self.memory_segment_fusion_attention = Attention(config)
        self.mlp = nn.Sequential(nn.Linear(config.mm_hidden_size, config.mm_intermediate_size, dtype=dt), _ReLU())
        self._act = _ACTS[config.mm_hidden_act]
        self.residual = Residual(config.mm_intermediate_size, config.mm_hidden_size, config)

    def forward(self, query_states, kv_states, *, kv_projected=None, col_scores=False):
        a, probs = self.memory_segment_fusion_attention(query_states, kv_hidden_states=kv_states,
                                                        output_attentions=True, kv_projected=kv_projected,
                                                        col_scores=col_scores)
        up = ops.linear(a, self.mlp[0].weight, self.mlp[0].bias, act=self._act)
        return self.residual(up, a), probs


class TransformerProjector(nn.Module):
    """Recurrent memory transformer   (MemoryController.py:74-158).

    forward(image_features [C, P, D]) -> (memory_cache, frame_attn_scores): the live cache list (<= cache_size
    states of [M, P, D]) and the list of per-chunk frame scores ([C] each, filled when config.frame_scores;
    in bf16 they cost an extra fp32-tier pass over the last layer's scores).  State lives on the module like the reference
    (`rmt.memory_cache = []` resets it, llava_arch.py:532); frame_attn_scores is cleared with it."""

    def __init__(self, config=None):
        super().__init__()
        self.config = config or Config()
        self.layers = nn.ModuleList([TransformerLayer(self.config) for _ in range(self.config.depth)])
        self.num_memory_tokens = self.config.num_memory_tokens
        self.hidden_size = self.config.mm_hidden_size
        self.patch_size = self.config.patch_size
        self.initial_memory = nn.Parameter(torch.empty(self.num_memory_tokens, self.patch_size, self.hidden_size))
        self.memory_pos_embed = nn.Parameter(torch.randn(self.num_memory_tokens, 1, self.hidden_size))
        nn.init.xavier_uniform_(self.initial_memory)
        self._memory_cache: List[torch.Tensor] = []
        self.memory_update_attention = Attention(self.config)
        self.frame_attn_scores: List[torch.Tensor] = []
        self.cache_size = getattr(self.config, "cache_size", 10)
        self._kv_cache: list = []      # (state, version, projected k|v) per cached state, evolution attention

    # the reference resets state by assigning `memory_cache = []`; keep derived caches in sync
    @property
    def memory_cache(self) -> List[torch.Tensor]:
        return self._memory_cache

    @memory_cache.setter
    def memory_cache(self, value):
        self._memory_cache = value
        self._kv_cache = []
        if not value:
            self.frame_attn_scores = []

    def initial_state(self, dtype: torch.dtype) -> torch.Tensor:
        """(initial_memory + memory_pos_embed).to(dtype)   (MemoryController.py:123-124) via the PE-add kernel."""
        im = self.initial_memory
        pos = self.memory_pos_embed
        table = pos.reshape(self.num_memory_tokens, self.hidden_size)
        if torch.is_grad_enabled() and (im.requires_grad or pos.requires_grad):
            return ops.add_rows(im if im.dtype == dtype else im.to(dtype), table)   # training: part of the graph
        # inference: the sum only changes when a parameter does -- cached (it was 4 small launches per video)
        key = (im.data_ptr(), im._version, pos.data_ptr(), pos._version, dtype, im.device)
        if getattr(self, "_init_state_key", None) != key:
            with torch.no_grad():
                self._init_state = ops.add_rows(im if im.dtype == dtype else im.to(dtype), table)
            self._init_state_key = key
        return self._init_state

    def _update_memory_tokens_with_cache(self, current_memory: torch.Tensor) -> torch.Tensor:
        """Memory evolution (MemoryController.py:89-115): Q = last state, K/V = all cached states.
        The K/V projection of older states is cached (the reference re-projects the whole cache)."""
        if not self._memory_cache:
            return current_memory
        att = self.memory_update_attention
        m, p, d = current_memory.shape
        # entries are (state, its version, projection) matched by IDENTITY against the live list, so a cache the caller
        # mutated in place (pop / clear / slice assignment / in-place edits of a state) never meets a stale projection
        old = {id(e[0]): e for e in self._kv_cache}
        entries = []
        for s in self._memory_cache:
            e = old.get(id(s))
            if e is None or e[0] is not s or e[1] != s._version:
                e = (s, s._version, att.project_kv(s.reshape(1, m * p, d)))
            entries.append(e)
        self._kv_cache = entries
        kv = entries[0][2] if len(entries) == 1 else torch.cat([e[2] for e in entries], dim=1)
        out, _ = att(current_memory.reshape(1, m * p, d), kv_projected=kv)
        return out.reshape(m, p, d)

    def forward(self, image_features: torch.Tensor):
        f, p, d = image_features.shape
        dtype = image_features.dtype
        if self._memory_cache:
            memory_tokens = self._update_memory_tokens_with_cache(self._memory_cache[-1])
        else:
            memory_tokens = self.initial_state(dtype)
        memory_2d = memory_tokens.reshape(1, self.num_memory_tokens * p, d)
        image_2d = image_features.reshape(1, f * p, d)
        want_scores = bool(getattr(self.config, "frame_scores", False))
        last = len(self.layers) - 1
        for i, layer in enumerate(self.layers):
            memory_2d, _ = layer(memory_2d, image_2d, col_scores=want_scores and i == last)
        final_memory = memory_2d.view(self.num_memory_tokens, p, d)
        self._memory_cache.append(final_memory)
        if len(self._memory_cache) > self.cache_size:
            self._memory_cache = self._memory_cache[-self.cache_size:]
            live = {id(t) for t in self._memory_cache}
            self._kv_cache = [e for e in self._kv_cache if id(e[0]) in live]
        if want_scores:
            cs = self.layers[last].memory_segment_fusion_attention.last_col_scores
            self.frame_attn_scores.append(cs.view(f, p).mean(dim=1).detach())   # MemoryController.py:135-137
        return self._memory_cache, self.frame_attn_scores


class TemporalPositionalEncoding(nn.Module):
    """Fixed sinusoidal (or learnable) temporal PE   (position_encoding.py:13-80)."""

    def __init__(self, max_frames, embed_dim, learnable=True):
        super().__init__()
        self.max_frames = max_frames
        self.embed_dim = embed_dim
        self.learnable = learnable
        if learnable:
            self.frame_embed = nn.Embedding(max_frames, embed_dim)
        else:
            pe = torch.zeros(max_frames, embed_dim, dtype=torch.float32)
            position = torch.arange(0, max_frames).unsqueeze(1).float()
            div_term = torch.exp(torch.arange(0, embed_dim, 2).float() * -(math.log(10000.0) / embed_dim))
            pe[:, 0::2] = torch.sin(position * div_term)
            pe[:, 1::2] = torch.cos(position * div_term)
            self.register_buffer("frame_embed", pe)

    def table(self) -> torch.Tensor:
        t = self.frame_embed.weight if self.learnable else self.frame_embed
        return t.detach().float().contiguous()

    def validate(self, frame_indices) -> None:
        """Index checks of position_encoding.py:73-76, done on the host copy of the indices so that
        the kernel path stays free of device->host syncs."""
        idx = frame_indices.detach().cpu() if isinstance(frame_indices, torch.Tensor) else torch.as_tensor(frame_indices)
        if idx.numel() and int(idx.max()) >= self.max_frames:
            raise ValueError(f"indices exceed max_frames: max {int(idx.max())} vs limit {self.max_frames}")
        if idx.numel() and int(idx.min()) < 0:
            raise ValueError(f"indices contains negative values: min {int(idx.min())}")

    def forward(self, x, frame_indices=None):
        if x.dim() not in (3, 4):
            raise ValueError(f"Expected 3D or 4D input, got {x.dim()}D.")
        if x.dim() == 4:
            b, t = x.shape[:2]
            if frame_indices is None:
                frame_indices = torch.arange(t).expand(b, t)
            self.validate(frame_indices)
            y = ops.add_pe(x.reshape(b * t, x.shape[2], x.shape[3]), self.table(),
                           frame_indices.reshape(-1).to(x.device))
            return y.reshape(x.shape)
        if frame_indices is None:
            frame_indices = torch.arange(x.size(0))
        self.validate(frame_indices)
        return ops.add_pe(x, self.table(), frame_indices.to(x.device))


class _FusedMLP(nn.Sequential):
    """Linear -> GELU(erf) -> Linear with the reference's Sequential key names (`0.*`, `2.*`)."""

    def forward(self, x, *, out=None, addvec=None):
        h = ops.linear(x, self[0].weight, self[0].bias, act=ACT_GELU_ERF)
        return ops.linear(h, self[2].weight, self[2].bias, out=out, addvec=addvec)


class VisionProjector(_FusedMLP):
    """mlp2x_gelu mm_projector   (multimodal_projector/builder.py:41-48)."""


class MemoryFuserMLP(_FusedMLP):
    """The live memory_fuser   (llava_arch.py:132-136)."""


class MemoryFuser(nn.Module):
    """The reference's encoder-style fuser (MemoryFuser.py:4-30; imported by llava_arch.py:40, use commented
    out at :137-143): Linear -> nn.TransformerEncoder (post-LN, GELU(erf) FFN 4D, `num_heads` heads, LN eps
    1e-5) -> Linear, self-attention over the tokens of each batch row.  Same constructor, same state_dict keys
    (the nn.TransformerEncoder modules are kept as parameter containers).  Inference mode only: train-mode
    dropout (0.1) is not implemented on this path."""

    def __init__(self, hidden_dim, num_layers=2, num_heads=4, dropout=0.1, device="cuda"):
        super().__init__()
        self.device = device
        self.num_heads = num_heads
        self.dropout = dropout
        self.input_proj = nn.Linear(hidden_dim, hidden_dim)
        encoder_layer = nn.TransformerEncoderLayer(d_model=hidden_dim, nhead=num_heads, dim_feedforward=hidden_dim * 4,
                                                   dropout=dropout, batch_first=True, activation="gelu")
        self.transformer_encoder = nn.TransformerEncoder(encoder_layer, num_layers=num_layers,
                                                         enable_nested_tensor=False)
        self.output_proj = nn.Linear(hidden_dim, hidden_dim)

    def forward(self, memory_tokens: torch.Tensor) -> torch.Tensor:
        if self.training and self.dropout > 0:
            raise NotImplementedError("MemoryFuser: train-mode dropout is not implemented on the B200 path; "
                                      "call .eval()")
        b, s_len, d = memory_tokens.shape
        h = self.num_heads
        dh = d // h
        dt = memory_tokens.dtype
        tc = dt in (torch.bfloat16, torch.float16) and dh in (128, 448)   # head dims the fused tensor-core kernel handles
        # other head dims in bf16 / fp16 (7B: dh = 896, 0.5B: 224): attention as batched tcgen05 GEMMs around a row softmax,
        # on sequences padded to a multiple of 8 tokens (the GEMM's K alignment); the pad keys get zero probability
        gemm_attn = dt in (torch.bfloat16, torch.float16) and not tc and dh % 8 == 0 and d % 8 == 0
        s_pad = (s_len + 7) // 8 * 8 if gemm_attn else s_len
        if s_pad != s_len:
            padded = memory_tokens.new_zeros((b, s_pad, d))
            padded[:, :s_len].copy_(memory_tokens)
            memory_tokens = padded
        x = ops.linear(memory_tokens, self.input_proj.weight, self.input_proj.bias)
        for layer in self.transformer_encoder.layers:
            sa = layer.self_attn
            if gemm_attn:
                qkv = ops.linear(x, sa.in_proj_weight, sa.in_proj_bias)
                ctx = self._gemm_attention(qkv, b, s_pad, s_len, h, dh)
            elif tc or dt == torch.float32:
                qkv = ops.linear(x, sa.in_proj_weight, sa.in_proj_bias)
                ctx, _, _ = ops.xattn(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], h)
            else:                                                # e.g. 7B: dh = 896 -> fp32-tier attention on fp32 q/k/v
                qkv = ops.linear(x, sa.in_proj_weight, sa.in_proj_bias, out_dtype=torch.float32)
                ctx32, _, _ = ops.xattn(qkv[..., :d], qkv[..., d:2 * d], qkv[..., 2 * d:], h)
                ctx = ops.cast(ctx32, dt)
            pre = ops.linear(ctx, sa.out_proj.weight, sa.out_proj.bias, resid=x, out_dtype=torch.float32)
            x = ops.layernorm(pre, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps, out_dtype=dt)
            f = ops.linear(x, layer.linear1.weight, layer.linear1.bias, act=ACT_GELU_ERF)
            pre = ops.linear(f, layer.linear2.weight, layer.linear2.bias, resid=x, out_dtype=torch.float32)
            x = ops.layernorm(pre, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps, out_dtype=dt)
        y = ops.linear(x, self.output_proj.weight, self.output_proj.bias)
        return y[:, :s_len] if s_pad != s_len else y

    @staticmethod
    @on_tensor_device
    def _gemm_attention(qkv: torch.Tensor, b: int, s_pad: int, s_len: int, h: int, dh: int) -> torch.Tensor:
        """softmax(q k^T / sqrt(dh)) v per (row, head) of qkv [b, s_pad, 3 h dh] (bf16): two batched GEMMs (fp32 scores)
        and one row-softmax kernel that masks the pad keys."""
        from . import _lib
        from .autograd import _DT, _p, _s
        d = h * dh
        lib = _lib.load()
        dev = qkv.device
        scores = torch.empty((b * h * s_pad, s_pad), dtype=torch.float32, device=dev)
        ctx = torch.empty((b, s_pad, d), dtype=qkv.dtype, device=dev)
        import ctypes
        ld = 3 * d

        def strides(*v):
            return (ctypes.c_int64 * 6)(*v)

        q, k, v = qkv, qkv[..., d:], qkv[..., 2 * d:]
        # scores[b, h] = q[b, :, h] k[b, :, h]^T      (outer = rows of the batch, inner = heads)
        _lib.check(lib.mavlm_gemm_ex(_p(q), ld, 0, k.data_ptr(), ld, 1, _p(scores), s_pad, s_pad, s_pad, dh, 1.0, 0, b, h,
                                     strides(s_pad * ld, dh, s_pad * ld, dh, h * s_pad * s_pad, s_pad * s_pad),
                                     _DT[qkv.dtype], _DT[torch.float32], _s()), "gemm_ex")
        probs = ops.softmax_rows(scores, s_len, 1.0 / math.sqrt(dh), qkv.dtype)
        # ctx[b, :, h] = probs[b, h] v[b, :, h]
        _lib.check(lib.mavlm_gemm_ex(_p(probs), s_pad, 0, v.data_ptr(), ld, 0, _p(ctx), d, s_pad, dh, s_pad, 1.0, 0, b, h,
                                     strides(h * s_pad * s_pad, s_pad * s_pad, s_pad * ld, dh, s_pad * d, dh),
                                     _DT[qkv.dtype], _DT[qkv.dtype], _s()), "gemm_ex")
        return ctx


def build_vision_projector(config, delay_load=False, **kwargs):
    """Same entry point as multimodal_projector/builder.py:32-65; only the projector type the
    OneVision checkpoints use (mlp2x_gelu) is on this path."""
    projector_type = getattr(config, "mm_projector_type", "linear")
    m = re.match(r"^mlp(\d+)x_gelu$", projector_type)
    if m and int(m.group(1)) == 2:
        return VisionProjector(nn.Linear(config.mm_hidden_size, config.hidden_size), nn.GELU(),
                               nn.Linear(config.hidden_size, config.hidden_size))
    raise ValueError(f"Unknown projector type: {projector_type} (this path implements mlp2x_gelu)")


def build_memory_fuser(hidden_dim: int) -> MemoryFuserMLP:
    return MemoryFuserMLP(nn.Linear(hidden_dim, hidden_dim * 4), nn.GELU(), nn.Linear(hidden_dim * 4, hidden_dim))


def get_2dPool(image_feature: torch.Tensor, stride: int = 2, *, mode: str = "bilinear",
               num_patches_per_side: Optional[int] = None) -> torch.Tensor:
    """[F, side*side, D] -> [F, ceil(side/stride)^2, D]   (LlavaMetaForCausalLM.get_2dPool, llava_arch.py:277-297)."""
    side = num_patches_per_side or int(round(math.sqrt(image_feature.shape[1])))
    return ops.pool_pe(image_feature, side=side, stride=stride, mode=mode)


def uniform_segment_variant(features, d=32):
    """Chunk boundaries [0, d, 2d, ..., T]   (segment.py:169-192); only shape[0] of `features` is read."""
    t = features if isinstance(features, int) else features.shape[0]
    bounds = [0]
    cur = 0
    while cur + d <= t:
        cur += d
        bounds.append(cur)
    if cur < t:
        bounds.append(t)
    return bounds


def sample_frame_indices(num_frames: int) -> torch.Tensor:
    """Frame sampling of llava_arch.py:437-451 (host index math, torch.linspace like the reference)."""
    if num_frames < 32:
        n = num_frames
    else:
        n = max((num_frames // 32) * 32, 64)
    return torch.linspace(0, num_frames - 1, steps=n).long()


def fine_frame_indices(num_sampled: int, max_fine: int = 32) -> torch.Tensor:
    """Fine-grained frame pick of llava_arch.py:513-522."""
    n = min(max_fine, num_sampled)
    idx = torch.round(torch.linspace(0, num_sampled - 1, steps=n)).long()
    return torch.clamp(idx, 0, num_sampled - 1)
