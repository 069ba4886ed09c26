"""Differentiable CPU oracle for the TRAINING side of the visual-memory path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A plain-PyTorch (CPU, fp32 or fp64, ATen ops only) restatement of the recurrent memory transformer, the
memory fuser and the token assembly, written so that ``torch.autograd`` yields the gradients the reference's
own training step computes (train.py:1694-1728 unfreezes recurrent_memory_transformer, memory_fuser and
token_type_embedding; BPTT runs through the un-detached state cache, MemoryController.py:126,152; frame
features are detached, llava_arch.py:302,481).  It is what the GPU backward tests hold the bf16 / fp32 CUDA
gradients to at sizes where no committed reference-autograd golden fits in the repository (OV-7B dims:
0.47 G parameters).

Only ``tests/`` may import it (as the checker).  The product package never imports anything under ``oracle/``.

Parity pinning: ``tests/test_oracle_golden.py::test_torch_oracle_gradients_*`` hold this file to the gradients
the UNMODIFIED reference modules produced under autograd in float64 (tools/gen_golden.py ->
tests/golden/rmt_grads.npz, path_grads.npz) to 1e-9, and its forward to the numpy oracle.

Weights are a flat dict keyed by the reference ``state_dict`` names (as in vismem_oracle.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MEMORY_PROMPT_IDS = (1986, 374, 264, 1550, 11591, 12126, 315, 279, 2766, 25)      # llava_arch.py:708
FRAME_PROMPT_IDS = (9485, 525, 48876, 9124, 14087, 504, 279, 2766, 25)            # llava_arch.py:714

Weights = Dict[str, torch.Tensor]


def leaf_weights(w_np: Dict[str, np.ndarray], dtype: torch.dtype = torch.float32, *, skip: Sequence[str] = ()) -> Weights:
    """numpy weight dict -> torch leaves with requires_grad (floating tensors only)."""
    out = {}
    for k, v in w_np.items():
        if any(k.startswith(s) for s in skip):
            continue
        t = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
        out[k] = t.requires_grad_(True)
    return out


def _linear(x, w: Weights, p: str):
    return F.linear(x, w[p + "weight"], w[p + "bias"])                      # MemoryController.py:23,37-39,64


def _residual(h, x, w: Weights, p: str, eps: float):
    """Residual.forward: LayerNorm(dense(h) + x)   (MemoryController.py:26-29)."""
    y = _linear(h, w, p + "dense.") + x
    return F.layer_norm(y, (y.shape[-1],), w[p + "layernorm.weight"], w[p + "layernorm.bias"], eps)


def _attention(xq, xkv, w: Weights, p: str, heads: int, eps: float):
    """Attention.forward (MemoryController.py:47-57): q/k/v projections, split heads, softmax(q k^T / sqrt(dh)) v,
    merge heads, Residual.  xq [Lq, D], xkv [Lk, D]."""
    d = xq.shape[-1]
    dh = d // heads
    q = _linear(xq, w, p + "q_proj.").view(-1, heads, dh).transpose(0, 1)
    k = _linear(xkv, w, p + "k_proj.").view(-1, heads, dh).transpose(0, 1)
    v = _linear(xkv, w, p + "v_proj.").view(-1, heads, dh).transpose(0, 1)
    probs = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    ctx = (probs @ v).transpose(0, 1).reshape(-1, d)
    return _residual(ctx, xq, w, p + "residual.", eps)


def _layer(mem, frames, w: Weights, p: str, heads: int, eps: float):
    """TransformerLayer.forward (MemoryController.py:69-72)."""
    a = _attention(mem, frames, w, p + "memory_segment_fusion_attention.", heads, eps)
    up = torch.relu(_linear(a, w, p + "mlp.0."))
    return _residual(up, a, w, p + "residual.", eps)


def rmt_chunk(chunk, cache: List[torch.Tensor], w: Weights, *, prefix: str = "recurrent_memory_transformer.",
              heads: int = 8, depth: int = 2, eps: float = 1e-12, cache_cap: int = 10) -> List[torch.Tensor]:
    """TransformerProjector.forward for one chunk (MemoryController.py:118-158); states are NOT detached."""
    c, p_, d = chunk.shape
    init = w[prefix + "initial_memory"] + w[prefix + "memory_pos_embed"]    # :123
    m_slots = init.shape[0]
    if cache:                                                               # :125-127, :89-97
        mem = cache[-1].reshape(-1, d)
        kv = torch.cat([s.reshape(-1, d) for s in cache], dim=0)
        mem = _attention(mem, kv, w, prefix + "memory_update_attention.", heads, eps)
    else:
        mem = init.to(chunk.dtype).reshape(-1, d)
    frames = chunk.reshape(c * p_, d)
    for l in range(depth):
        mem = _layer(mem, frames, w, f"{prefix}layers.{l}.", heads, eps)
    cache = list(cache) + [mem.reshape(m_slots, p_, d)]                     # :152
    if len(cache) > cache_cap:                                              # :153-154
        cache = cache[-cache_cap:]
    return cache


def chunk_bounds(t: int, d: int) -> List[int]:
    """uniform_segment_variant (segment.py:169-192)."""
    b, cur = [0], 0
    while cur + d <= t:
        cur += d
        b.append(cur)
    if cur < t:
        b.append(t)
    return b


def rmt_video(z, w: Weights, *, chunk: int = 32, **kw) -> List[torch.Tensor]:
    """Chunk scheduler (llava_arch.py:528-537)."""
    bounds = chunk_bounds(z.shape[0], chunk)
    cache: List[torch.Tensor] = []
    for i in range(len(bounds) - 1):
        cache = rmt_chunk(z[bounds[i]:bounds[i + 1]], cache, w, **kw)
    return cache


def fine_frame_indices(n_frames: int, max_fine: int = 32) -> torch.Tensor:
    n = min(max_fine, n_frames)                                             # llava_arch.py:513-522
    return torch.clamp(torch.round(torch.linspace(0, n_frames - 1, steps=n)).long(), 0, n_frames - 1)


def memory_path(z, w: Weights, *, prompt_mem, prompt_frm, chunk: int = 32, max_fine: int = 32, heads: int = 8,
                depth: int = 2, eps: float = 1e-12, cache_cap: int = 10, drop_frames: bool = False
                ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """One video from pooled + PE'd frames z [F, P, D] (detached input) to the assembled sequence:
    RMT over chunks -> fuser MLP (llava_arch.py:132-136, 545-546) -> type embeddings (:548-554) -> flatten +
    image_newline (:620-629) -> cat with the prompt embeddings (:705-731)."""
    z = z.detach()
    cache = rmt_video(z, w, chunk=chunk, heads=heads, depth=depth, eps=eps, cache_cap=cache_cap)
    cat = torch.cat(cache, dim=0)                                           # [n*M, P, D]
    h = F.gelu(_linear(cat, w, "memory_fuser.0."))                          # erf form
    mem = _linear(h, w, "memory_fuser.2.") + w["token_type_embedding.weight"][0]
    fine = z[fine_frame_indices(z.shape[0], max_fine)] + w["token_type_embedding.weight"][1]
    d = z.shape[-1]
    nl = w["image_newline"][None]
    parts = [prompt_mem, mem.reshape(-1, d), nl]
    if not drop_frames:
        parts += [prompt_frm, fine.reshape(-1, d), nl]
    return torch.cat(parts, dim=0), cache


def path_gradients(z_np: np.ndarray, w_np: Dict[str, np.ndarray], *, chunk: int, dtype: torch.dtype = torch.float32,
                   prompt_rows: Optional[Tuple[np.ndarray, np.ndarray]] = None, **kw
                   ) -> Tuple[float, Dict[str, np.ndarray], np.ndarray]:
    """loss = mean(sequence^2) over a BATCH of videos z_np [B, F, P, D] (the loss of the config-4 parity tests,
    SURVEY.md §8d); returns (loss, {state_dict key: gradient}, sequences [B, L, D]).  `prompt_rows`: the embedding
    rows of the two fixed prompts (leaves `embed.prompt_mem` / `embed.prompt_frm`); by default they come from
    w_np['embed_tokens.weight']."""
    w = leaf_weights(w_np, dtype, skip=("embed_tokens.", "positional_encoding.", "mm_projector."))
    if prompt_rows is None:
        tab = w_np["embed_tokens.weight"]
        prompt_rows = (tab[list(MEMORY_PROMPT_IDS)], tab[list(FRAME_PROMPT_IDS)])
    pm = torch.from_numpy(np.ascontiguousarray(prompt_rows[0])).to(dtype).requires_grad_(True)
    pf = torch.from_numpy(np.ascontiguousarray(prompt_rows[1])).to(dtype).requires_grad_(True)
    z = torch.from_numpy(np.ascontiguousarray(z_np)).to(dtype)
    seqs = []
    total = None
    n_el = 0
    for b in range(z.shape[0]):
        seq, _ = memory_path(z[b], w, prompt_mem=pm, prompt_frm=pf, chunk=chunk, **kw)
        seqs.append(seq.detach())
        s = (seq * seq).sum()
        total = s if total is None else total + s
        n_el += seq.numel()
    loss = total / n_el
    loss.backward()
    def _np(t):                                                             # bf16 (noise-floor runs) has no numpy dtype
        return t.detach().numpy() if t.dtype in (torch.float32, torch.float64) else t.detach().double().numpy()

    grads = {k: _np(v.grad) for k, v in w.items() if v.grad is not None}
    grads["embed.prompt_mem"] = _np(pm.grad)
    grads["embed.prompt_frm"] = _np(pf.grad)
    return float(loss.detach()), grads, _np(torch.stack(seqs))
