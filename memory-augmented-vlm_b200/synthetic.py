"""Random-init weights + synthetic inputs of the shapes BASELINE.json names (there are no
checkpoints or datasets in this environment).  Default PyTorch initialisers, like the reference's
constructors: nn.Linear kaiming-uniform, xavier_uniform_ initial_memory, randn memory_pos_embed
(MemoryController.py:82-84), nn.Embedding N(0,1), image_newline randn * D^-0.5.

`export_weights` returns the numpy dict (reference state_dict key names) that the test-side oracle
consumes; nothing here imports the oracle.
"""
from __future__ import annotations

import types
from typing import Dict, Tuple

import numpy as np
import torch

from .modules import (Config, TemporalPositionalEncoding, TransformerProjector, build_memory_fuser,
                      build_vision_projector)
from .pipeline import VisualMemoryPipeline

OV_DIMS = {"0.5b": 896, "7b": 3584}
SIGLIP_DIM = 1152
SIGLIP_SIDE = 27


def build_pipeline(hidden: int, vision_dim: int = SIGLIP_DIM, *, dtype: torch.dtype = torch.bfloat16, seed: int = 0,
                   chunk_size: int = 32, depth: int = 2, max_frames: int = 600, vocab: int = 50000,
                   num_memory_tokens: int = 8, cache_size: int = 10, device: str = "cuda",
                   q_scale: float = 1.0) -> Tuple[VisualMemoryPipeline, Dict[str, np.ndarray]]:
    cfg = Config()
    cfg.mm_hidden_size = hidden
    cfg.mm_intermediate_size = 4 * hidden
    cfg.depth = depth
    cfg.mm_dtype = torch.float32
    cfg.num_memory_tokens = num_memory_tokens
    cfg.cache_size = cache_size
    torch.manual_seed(seed)
    rmt = TransformerProjector(cfg)
    fuser = build_memory_fuser(hidden)
    proj = build_vision_projector(types.SimpleNamespace(mm_projector_type="mlp2x_gelu", mm_hidden_size=vision_dim,
                                                        hidden_size=hidden))
    pe = TemporalPositionalEncoding(max_frames, hidden, learnable=False)
    tte = torch.nn.Embedding(2, hidden)
    newline = torch.randn(hidden) * hidden ** -0.5
    emb = torch.nn.Embedding(vocab, hidden)
    if q_scale != 1.0:   # stress variant of SURVEY.md §8d: sharp softmax
        with torch.no_grad():
            for n, p in rmt.named_parameters():
                if "q_proj" in n:
                    p.mul_(q_scale)
    weights = export_weights(rmt, fuser, proj, pe, tte, newline, emb)
    pipe = VisualMemoryPipeline(mm_projector=proj, recurrent_memory_transformer=rmt, memory_fuser=fuser,
                                positional_encoding=pe, token_type_embedding=tte, image_newline=newline,
                                embed_tokens=emb, chunk_size=chunk_size)
    pipe = pipe.to(device)
    table = pe.frame_embed
    if dtype != torch.float32:
        pipe = pipe.to(dtype)
        pe.frame_embed = table.float().to(device)   # the PE table stays fp32 (position_encoding.py:36)
    pipe.image_newline = newline.to(device=device, dtype=dtype)
    return pipe, weights


def export_weights(rmt, fuser, proj, pe, tte, newline, emb) -> Dict[str, np.ndarray]:
    w: Dict[str, np.ndarray] = {}
    for pref, mod in (("recurrent_memory_transformer.", rmt), ("memory_fuser.", fuser), ("mm_projector.", proj),
                      ("token_type_embedding.", tte)):
        for k, v in mod.state_dict().items():
            w[pref + k] = v.detach().float().cpu().numpy().astype(np.float64)
    w["image_newline"] = newline.detach().float().cpu().numpy().astype(np.float64)
    w["positional_encoding.frame_embed"] = pe.frame_embed.detach().float().cpu().numpy().astype(np.float64)
    w["embed_tokens.weight"] = emb.weight.detach().float().cpu().numpy().astype(np.float64)
    return w


def round_weights_like(weights: Dict[str, np.ndarray], dtype: torch.dtype) -> Dict[str, np.ndarray]:
    """What the modules actually hold after `.to(dtype)` (the oracle must see the same operands)."""
    if dtype == torch.float32:
        return {k: torch.from_numpy(v).float().double().numpy() for k, v in weights.items()}
    out = {}
    for k, v in weights.items():
        if k == "positional_encoding.frame_embed":
            out[k] = v
        else:
            out[k] = torch.from_numpy(v).to(dtype).double().numpy()
    return out


def synthetic_tower_tokens(videos: int, frames: int, vision_dim: int = SIGLIP_DIM, *, seed: int = 1234,
                           dtype: torch.dtype = torch.bfloat16, pin: bool = False) -> torch.Tensor:
    """N(0,1) SigLIP-shaped tokens [videos, frames, 729, vision_dim] on the host (seed 1234 + video)."""
    out = torch.empty((videos, frames, SIGLIP_SIDE * SIGLIP_SIDE, vision_dim), dtype=dtype,
                      pin_memory=pin)
    for v in range(videos):
        g = torch.Generator().manual_seed(seed + v)
        out[v] = torch.randn(frames, SIGLIP_SIDE * SIGLIP_SIDE, vision_dim, generator=g).to(dtype)
    return out
