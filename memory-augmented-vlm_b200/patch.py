"""Drop the B200 path into a reference LLaVA model (`LlavaQwenForCausalLM`), in place.

    import mavlm_b200
    mavlm_b200.patch_llava(model)        # after from_pretrained(..., torch_dtype=torch.bfloat16).cuda()

What changes: `model.get_model().{mm_projector, recurrent_memory_transformer, memory_fuser,
positional_encoding}` are replaced by the same-named drop-ins holding THE SAME parameter tensors
(state_dict keys unchanged, so checkpoints keep loading/saving), and `model.get_2dPool` is rebound to the
fused kernel.  The tower, the LLM, `token_type_embedding`, `image_newline` and
`prepare_inputs_labels_for_multimodal` itself are untouched: the reference's own loop
(llava_arch.py:481-557) now calls into libmavlm.so through the unchanged call signatures.
`model.mavlm_pipeline` additionally exposes the fused whole-path call.
"""
from __future__ import annotations

import types

import torch
from torch import nn

from .modules import (Config, MemoryFuserMLP, TemporalPositionalEncoding, TransformerProjector, VisionProjector,
                      get_2dPool)
from .pipeline import VisualMemoryPipeline


def _adopt(dst: nn.Module, src: nn.Module) -> None:
    """Make `dst` use src's parameter / buffer tensors (no copy; names must match)."""
    sp = dict(src.named_parameters())
    sb = dict(src.named_buffers())
    for name, _ in list(dst.named_parameters()):
        mod, leaf = _owner(dst, name)
        mod._parameters[leaf] = sp[name]
    for name, _ in list(dst.named_buffers()):
        mod, leaf = _owner(dst, name)
        mod._buffers[leaf] = sb[name]


def _owner(root: nn.Module, dotted: str):
    parts = dotted.split(".")
    m = root
    for p in parts[:-1]:
        m = getattr(m, p)
    return m, parts[-1]


def convert_rmt(ref_rmt: nn.Module) -> TransformerProjector:
    c = ref_rmt.config
    cfg = Config()
    for k in ("mm_hidden_size", "mm_hidden_act", "mm_num_attention_heads", "patch_size", "mm_layer_norm_eps",
              "mm_intermediate_size", "num_memory_tokens", "depth"):
        setattr(cfg, k, getattr(c, k))
    cfg.mm_dtype = next(ref_rmt.parameters()).dtype
    new = TransformerProjector(cfg)
    _adopt(new, ref_rmt)
    return new


def patch_llava(model: nn.Module, *, chunk_size: int = 32) -> nn.Module:
    inner = model.get_model()
    dtype = next(inner.mm_projector.parameters()).dtype
    if dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise TypeError(f"mavlm: {dtype} is not supported on this path (fp32, bf16 and fp16 only)")
    proj = VisionProjector(*[m for m in inner.mm_projector])
    fuser = MemoryFuserMLP(*[m for m in inner.memory_fuser])
    rmt = convert_rmt(inner.recurrent_memory_transformer)
    old_pe = inner.positional_encoding
    pe = TemporalPositionalEncoding(old_pe.max_frames, old_pe.embed_dim, learnable=old_pe.learnable)
    _adopt(pe, old_pe)
    inner.mm_projector = proj
    inner.memory_fuser = fuser
    inner.recurrent_memory_transformer = rmt
    inner.positional_encoding = pe
    side = model.get_vision_tower().num_patches_per_side

    def _get_2dpool(self, image_feature, stride=2):                      # llava_arch.py:277
        return get_2dPool(image_feature, stride, mode=self.config.mm_spatial_pool_mode, num_patches_per_side=side)

    model.get_2dPool = types.MethodType(_get_2dpool, model)
    # not registered as a sub-module: the state_dict must stay exactly the reference's
    model.__dict__["mavlm_pipeline"] = VisualMemoryPipeline(
        mm_projector=proj, recurrent_memory_transformer=rmt, memory_fuser=fuser, positional_encoding=pe,
        token_type_embedding=inner.token_type_embedding, image_newline=inner.image_newline,
        embed_tokens=inner.embed_tokens, chunk_size=chunk_size, num_patches_per_side=side)
    return model
