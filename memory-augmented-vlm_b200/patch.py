"""Drop the B200 path into a reference LLaVA model (`LlavaQwenForCausalLM`), in place.

    import mavlm_b200
    mavlm_b200.patch_llava(model)        # after from_pretrained(..., torch_dtype=torch.bfloat16).cuda()

What changes: `model.get_model().{mm_projector, recurrent_memory_transformer, memory_fuser,
positional_encoding}` are replaced by the same-named drop-ins holding THE SAME parameter tensors
(state_dict keys unchanged, so checkpoints keep loading/saving), and `model.get_2dPool` is rebound to the
fused kernel.  The tower, the LLM, `token_type_embedding`, `image_newline` and
`prepare_inputs_labels_for_multimodal` itself are untouched: the reference's own loop
(llava_arch.py:481-557) now calls into libmavlm.so through the unchanged call signatures.  The legacy
`MultimodalOpsMixin` methods (`compress_temporal_features`, ...) and an `attention_model`, when the model has
them, are rebound to `legacy.py` the same way.
`model.mavlm_pipeline` additionally exposes the fused whole-path call.
"""
from __future__ import annotations

import types

import torch
from torch import nn

from .modules import (sample_frame_indices, Config, MemoryFuser, MemoryFuserMLP, TemporalPositionalEncoding, TransformerProjector, VisionProjector,
                      get_2dPool)
from .pipeline import VisualMemoryPipeline
from .splice import splice_text_and_vision


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


def _fused_layout_supported(cfg) -> bool:
    """The fused pipeline emits ONE layout: bilinear pooling, video tokens flattened with `image_newline` appended to
    each of {memory, frames} -- what llava_arch.py:571-629 produces for an mm_patch_merge_type that starts with 'spatial'
    AND contains 'unpad' (e.g. 'spatial_unpad', the OneVision setting) with mm_newline_position 'one_token'.  Every other
    combination the reference honours ('flat' = no newline at all, :567-568; 'spatial' without 'unpad' = flattened, no
    newline, :620-629; 'frame' / 'grid' / 'no_token' newline positions, :586-631; average / max pooling, :287-296) has a
    different sequence length or content, so those go to the reference's own method (which then runs on the patched
    modules)."""
    merge = getattr(cfg, "mm_patch_merge_type", "flat")
    newline = getattr(cfg, "mm_newline_position", "one_token")
    pool = getattr(cfg, "mm_spatial_pool_mode", "bilinear")
    return merge.startswith("spatial") and "unpad" in merge and newline == "one_token" and pool == "bilinear"


def _fused_prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels,
                                                images, modalities=["image"], image_sizes=None):
    """Whole-function replacement of LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal (llava_arch.py:388-878)
    for what the fork actually runs -- video samples at inference (Appendix E of SURVEY.md: non-video entries are
    dropped by the fork itself): frame sampling (:437-457) -> the reference's own vision tower -> the fused
    visual-memory pipeline (projector, pool, PE, recurrent memory, fuser, token assembly; one CUDA library) ->
    text / vision splice, truncation, padding (:745-878, one gather kernel).  Same signature, same 6-tuple.
    Training (autograd, dropout_frames coin) and non-video inputs go to the reference's method, which by then runs on
    the patched modules."""
    vision_tower = self.get_vision_tower()
    if vision_tower is None or images is None or input_ids.shape[1] == 1:                     # :392-394
        return input_ids, position_ids, attention_mask, past_key_values, None, labels
    if isinstance(modalities, str):
        modalities = [modalities]
    is_list = type(images) is list
    if (not (is_list or images.ndim == 5) or self.training or any(m != "video" for m in modalities)
            or not _fused_layout_supported(self.config)):
        return self._mavlm_reference_prepare(input_ids, position_ids, attention_mask, past_key_values, labels, images,
                                             modalities, image_sizes)
    pipe = self.mavlm_pipeline
    wdtype = pipe.mm_projector[0].weight.dtype
    dev = pipe.mm_projector[0].weight.device
    feats = []
    for image in images:
        if image.ndim == 3:
            image = image.unsqueeze(0)
        idx = sample_frame_indices(image.shape[0])                                            # :437-451
        tokens = vision_tower(image[idx.to(image.device)])                                    # [F', side^2, Dv]
        seq = pipe(tokens.to(device=dev, dtype=wdtype)[None], idx[None])["sequence"][0]
        feats.append(seq)
    cfg = self.config
    pos, mask, embeds, new_labels = splice_text_and_vision(
        input_ids, position_ids, attention_mask, labels, feats, self.get_model().embed_tokens.weight.detach(),
        tokenizer_model_max_length=getattr(cfg, "tokenizer_model_max_length", None),
        padding_side=getattr(cfg, "tokenizer_padding_side", "right"),
        use_pos_skipping=getattr(cfg, "use_pos_skipping", False),
        pos_skipping_range=getattr(cfg, "pos_skipping_range", 0), training=False)
    return None, pos, mask, past_key_values, embeds, new_labels                             # :878


def patch_llava(model: nn.Module, *, chunk_size: int = 32, fused: bool = False) -> nn.Module:
    inner = model.get_model()
    dtype = next(inner.mm_projector.parameters()).dtype
    if dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise TypeError(f"mavlm: {dtype} is not supported on this path (fp32, bf16 and fp16 only)")
    proj = VisionProjector(*[m for m in inner.mm_projector])
    if isinstance(inner.memory_fuser, nn.Sequential):
        fuser = MemoryFuserMLP(*[m for m in inner.memory_fuser])
    elif hasattr(inner.memory_fuser, "transformer_encoder"):      # the encoder variant (MemoryFuser.py, llava_arch.py:137-143)
        ref = inner.memory_fuser
        layers = ref.transformer_encoder.layers
        fuser = MemoryFuser(ref.input_proj.in_features, num_layers=len(layers), num_heads=layers[0].self_attn.num_heads,
                            dropout=layers[0].dropout.p, device=str(ref.input_proj.weight.device))
        _adopt(fuser, ref)
        fuser.train(ref.training)
    else:
        raise TypeError(f"mavlm: unsupported memory_fuser {type(inner.memory_fuser).__name__}")
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
    # legacy memories: the model class mixes in MultimodalOpsMixin (llava_arch.py:267); rebind its methods
    # to the B200 ones and, when the model carries the Turing-memory module, swap it for the drop-in (same parameters)
    from . import legacy
    if hasattr(model, "compress_temporal_features"):         # the mixin's marker method; `attention` alone is too generic
        for name in ("attention", "attention2", "compress_spatial_features", "compress_temporal_features"):
            if hasattr(model, name):
                setattr(model, name, types.MethodType(getattr(legacy.MultimodalOpsMixin, name), model))
    ntm = getattr(inner, "attention_model", None)
    if isinstance(ntm, nn.Module) and not isinstance(ntm, legacy.NeuralTuringMachine):
        new_ntm = legacy.NeuralTuringMachine(ntm.input_dim, ntm.output_dim, attention_dropout=ntm.dropout.p)
        _adopt(new_ntm, ntm)
        new_ntm.train(ntm.training)
        inner.attention_model = new_ntm
    if fused:   # replace the whole method; the reference's own stays reachable for training / non-video inputs
        ref = getattr(type(model), "prepare_inputs_labels_for_multimodal", None)
        if ref is not None:
            model._mavlm_reference_prepare = types.MethodType(ref, model)
        model.prepare_inputs_labels_for_multimodal = types.MethodType(_fused_prepare_inputs_labels_for_multimodal, model)
    return model
