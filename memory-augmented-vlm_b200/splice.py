"""Text / vision splice, truncation and padding: the tail of prepare_inputs_labels_for_multimodal
(llava_arch.py:745-878), i.e. the immediate consumer of the assembled video-token sequence.

Host side: the same index logic as the reference (strip padding with the attention mask, split the ids at
IMAGE_TOKEN_INDEX, interleave text chunks and image features, IGNORE_INDEX labels for the vision rows, truncate to
tokenizer_model_max_length, pad left or right, attention mask, position ids) expressed as ONE row-source table per
batch; device side: a single gather kernel (mavlm_gather_rows_fwd) writes the padded [B, Lmax, D] embedding tensor
(the reference does ~6 full-sequence torch.cat copies per sample).  Inference / frozen-embedding path: the gather is
not differentiable (the differentiable assembly is VisualMemoryPipeline.memory_forward_train).
"""
from __future__ import annotations

import random
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from ._device import on_tensor_device

IGNORE_INDEX = -100          # llava/constants.py:7
IMAGE_TOKEN_INDEX = -200     # llava/constants.py:8


@on_tensor_device
@torch.no_grad()
def splice_text_and_vision(input_ids: torch.Tensor, position_ids: Optional[torch.Tensor],
                           attention_mask: Optional[torch.Tensor], labels: Optional[torch.Tensor],
                           image_features: Sequence[torch.Tensor], embed_table: torch.Tensor, *,
                           tokenizer_model_max_length: Optional[int] = None, padding_side: str = "right",
                           use_pos_skipping: bool = False, pos_skipping_range: int = 0, training: bool = False
                           ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], torch.Tensor, Optional[torch.Tensor]]:
    """Returns (position_ids, attention_mask, inputs_embeds [B, Lmax, D], labels) with the reference's None rules
    (llava_arch.py:857-867).  image_features: the per-sample vision sequences ([L_i, D], same device / dtype as
    embed_table), consumed in order at each IMAGE_TOKEN_INDEX like the reference's cur_image_idx."""
    dev = embed_table.device
    if not embed_table.is_cuda:
        raise RuntimeError("mavlm.splice: embed_table is not on a CUDA device; this path has no CPU fallback")
    if embed_table.dim() != 2 or not embed_table.is_contiguous():
        raise RuntimeError("mavlm.splice: embed_table must be a contiguous [vocab, D] tensor")
    vocab, d = embed_table.shape
    # the kernel reinterprets both sources with ONE dtype code: features in another dtype (an fp32 projector feeding a
    # bf16 LLM) are cast here, like the reference's `.to(self.device)` / torch.cat promotion would force the caller to
    image_features = [f if f.dtype == embed_table.dtype else ops.cast(f.contiguous(), embed_table.dtype)
                      for f in image_features]
    for f in image_features:
        if f.device != dev or f.dim() != 2 or f.shape[1] != d:
            raise RuntimeError(f"mavlm.splice: image features must be [L, {d}] tensors on {dev}, got {tuple(f.shape)} "
                               f"on {f.device}")
    ids_cpu = input_ids.detach().cpu()
    _labels, _position_ids, _attention_mask = labels, position_ids, attention_mask
    mask_cpu = (torch.ones_like(ids_cpu, dtype=torch.bool) if attention_mask is None
                else attention_mask.detach().cpu().bool())                                     # :749-752
    labels_cpu = torch.full_like(ids_cpu, IGNORE_INDEX) if labels is None else labels.detach().cpu()   # :755-756
    feat_rows = [int(f.shape[0]) for f in image_features]
    feat_base = [0]
    for n in feat_rows:
        feat_base.append(feat_base[-1] + n)
    feats = image_features[0] if len(image_features) == 1 else torch.cat(list(image_features), dim=0)
    feats = feats.contiguous()

    src_rows: List[torch.Tensor] = []
    new_labels: List[torch.Tensor] = []
    cur_image_idx = 0
    for b in range(ids_cpu.shape[0]):
        cur_ids = ids_cpu[b][mask_cpu[b]]                                                      # :760
        cur_lab = labels_cpu[b][mask_cpu[b]]
        img_pos = torch.where(cur_ids == IMAGE_TOKEN_INDEX)[0].tolist()
        if not img_pos:                                                                        # :770-778
            src_rows.append(cur_ids.clone())
            new_labels.append(cur_lab)
            cur_image_idx += 1
            continue
        bounds = [-1] + img_pos + [cur_ids.shape[0]]                                           # :780
        src, lab = [], []
        for i in range(len(bounds) - 1):
            src.append(cur_ids[bounds[i] + 1:bounds[i + 1]])                                   # text chunk -> table rows
            lab.append(cur_lab[bounds[i] + 1:bounds[i + 1]])
            if i < len(img_pos):
                fi = cur_image_idx if cur_image_idx < len(feat_rows) else cur_image_idx - 1    # :799-802 (IndexError hack)
                cur_image_idx += 1
                n = feat_rows[fi]
                src.append(-(torch.arange(n, dtype=torch.int64) + feat_base[fi]) - 2)
                lab.append(torch.full((n,), IGNORE_INDEX, dtype=cur_lab.dtype))                # :805
        src_rows.append(torch.cat(src))
        new_labels.append(torch.cat(lab))
    if tokenizer_model_max_length is not None:                                                 # :822-823
        src_rows = [s[:tokenizer_model_max_length] for s in src_rows]
        new_labels = [l[:tokenizer_model_max_length] for l in new_labels]
    max_len = max(s.shape[0] for s in src_rows)
    bsz = len(src_rows)
    table = torch.full((bsz, max_len), -1, dtype=torch.int64)
    labels_pad = torch.full((bsz, max_len), IGNORE_INDEX, dtype=new_labels[0].dtype)
    mask_pad = torch.zeros((bsz, max_len), dtype=torch.bool)
    pos_pad = torch.zeros((bsz, max_len), dtype=torch.long if position_ids is None else position_ids.dtype)
    for i, (s, l) in enumerate(zip(src_rows, new_labels)):                                     # :838-851
        n = s.shape[0]
        if n == 0:
            continue
        sl = slice(max_len - n, max_len) if padding_side == "left" else slice(0, n)
        table[i, sl] = s
        labels_pad[i, sl] = l
        mask_pad[i, sl] = True
        pos_pad[i, sl] = torch.arange(n, dtype=pos_pad.dtype)
    text = table[table >= 0]
    if text.numel() and int(text.max()) >= vocab:                                              # nn.Embedding raises here too
        raise IndexError(f"mavlm.splice: token id {int(text.max())} is out of range for an embedding table of {vocab} rows")
    if int(table.min()) < -(feat_base[-1] + 1):
        raise RuntimeError("mavlm.splice: inconsistent row-source table")                      # pragma: no cover
    out = torch.empty((bsz, max_len, d), dtype=embed_table.dtype, device=dev)
    tab_dev = table.to(dev)
    st = _lib.load().mavlm_gather_rows_fwd(out.data_ptr(), d, embed_table.data_ptr(), vocab, feats.data_ptr(),
                                           feats.shape[0], tab_dev.data_ptr(), bsz * max_len, d, ops.dtype_code(out),
                                           torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "gather_rows_fwd")
    new_labels_out = None if _labels is None else labels_pad.to(dev)                           # :857-860
    mask_out = None if _attention_mask is None else mask_pad.to(device=dev, dtype=_attention_mask.dtype)   # :862-865
    pos_out = None if _position_ids is None else pos_pad.to(dev)                               # :867-868
    if use_pos_skipping and training:                                                          # :869-875
        pos_out = torch.arange(max_len, device=dev).unsqueeze(0)
        split_position = random.randint(0, max_len)
        left_add = random.randint(0, pos_skipping_range)
        right_add = random.randint(left_add, pos_skipping_range)
        pos_out[:, :split_position] += left_add
        pos_out[:, split_position:] += right_add
    return pos_out, mask_out, out, new_labels_out
