"""Text / vision splice (llava_arch.py:745-878) against goldens from the reference's own
prepare_inputs_labels_for_multimodal: labels / attention mask / position ids bit-exact, embeddings row-exact."""
import os

import numpy as np
import pytest
import torch

import mavlm_b200 as M

from conftest import GOLDEN

DEV = "cuda:0"


@pytest.mark.gpu
@pytest.mark.parametrize("tag,side,maxlen,row", [("right", "right", 32768, 0), ("left_trunc", "left", 9000, 1)])
def test_splice_matches_reference(tag, side, maxlen, row):
    z = np.load(os.path.join(GOLDEN, "splice.npz"))
    table = torch.from_numpy(z["embed_table"]).to(DEV)
    feats = torch.from_numpy(z["video_sequence"]).to(DEV)
    ids = torch.from_numpy(z["input_ids"][row:row + 1]).to(DEV)
    pos, mask, emb, lab = M.splice_text_and_vision(
        ids, torch.from_numpy(z["pos"][row:row + 1]).to(DEV), torch.from_numpy(z["mask"][row:row + 1]).to(DEV),
        torch.from_numpy(z["labels"][row:row + 1]).to(DEV), [feats], table, tokenizer_model_max_length=maxlen,
        padding_side=side)
    assert tuple(emb.shape) == tuple(z[f"{tag}.shape"])
    assert np.array_equal(pos.cpu().numpy(), z[f"{tag}.position_ids"])
    assert np.array_equal(mask.cpu().numpy(), z[f"{tag}.attention_mask"])
    assert np.array_equal(lab.cpu().numpy(), z[f"{tag}.labels"])
    e = emb.cpu()
    assert np.array_equal(e[:, :12].numpy(), z[f"{tag}.rows_head"])
    assert np.array_equal(e[:, -12:].numpy(), z[f"{tag}.rows_tail"])
    assert np.array_equal(e[:, ::97].numpy(), z[f"{tag}.rows_stride"])
    assert np.allclose(e.double().sum(dim=1).numpy(), z[f"{tag}.colsum"], rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_splice_none_rules_and_text_only_and_batch_padding():
    torch.manual_seed(0)
    table = torch.randn(100, 32, device=DEV).bfloat16()
    feats = torch.randn(7, 32, device=DEV).bfloat16()
    ids = torch.tensor([[3, M.IMAGE_TOKEN_INDEX, 4, 5], [6, 7, 0, 0]], device=DEV)
    mask = torch.tensor([[1, 1, 1, 1], [1, 1, 0, 0]], device=DEV)
    pos, m2, emb, lab = M.splice_text_and_vision(ids, None, mask, None, [feats], table)
    assert pos is None and lab is None                                         # llava_arch.py:857-868
    assert emb.shape == (2, 10, 32) and m2.dtype == mask.dtype
    assert m2.sum(1).tolist() == [10, 2]
    assert torch.equal(emb[0, 0], table[3]) and torch.equal(emb[0, 1:8], feats) and torch.equal(emb[0, 8], table[4])
    assert torch.equal(emb[1, :2], table[torch.tensor([6, 7], device=DEV)]) and emb[1, 2:].abs().sum() == 0
    pos, m2, emb, lab = M.splice_text_and_vision(ids, None, mask, None, [feats], table, padding_side="left")
    assert emb[1, :8].abs().sum() == 0 and torch.equal(emb[1, 8], table[6])
    # features in another dtype are cast (never reinterpreted bytewise); out-of-range ids raise like nn.Embedding
    _, _, emb32, _ = M.splice_text_and_vision(ids, None, mask, None, [feats.float()], table)
    assert torch.equal(emb32, M.splice_text_and_vision(ids, None, mask, None, [feats], table)[2])
    bad = torch.tensor([[3, M.IMAGE_TOKEN_INDEX, 100, 5]], device=DEV)
    with pytest.raises(IndexError):
        M.splice_text_and_vision(bad, None, None, None, [feats], table)
    with pytest.raises(RuntimeError):
        M.splice_text_and_vision(ids, None, mask, None, [feats[:, :16]], table)                    # wrong width
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        M.splice_text_and_vision(ids.cpu(), None, mask.cpu(), None, [feats.cpu()], table.cpu())
