"""patch_llava on a reference-SHAPED model built from plain torch modules (the reference itself is not
available on the GPU box): attribute names / state_dict keys are the reference's (llava_arch.py:117-150)."""
import types

import pytest
import torch
from torch import nn

import mavlm_b200 as M


class _RefLikeRMT(nn.Module):
    """Parameter container with the reference TransformerProjector's key names (MemoryController.py:74-87)."""

    def __init__(self, d):
        super().__init__()
        cfg = M.Config()
        cfg.mm_hidden_size, cfg.mm_intermediate_size, cfg.depth, cfg.mm_dtype = d, 4 * d, 2, torch.float32
        self.config = cfg
        donor = M.TransformerProjector(cfg)
        self.layers = donor.layers
        self.memory_update_attention = donor.memory_update_attention
        self.initial_memory = donor.initial_memory
        self.memory_pos_embed = donor.memory_pos_embed


class _Tower(nn.Module):
    num_patches_per_side = 27


class _Inner(nn.Module):
    def __init__(self, d=32, dv=8):
        super().__init__()
        self.vision_tower = _Tower()
        self.mm_projector = nn.Sequential(nn.Linear(dv, d), nn.GELU(), nn.Linear(d, d))
        self.recurrent_memory_transformer = _RefLikeRMT(d)
        self.memory_fuser = nn.Sequential(nn.Linear(d, 4 * d), nn.GELU(), nn.Linear(4 * d, d))
        self.positional_encoding = M.TemporalPositionalEncoding(600, d, learnable=False)
        self.token_type_embedding = nn.Embedding(2, d)
        self.image_newline = nn.Parameter(torch.randn(d))
        self.embed_tokens = nn.Embedding(50000, d)


class _Model(nn.Module):
    def __init__(self):
        super().__init__()
        self.model = _Inner()
        self.config = types.SimpleNamespace(mm_spatial_pool_mode="bilinear")

    def get_model(self):
        return self.model

    def get_vision_tower(self):
        return self.model.vision_tower


def test_patch_keeps_state_dict_and_shares_parameters():
    m = _Model()
    before = {k: v.data_ptr() for k, v in m.state_dict().items()}
    M.patch_llava(m)
    after = {k: v.data_ptr() for k, v in m.state_dict().items()}
    assert before == after                                   # same keys, same storage: checkpoints load/save unchanged
    inner = m.get_model()
    assert isinstance(inner.recurrent_memory_transformer, M.TransformerProjector)
    assert isinstance(inner.mm_projector, M.VisionProjector) and isinstance(inner.memory_fuser, M.MemoryFuserMLP)
    assert isinstance(m.mavlm_pipeline, M.VisualMemoryPipeline)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.get_2dPool(torch.zeros(1, 729, 32))
    m.config.mm_spatial_pool_mode = "nearest"
    with pytest.raises(ValueError, match="Unexpected mm_spatial_pool_mode"):
        m.get_2dPool(torch.zeros(1, 729, 32))


def test_patch_accepts_fp16_and_rejects_other_dtypes():
    m = _Model().half()                                      # the reference inference loader's default (builder.py:27)
    M.patch_llava(m)
    assert next(m.get_model().recurrent_memory_transformer.parameters()).dtype == torch.float16
    with pytest.raises(TypeError):
        M.patch_llava(_Model().double())


@pytest.mark.gpu
def test_patched_reference_loop_runs_on_gpu():
    """The reference's own per-video loop (llava_arch.py:491-557) over the patched attributes."""
    m = _Model().cuda()
    M.patch_llava(m, chunk_size=4)
    inner = m.get_model()
    feats = inner.mm_projector(torch.randn(6, 729, 8, device="cuda"))
    image = m.get_2dPool(feats)
    image = inner.positional_encoding(image, torch.arange(6))
    rmt = inner.recurrent_memory_transformer
    rmt.memory_cache = []
    for seg in (image[0:4], image[4:6]):
        cache, _ = rmt(seg)
    fused = inner.memory_fuser(torch.cat(cache, dim=0))
    assert fused.shape == (16, 196, 32) and torch.isfinite(fused).all()
    both = m.mavlm_pipeline.memory_forward(image[None])      # fused path == module-by-module path
    ref = fused + inner.token_type_embedding.weight[0]
    got = both["sequence"][0, 10:10 + 16 * 196].view(16, 196, 32)
    assert (got - ref).abs().max() < 1e-4
