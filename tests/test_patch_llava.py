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


def test_patch_rebinds_the_legacy_memory_methods():
    """A model that mixes in the reference-style MultimodalOpsMixin and owns a Turing-memory module gets the B200
    methods and the drop-in NeuralTuringMachine on the same parameters."""
    from mavlm_b200 import legacy

    class _RefNTM(nn.Module):                                 # parameter container with memory_builder.py:8-19's names
        def __init__(self, d=16):
            super().__init__()
            self.input_dim = self.output_dim = d
            self.q_proj, self.k_proj, self.v_proj = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)
            self.dropout, self.out_proj, self.out_dropout = nn.Dropout(0.1), nn.Linear(d, d), nn.Dropout(0.1)
            self.out_ln = nn.LayerNorm(d, eps=1e-12)

    class _LegacyModel(_Model):
        def compress_temporal_features(self, image_features, video_idx_in_batch, all_video=False):
            raise AssertionError("the reference method should have been replaced")

        def compress_spatial_features(self, image_features, compress_size=1):
            raise AssertionError("the reference method should have been replaced")

        def attention(self, turing_memory, new_feature, update_ratio=0.2):
            raise AssertionError("the reference method should have been replaced")

    m = _LegacyModel()
    m.get_model().attention_model = _RefNTM().eval()
    before = {k: v.data_ptr() for k, v in m.state_dict().items()}
    M.patch_llava(m)
    assert {k: v.data_ptr() for k, v in m.state_dict().items()} == before
    assert isinstance(m.get_model().attention_model, legacy.NeuralTuringMachine)
    assert not m.get_model().attention_model.training
    assert m.compress_temporal_features.__func__ is legacy.MultimodalOpsMixin.compress_temporal_features
    x = torch.zeros(2, 36, 8)
    assert m.compress_spatial_features(x, 6) is x            # 6 x 6 already: returned untouched, like the reference
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.compress_spatial_features(x, 3)
    with pytest.raises(NotImplementedError):
        m.config.video_sample_type = "bogus"
        m.compress_temporal_features([x], [0])


def test_patch_adopts_the_encoder_variant_fuser():
    class _RefEncoderFuser(nn.Module):                        # MemoryFuser.py:4-22's attribute names
        def __init__(self, d):
            super().__init__()
            self.input_proj = nn.Linear(d, d)
            layer = nn.TransformerEncoderLayer(d_model=d, nhead=4, dim_feedforward=4 * d, dropout=0.1, batch_first=True,
                                               activation="gelu")
            self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=2, enable_nested_tensor=False)
            self.output_proj = nn.Linear(d, d)

    m = _Model()
    m.get_model().memory_fuser = _RefEncoderFuser(32).eval()
    before = {k: v.data_ptr() for k, v in m.state_dict().items()}
    M.patch_llava(m)
    assert {k: v.data_ptr() for k, v in m.state_dict().items()} == before
    f = m.get_model().memory_fuser
    assert isinstance(f, M.MemoryFuser) and f.num_heads == 4 and len(f.transformer_encoder.layers) == 2 and not f.training
    assert m.mavlm_pipeline.memory_fuser is f
    m.get_model().memory_fuser = nn.Identity()
    with pytest.raises(TypeError):
        M.patch_llava(m)


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


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["full_path.npz", "full_path_short.npz"])
def test_fused_prepare_inputs_matches_the_reference_function(name):
    """patch_llava(fused=True) replaces prepare_inputs_labels_for_multimodal as a whole: its inputs_embeds equal what
    the REAL reference function produced for the same video, ids and weights (tests/golden/full_path*.npz were
    generated by running llava_arch.py's method through tools/gen_golden.py)."""
    import os
    import numpy as np
    from conftest import GOLDEN
    z0 = np.load(os.path.join(GOLDEN, "full_path.npz"))
    w = {k[3:]: z0[k] for k in z0.files if k.startswith("w::")}
    z = np.load(os.path.join(GOLDEN, name))
    d, dv = 16, 4

    class Tower(nn.Module):                                   # the golden harness's fake tower: [N, C, 27, 27] -> [N, 729, C]
        num_patches_per_side = 27

        def forward(self, images):
            return images.flatten(2).transpose(1, 2)

    class Inner(nn.Module):
        def __init__(self):
            super().__init__()
            self.vision_tower = Tower()
            self.mm_projector = nn.Sequential(nn.Linear(dv, d), nn.GELU(), nn.Linear(d, d))
            self.recurrent_memory_transformer = _RefLikeRMT(d)
            self.memory_fuser = nn.Sequential(nn.Linear(d, 4 * d), nn.GELU(), nn.Linear(4 * d, d))
            self.positional_encoding = M.TemporalPositionalEncoding(600, d, learnable=False)
            self.token_type_embedding = nn.Embedding(2, d)
            self.image_newline = nn.Parameter(torch.zeros(d))
            self.embed_tokens = nn.Embedding(50000, d)

    class Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.model = Inner()
            self.config = types.SimpleNamespace(mm_spatial_pool_mode="bilinear", tokenizer_model_max_length=32768,
                                                tokenizer_padding_side="right", mm_patch_merge_type="spatial_unpad",
                                                mm_newline_position="one_token")    # the golden harness's setting

        def get_model(self):
            return self.model

        def get_vision_tower(self):
            return self.model.vision_tower

    m = Model()
    inner = m.model
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    inner.mm_projector.load_state_dict({k[len("mm_projector."):]: v for k, v in sd.items() if k.startswith("mm_projector.")})
    inner.recurrent_memory_transformer.load_state_dict(
        {k[len("recurrent_memory_transformer."):]: v for k, v in sd.items() if k.startswith("recurrent_memory_transformer.")})
    inner.memory_fuser.load_state_dict({k[len("memory_fuser."):]: v for k, v in sd.items() if k.startswith("memory_fuser.")})
    inner.token_type_embedding.load_state_dict({"weight": sd["token_type_embedding.weight"]})
    inner.image_newline.data = sd["image_newline"]
    tab = sd["_emb.weight"]
    inner.embed_tokens.weight.data = tab[torch.arange(50000) % tab.shape[0]]          # the harness embeds ids mod 64
    m = m.cuda().eval()
    M.patch_llava(m, fused=True)
    ids = torch.from_numpy(z0["input_ids"]).cuda()
    video = torch.from_numpy(z["video"]).cuda()
    out = m.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, [video], modalities=["video"])
    assert out[0] is None and out[1] is None and out[2] is None and out[5] is None         # the reference's None rules
    ref = torch.from_numpy(z["inputs_embeds"])
    emb = out[4].float().cpu()
    assert emb.shape == ref.shape
    assert float((emb - ref).abs().max() / ref.abs().max()) < 2e-5
    # decode step: one new token -> untouched inputs (llava_arch.py:392-394)
    same = m.prepare_inputs_labels_for_multimodal(ids[:, :1], None, None, "kv", None, [video], modalities=["video"])
    assert same[0] is ids[:, :1] or torch.equal(same[0], ids[:, :1])
    assert same[3] == "kv" and same[4] is None


def test_fused_prepare_routes_other_layouts_to_the_reference_method():
    """The fused whole-function replacement only emits the spatial_unpad + one_token + bilinear layout; 'flat' (the
    reference's DEFAULT merge type: no newline, llava_arch.py:562-568), 'spatial' without 'unpad', the other newline
    positions and average / max pooling must reach the reference's own method instead of silently producing a
    sequence of different length.  CPU: the dispatch happens before any kernel call."""
    d = 16

    class Tower(nn.Module):
        num_patches_per_side = 27

        def forward(self, images):
            return images.flatten(2).transpose(1, 2)

    class Inner(nn.Module):
        def __init__(self):
            super().__init__()
            self.vision_tower = Tower()
            self.mm_projector = nn.Sequential(nn.Linear(4, d), nn.GELU(), nn.Linear(d, d))
            self.recurrent_memory_transformer = _RefLikeRMT(d)
            self.memory_fuser = nn.Sequential(nn.Linear(d, 4 * d), nn.GELU(), nn.Linear(4 * d, d))
            self.positional_encoding = M.TemporalPositionalEncoding(600, d, learnable=False)
            self.token_type_embedding = nn.Embedding(2, d)
            self.image_newline = nn.Parameter(torch.zeros(d))
            self.embed_tokens = nn.Embedding(100, d)

    class Model(nn.Module):
        def __init__(self):
            super().__init__()
            self.model = Inner()
            self.config = types.SimpleNamespace(mm_spatial_pool_mode="bilinear")

        def get_model(self):
            return self.model

        def get_vision_tower(self):
            return self.model.vision_tower

        def prepare_inputs_labels_for_multimodal(self, *a, **k):          # stands in for llava_arch.py:388
            return "reference"

    m = Model().eval()
    M.patch_llava(m, fused=True)
    ids = torch.tensor([[1, -200, 2]])
    video = torch.zeros(4, 4, 27, 27)
    call = lambda: m.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, [video], modalities=["video"])
    assert call() == "reference"                                          # mm_patch_merge_type defaults to 'flat'
    for merge, newline, pool in (("flat", "one_token", "bilinear"), ("spatial", "one_token", "bilinear"),
                                 ("spatial_unpad", "frame", "bilinear"), ("spatial_unpad", "grid", "bilinear"),
                                 ("spatial_unpad", "no_token", "bilinear"), ("spatial_unpad", "one_token", "average"),
                                 ("spatial_unpad", "one_token", "max")):
        m.config = types.SimpleNamespace(mm_spatial_pool_mode=pool, mm_patch_merge_type=merge, mm_newline_position=newline)
        assert call() == "reference", (merge, newline, pool)
    m.config = types.SimpleNamespace(mm_spatial_pool_mode="bilinear", mm_patch_merge_type="spatial_unpad",
                                     mm_newline_position="one_token")
    with pytest.raises(RuntimeError, match="no CPU fallback"):              # the fused path is taken (and needs the GPU)
        call()
