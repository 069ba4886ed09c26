"""CPU-side checks: the C-ABI library loads and exports every symbol include/mavlm.h declares (no
compute calls without a GPU), host logic (index math, state_dict keys, packing, error behaviour)."""
import json
import os
import re

import numpy as np
import pytest
import torch

import mavlm_b200 as M
from mavlm_b200 import _lib, ops, synthetic

from conftest import GOLDEN, ROOT


def _header_symbols():
    with open(os.path.join(ROOT, "include", "mavlm.h")) as fh:
        txt = fh.read()
    return sorted(set(re.findall(r"MAVLM_API\s+[\w\s\*]+?\b(mavlm_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 10
    lib = _lib.load()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mavlm.h but not exported by libmavlm.so"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert lib.mavlm_version() == 100
    assert lib.mavlm_launch_count() == 0          # nothing has been launched: no compute on the CPU box
    assert lib.mavlm_xattn_workspace_bytes(1, 8, 1568, 6272, 112, _lib.F32) == 8 * 1568 * 6272 * 4
    # bf16 tier: one partial-result slot + one completion flag per persistent CTA: 11 groups of 13 q-tile CTAs
    assert lib.mavlm_xattn_workspace_bytes(1, 8, 1568, 6272, 448, _lib.BF16) == 11 * 13 * ((128 * 448 + 256) * 4 + 4)


def test_no_cpu_fallback():
    x = torch.randn(4, 8)
    w = torch.randn(8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.linear(x, w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.pool_pe(torch.randn(1, 729, 8), side=27)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    # bad dtype / bad pool mode are rejected before any CUDA call
    assert lib.mavlm_pool_pe_fwd(None, None, None, None, 1, 27, 14, 2, 8, 7, _lib.F32, None) == _lib.E_INVALID
    assert b"mm_spatial_pool_mode" in lib.mavlm_last_error_string()
    assert lib.mavlm_gemm_bias_act_fwd(None, 0, None, 0, None, None, 0, None, None, 0, 4, 4, 0, 0, _lib.F32, _lib.F32,
                                       None) == _lib.E_INVALID
    with pytest.raises(ValueError, match="Unexpected mm_spatial_pool_mode"):
        M.get_2dPool(torch.zeros(1, 729, 8), mode="nearest")           # llava_arch.py:294 raises the same
    # SURVEY 8f-4 entry points: geometry is checked before any CUDA call
    assert lib.mavlm_stream_compress_fwd(None, 10, 64, 0, 1, None, None, None, None, None, 0, _lib.F32, None) == _lib.E_INVALID
    assert b"keep" in lib.mavlm_last_error_string()
    assert lib.mavlm_stream_compress_fwd(None, 10, 64, 65, 1, None, None, None, None, None, 0, _lib.F32, None) == _lib.E_INVALID
    assert lib.mavlm_stream_compress_fwd(None, 10, 63, 3, 4, None, None, None, None, None, 0, _lib.F32, None) == _lib.E_INVALID
    assert b"mode" in lib.mavlm_last_error_string()
    assert lib.mavlm_stream_compress_fwd(None, 10, 62, 3, 1, None, None, None, None, None, 0, _lib.BF16, None) == _lib.E_INVALID
    assert b"multiple of 8" in lib.mavlm_last_error_string()
    assert lib.mavlm_stream_compress_workspace_bytes(64, 65, 1, _lib.F32) == 0
    assert lib.mavlm_avg_pool_fwd(None, None, 1, 6, 7, 8, _lib.F32, None) == _lib.E_INVALID
    assert lib.mavlm_frame_mean_fwd(None, None, 1, 4, 6, _lib.BF16, None) == _lib.E_INVALID
    assert lib.mavlm_ntm_softmax_fwd(None, 8, 1, 9, 1.0, 1.0, None, 8, 8, None, None, 0, _lib.F32, None) == _lib.E_INVALID
    assert lib.mavlm_kmeans_iter_fwd(None, None, None, None, None, None, None, None, 4, 62, 2, None, 0, _lib.BF16,
                                     None) == _lib.E_INVALID
    assert lib.mavlm_depth_scores_fwd(None, None, 0, 0, None) == 0                          # nothing to do


def test_state_dict_keys_match_reference():
    z = np.load(os.path.join(GOLDEN, "rmt_small.npz"))
    ref_keys = sorted(k[len("w::recurrent_memory_transformer."):] for k in z.files if k.startswith("w::"))
    cfg = M.Config()
    cfg.mm_hidden_size, cfg.mm_intermediate_size, cfg.depth, cfg.mm_dtype = 32, 128, 2, torch.float32
    rmt = M.TransformerProjector(cfg)
    assert sorted(rmt.state_dict().keys()) == ref_keys
    for k, v in rmt.state_dict().items():
        assert tuple(v.shape) == z["w::recurrent_memory_transformer." + k].shape, k
    zf = np.load(os.path.join(GOLDEN, "projector_fuser.npz"))
    import types
    proj = M.build_vision_projector(types.SimpleNamespace(mm_projector_type="mlp2x_gelu", mm_hidden_size=48,
                                                          hidden_size=32))
    assert sorted("mm_projector." + k for k in proj.state_dict()) == sorted(
        k[3:] for k in zf.files if k.startswith("w::mm_projector."))
    fuser = M.build_memory_fuser(32)
    assert sorted("memory_fuser." + k for k in fuser.state_dict()) == sorted(
        k[3:] for k in zf.files if k.startswith("w::memory_fuser."))
    with pytest.raises(ValueError):
        M.build_vision_projector(types.SimpleNamespace(mm_projector_type="qformer", mm_hidden_size=48, hidden_size=32))


def test_index_math_matches_reference():
    with open(os.path.join(GOLDEN, "indices.json")) as fh:
        g = json.load(fh)
    for f, idx in g["sample"].items():
        assert M.sample_frame_indices(int(f)).tolist() == idx, f
    for n, idx in g["fine"].items():
        assert M.fine_frame_indices(int(n)).tolist() == idx, n
    for key, b in g["bounds"].items():
        t, d = map(int, key.split(","))
        assert M.uniform_segment_variant(torch.zeros(t, 1), d=d) == b, key
        assert M.uniform_segment_variant(t, d) == b, key


def test_pe_table_and_index_errors():
    z = np.load(os.path.join(GOLDEN, "pe.npz"))
    pe = M.TemporalPositionalEncoding(600, 32, learnable=False)
    assert np.array_equal(pe.frame_embed.numpy(), z["table32"])         # same torch ops as the reference: bit equal
    assert "frame_embed" in pe.state_dict()
    with pytest.raises(ValueError, match="exceed max_frames"):
        pe.validate(torch.tensor([0, 600]))
    with pytest.raises(ValueError, match="negative"):
        pe.validate(torch.tensor([-1, 3]))
    with pytest.raises(ValueError, match="3D or 4D"):
        pe(torch.zeros(4, 4))


def test_padded_head_packing_is_exact():
    """0.5B heads (112) are zero-padded to 128 for the tensor-core tier: packed weights must give the
    same q/k/v/out as the unpadded ones (checked with plain matmuls on the CPU)."""
    cfg = M.Config()
    cfg.mm_hidden_size, cfg.mm_intermediate_size, cfg.mm_dtype = 896, 3584, torch.bfloat16
    torch.manual_seed(0)
    att = M.Attention(cfg)
    pg = att.packed()                                                  # grad enabled: differentiable, not cached
    assert pg["wq"].requires_grad and pg["wkv"].requires_grad and att.packed() is not pg
    torch.set_grad_enabled(False)
    p = att.packed()
    assert p["dhp"] == 128 and p["wq"].shape == (1024, 896) and p["wkv"].shape == (2048, 896)
    x = torch.randn(5, 896).bfloat16()
    q = (x.float() @ att.q_proj.weight.float().T + att.q_proj.bias.float()).view(5, 8, 112)
    qp = (x.float() @ p["wq"].float().T + p["bq"].float()).view(5, 8, 128)
    assert torch.allclose(qp[..., :112], q, atol=1e-5) and torch.count_nonzero(qp[..., 112:]) == 0
    ctx = torch.randn(5, 8, 112)
    ctxp = torch.zeros(5, 8, 128)
    ctxp[..., :112] = ctx
    o = ctx.reshape(5, 896) @ att.residual.dense.weight.float().T
    op = ctxp.reshape(5, 1024) @ p["wo"].float().T
    assert torch.allclose(o, op, atol=1e-5)
    assert att.packed() is p and not p["wq"].requires_grad             # inference: cached until a weight changes
    att.q_proj.weight.mul_(2.0)
    assert att.packed() is not p
    torch.set_grad_enabled(True)
    cfg32 = M.Config()
    cfg32.mm_hidden_size, cfg32.mm_intermediate_size, cfg32.mm_dtype = 896, 3584, torch.float32
    assert M.Attention(cfg32).packed()["dhp"] == 112                   # fp32 tier: no padding


def test_cache_reset_semantics():
    cfg = M.Config()
    cfg.mm_hidden_size, cfg.mm_intermediate_size, cfg.mm_dtype = 16, 64, torch.float32
    rmt = M.TransformerProjector(cfg)
    rmt._memory_cache.append(torch.zeros(8, 196, 16))
    rmt.frame_attn_scores.append(torch.zeros(2))
    rmt._kv_cache.append(torch.zeros(1))
    rmt.memory_cache = []                                              # llava_arch.py:532
    assert rmt.memory_cache == [] and rmt._kv_cache == [] and rmt.frame_attn_scores == []


def test_sequence_length_formula():
    pipe, _ = synthetic.build_pipeline(16, 4, dtype=torch.float32, device="cpu", vocab=64)
    assert pipe.sequence_length(2, 32) == 10 + 2 * 1568 + 1 + 9 + 32 * 196 + 1 == 9429   # SURVEY.md §3.1 (+4 text = 9433)
    assert pipe.sequence_length(1, 5) == 10 + 1568 + 1 + 9 + 5 * 196 + 1
    assert pipe.sequence_length(3, 32, drop_frames=True) == 10 + 3 * 1568 + 1


def test_resize_coefficient_tables_match_the_oracle():
    """mavlm_resize_coeffs is host code (no GPU): Pillow's precompute_coeffs + normalize_coeffs_8bpc, bit for bit."""
    import numpy as np
    from oracle import preprocess_oracle as po
    lib = _lib.load()
    for in_size, out_size in ((640, 384), (1280, 384), (200, 384), (97, 384), (384, 384), (1080, 384), (1, 7), (500, 3)):
        k = lib.mavlm_resize_coeffs(in_size, out_size, None, None, 0)
        bounds = np.zeros((out_size, 2), dtype=np.int32)
        kk = np.zeros((out_size, k), dtype=np.int32)
        assert lib.mavlm_resize_coeffs(in_size, out_size, bounds.ctypes.data, kk.ctypes.data, k) == k
        k2, b2, kk2 = po.precompute_coeffs(in_size, out_size)
        assert k == k2 and np.array_equal(bounds, b2) and np.array_equal(kk, kk2)
    assert lib.mavlm_resize_coeffs(0, 4, None, None, 0) < 0                     # error code, never an exception
