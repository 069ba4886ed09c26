#!/usr/bin/env python
"""Generate tests/golden/legacy_memory.npz by EXECUTING THE UNMODIFIED REFERENCE (scene
segmentation, Flash-VStream-style compression, Turing memory) in this container.

    python tools/gen_golden_legacy.py

Inputs are rebuilt from seeds by `legacy_inputs()` below (imported by the tests), so the fixture only holds the
reference's outputs.  Random decisions of the reference come from `random` / `torch.randperm`; each case seeds both
right before the call and the tests replay the same streams.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import random
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import OUT, REF, _install_llava_stub  # noqa: E402


def scene_frames(seed: int, T: int, P: int, D: int, scenes: int = 4) -> np.ndarray:
    """Frames that fall into `scenes` runs of similar content with a different noise level per frame, so that the
    adjacent similarities are well separated (decisions do not hinge on rounding)."""
    g = np.random.default_rng(seed)
    base = g.standard_normal((scenes, P, D)).astype(np.float32)
    cuts = np.sort(g.choice(np.arange(1, T), size=scenes - 1, replace=False))
    scene_of = np.searchsorted(cuts, np.arange(T), side="right")
    noise = g.uniform(0.15, 0.9, size=T).astype(np.float32)
    x = base[scene_of] + noise[:, None, None] * g.standard_normal((T, P, D)).astype(np.float32)
    return x.astype(np.float32)


def legacy_inputs():
    return {
        "seg_feat": scene_frames(11, 40, 1, 32, scenes=5)[:, 0],
        "seg_feat_long": scene_frames(12, 300, 1, 24, scenes=9)[:, 0],
        "sim_rand": np.random.default_rng(13).uniform(0.2, 1.0, size=64).astype(np.float32),
        "scene_video": scene_frames(14, 50, 4, 16, scenes=6),
        "scene_video_busy": scene_frames(15, 60, 3, 8, scenes=24),
        "stream": scene_frames(16, 14, 4, 16, scenes=5),
        "ntm_frames": (0.5 * scene_frames(17, 8, 5, 16, scenes=3)).astype(np.float32),
        "spatial36": np.random.default_rng(18).standard_normal((3, 36, 8)).astype(np.float32),
        "spatial729": np.random.default_rng(19).standard_normal((2, 729, 8)).astype(np.float32),
        "temporal": (0.3 * scene_frames(20, 8, 729, 1152, scenes=3)).astype(np.float32),
    }


def ntm_params(seed: int, d: int):
    g = np.random.default_rng(seed)
    p = {}
    for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
        p[f"{n}.weight"] = (g.standard_normal((d, d)) * (1.5 / np.sqrt(d))).astype(np.float32)
        p[f"{n}.bias"] = (0.1 * g.standard_normal(d)).astype(np.float32)
    p["out_ln.weight"] = (1.0 + 0.1 * g.standard_normal(d)).astype(np.float32)
    p["out_ln.bias"] = (0.1 * g.standard_normal(d)).astype(np.float32)
    return p


def mlp_params(seed: int, d: int):
    g = np.random.default_rng(seed)
    return {"0.weight": (g.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32),
            "0.bias": (0.1 * g.standard_normal(d)).astype(np.float32),
            "2.weight": (g.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32),
            "2.bias": (0.1 * g.standard_normal(d)).astype(np.float32)}


TEMPORAL_SAMPLE = (slice(None), slice(0, 729, 31), slice(0, 1152, 29))


def _steps_json(steps):
    return json.dumps(steps)


def main():
    _install_llava_stub()
    import importlib
    seg = importlib.import_module("llava.model.memory_module.segment")
    cf = importlib.import_module("llava.model.memory_module.compress_functions")
    mb = importlib.import_module("llava.model.memory_module.memory_builder")
    inp = legacy_inputs()
    out = {}
    meta = {}
    t = torch.from_numpy

    # ---- segmentation
    out["depth_rand"] = seg.cal_depth_score(t(inp["sim_rand"])).numpy()
    out["left_depth_rand"] = seg.cal_left_depth_score(t(inp["sim_rand"])).numpy()
    for name in ("seg_feat", "seg_feat_long"):
        f = t(inp[name])
        b, d = seg.segment(f, alpha=0.5)
        meta[f"{name}.segment_a05"] = b
        out[f"{name}.depth"] = d.numpy()
        meta[f"{name}.segment_k3"] = seg.segment(f, k=3)[0]
        with contextlib.redirect_stdout(io.StringIO()):
            meta[f"{name}.adjusted_4_8"] = seg.adjusted_segment(f, alpha=0.5, min_distance=4, max_distance=8)
            meta[f"{name}.adjusted_default"] = seg.adjusted_segment(f)
            meta[f"{name}.adjusted_k5"] = seg.adjusted_segment(f, k=5, min_distance=2, max_distance=30)
        meta[f"{name}.segment_left"] = seg.segment_left(f, alpha=0.5)
    meta["uniform_segment"] = {f"{T}_{d}": seg.uniform_segment(torch.zeros(T, 1), d)
                               for T in (1, 31, 32, 33, 64, 100) for d in (16, 32)}
    meta["uniform_segment_variant"] = {f"{T}_{d}": seg.uniform_segment_variant(torch.zeros(T, 1), d)
                                       for T in (1, 31, 32, 33, 64, 100) for d in (16, 32)}
    for name, nums in (("scene_video", (8, 20, 45, 50)), ("scene_video_busy", (4, 12))):
        for n in nums:
            torch.manual_seed(100 + n)
            meta[f"{name}.sample_{n}"] = [int(i) for i in seg.sample_scenes_priority(t(inp[name]), sample_num=n)]

    # ---- streaming compression
    x = t(inp["stream"])
    for T0 in (3, 5):
        random.seed(200 + T0)
        f, s, st = cf.drop_feature(x.clone(), T0)
        out[f"drop{T0}.feat"], out[f"drop{T0}.sim"], meta[f"drop{T0}.steps"] = f.numpy(), s.numpy(), st
        f, s, st = cf.merge_feature(x.clone(), T0)
        out[f"merge{T0}.feat"], out[f"merge{T0}.sim"], meta[f"merge{T0}.steps"] = f.numpy(), s.numpy(), st
        random.seed(300 + T0)
        f, s, st = cf.k_drop_feature(x.clone(), T0)
        out[f"kdrop{T0}.feat"], meta[f"kdrop{T0}.steps"] = f.numpy(), st
        f, s, st = cf.k_merge_feature(x.clone(), T0)
        out[f"kmerge{T0}.feat"], out[f"kmerge{T0}.sim"], meta[f"kmerge{T0}.steps"] = f.numpy(), s.numpy(), st
        torch.manual_seed(400 + T0)
        random.seed(400 + T0)
        f, s, st = cf.kmeans_feature(x.clone(), T0)
        out[f"kmeans{T0}.feat"], meta[f"kmeans{T0}.steps"] = f.numpy(), st
        torch.manual_seed(500 + T0)
        random.seed(500 + T0)
        w = torch.linspace(0.5, 2.0, x.shape[0])
        f, s, st = cf.weighted_kmeans_feature(x.clone(), T0, w)
        out[f"wkmeans{T0}.feat"], out[f"wkmeans{T0}.weights"], meta[f"wkmeans{T0}.steps"] = f.numpy(), s.numpy(), st
    f, s, st = cf.merge_feature(x[:3].clone(), 5)        # T <= T0: identity
    out["merge_identity.feat"], meta["merge_identity.steps"] = f.numpy(), st

    # ---- Turing memory
    d = 16
    p = ntm_params(600, d)
    ntm = mb.NeuralTuringMachine(input_dim=d, output_dim=d).eval()
    ntm.load_state_dict({k: t(v) for k, v in p.items()})
    fr = t(inp["ntm_frames"])
    with torch.no_grad():
        out["ntm.weight"] = ntm.get_weight(fr[:3].reshape(-1, d), fr[3:5].reshape(-1, d)).numpy()
        out["ntm.forward"] = ntm(fr[:3].reshape(-1, d), fr[3:5].reshape(-1, d)).numpy()

        class Holder(mb.MultimodalOpsMixin):
            def __init__(self, cfg, ntm_mod):
                super().__init__(cfg)
                self.attention_model = ntm_mod

            def get_model(self):
                return self

        h = Holder(types.SimpleNamespace(), ntm)
        out["ntm.attention_feature"] = cf.attention_feature(fr, 3, h.attention, update_ratio=0.2)[0].numpy()
        out["ntm.attention_feature_r05"] = cf.attention_feature(fr, 2, h.attention, update_ratio=0.5)[0].numpy()

        # ---- spatial compression
        for cs in (1, 2, 3, 6):
            out[f"spatial36.c{cs}"] = h.compress_spatial_features(t(inp["spatial36"]), cs).numpy()
        for cs in (1, 4, 9, 27):
            out[f"spatial729.c{cs}"] = h.compress_spatial_features(t(inp["spatial729"]), cs).numpy()

        # ---- whole legacy path for one video (729 x 1152 is hard-coded in the reference)
        big = mb.NeuralTuringMachine(input_dim=1152, output_dim=1152).eval()
        big.load_state_dict({k: t(v) for k, v in ntm_params(700, 1152).items()})
        for sample_type, seed in (("weighted_kmeans", 800), ("merge", 801)):
            hb = Holder(types.SimpleNamespace(video_sample_type=sample_type), big)
            hb.memory_mlp.load_state_dict({k: t(v) for k, v in mlp_params(701, 1152).items()})
            torch.manual_seed(seed)
            random.seed(seed)
            with contextlib.redirect_stdout(io.StringIO()):
                res = hb.compress_temporal_features([t(inp["temporal"])], [0])[0]
            out[f"temporal.{sample_type}.shape"] = np.asarray(res.shape)
            out[f"temporal.{sample_type}.sample"] = res.numpy()[TEMPORAL_SAMPLE].copy()
            out[f"temporal.{sample_type}.frame_sums"] = res.double().sum(dim=(1, 2)).numpy()

    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(OUT, "legacy_memory.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(out), "arrays,", len(meta), "index lists")


if __name__ == "__main__":
    assert os.path.isdir(REF), REF
    main()
