"""oracle/legacy_memory_oracle.py against the reference executed in the build container
(tests/golden/legacy_memory.npz, tools/gen_golden_legacy.py).  CPU only."""
import json
import os
import random
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from gen_golden_legacy import TEMPORAL_SAMPLE, legacy_inputs, mlp_params, ntm_params  # noqa: E402
from oracle import legacy_memory_oracle as O  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "legacy_memory.npz"))
META = json.loads(bytes(G["meta"]).decode())
INP = legacy_inputs()


def close(a, b, tol=1e-5):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-30), np.abs(a - b).max()


def coins(seed, n):
    random.seed(seed)
    return [random.randint(0, 1) for _ in range(n)]


def test_depth_scores_are_bit_exact():
    assert np.array_equal(O.cal_depth_score(INP["sim_rand"]), G["depth_rand"])
    assert np.array_equal(O.cal_depth_score(INP["sim_rand"], left_only=True), G["left_depth_rand"])


@pytest.mark.parametrize("name", ["seg_feat", "seg_feat_long"])
def test_scene_segmentation(name):
    f = INP[name]
    b, d = O.segment(f, alpha=0.5)
    assert b == META[f"{name}.segment_a05"]
    close(d, G[f"{name}.depth"], 2e-5)
    assert O.segment(f, k=3)[0] == META[f"{name}.segment_k3"]
    assert O.adjusted_segment(f, alpha=0.5, min_distance=4, max_distance=8) == META[f"{name}.adjusted_4_8"]
    assert O.adjusted_segment(f) == META[f"{name}.adjusted_default"]
    assert O.adjusted_segment(f, k=5, min_distance=2, max_distance=30) == META[f"{name}.adjusted_k5"]
    assert O.segment_left(f, alpha=0.5) == META[f"{name}.segment_left"]


def test_uniform_segments():
    from mavlm_b200.modules import uniform_segment_variant
    for key, want in META["uniform_segment"].items():
        T, d = map(int, key.split("_"))
        assert O.uniform_segment(T, d) == want, key
    for key, want in META["uniform_segment_variant"].items():
        T, d = map(int, key.split("_"))
        assert uniform_segment_variant(T, d) == want, key


@pytest.mark.parametrize("name,n", [("scene_video", 8), ("scene_video", 20), ("scene_video", 45), ("scene_video", 50),
                                    ("scene_video_busy", 4), ("scene_video_busy", 12)])
def test_scene_priority_sampling(name, n):
    torch.manual_seed(100 + n)
    got = O.sample_scenes_priority(INP[name], sample_num=n, randperm=lambda m: torch.randperm(m).tolist())
    assert got == META[f"{name}.sample_{n}"]


@pytest.mark.parametrize("T0", [3, 5])
def test_streaming_compression(T0):
    x = INP["stream"]
    T = x.shape[0]
    f, s, st = O.drop_feature(x, T0, coins(200 + T0, T - T0))
    assert np.array_equal(f, G[f"drop{T0}.feat"]) and st == META[f"drop{T0}.steps"]
    close(s, G[f"drop{T0}.sim"], 1e-5)
    f, s, st = O.merge_feature(x, T0)
    assert np.array_equal(f, G[f"merge{T0}.feat"]) and st == META[f"merge{T0}.steps"]
    close(s, G[f"merge{T0}.sim"], 1e-5)
    f, _, st = O.k_drop_feature(x, T0, coins(300 + T0, T - T0))
    assert np.array_equal(f, G[f"kdrop{T0}.feat"]) and st == META[f"kdrop{T0}.steps"]
    f, s, st = O.k_merge_feature(x, T0)
    assert np.array_equal(f, G[f"kmerge{T0}.feat"]) and st == META[f"kmerge{T0}.steps"]
    close(s, G[f"kmerge{T0}.sim"], 1e-5)


@pytest.mark.parametrize("T0", [3, 5])
def test_kmeans(T0):
    x = INP["stream"]
    T = x.shape[0]
    torch.manual_seed(400 + T0)
    random.seed(400 + T0)
    f, _, st = O.kmeans_feature(x, T0, torch.randperm(T)[:T0].tolist(), random.randint)
    close(f, G[f"kmeans{T0}.feat"], 1e-5)
    assert st == META[f"kmeans{T0}.steps"]
    torch.manual_seed(500 + T0)
    random.seed(500 + T0)
    w = torch.linspace(0.5, 2.0, T).numpy()
    f, ws, st = O.kmeans_feature(x, T0, torch.randperm(T)[:T0].tolist(), random.randint, weights=w, weighted=True)
    close(f, G[f"wkmeans{T0}.feat"], 1e-5)
    close(ws, G[f"wkmeans{T0}.weights"], 1e-5)
    assert st == META[f"wkmeans{T0}.steps"]


def test_identity_when_short():
    f, _, st = O.merge_feature(INP["stream"][:3], 5)
    assert np.array_equal(f, G["merge_identity.feat"]) and st == META["merge_identity.steps"]


def test_turing_memory():
    p = ntm_params(600, 16)
    fr = INP["ntm_frames"]
    a, b = fr[:3].reshape(-1, 16), fr[3:5].reshape(-1, 16)
    qk = (p["q_proj.weight"], p["q_proj.bias"], p["k_proj.weight"], p["k_proj.bias"])
    close(O.ntm_get_weight(a, b, *qk), G["ntm.weight"], 1e-5)
    close(O.ntm_forward(a, b, p), G["ntm.forward"], 1e-5)
    close(O.attention_feature(fr, 3, *qk, update_ratio=0.2), G["ntm.attention_feature"], 1e-5)
    close(O.attention_feature(fr, 2, *qk, update_ratio=0.5), G["ntm.attention_feature_r05"], 1e-5)


def test_spatial_compression():
    for cs in (1, 2, 3, 6):
        close(O.compress_spatial_features(INP["spatial36"], cs), G[f"spatial36.c{cs}"], 1e-6)
    for cs in (1, 4, 9, 27):
        close(O.compress_spatial_features(INP["spatial729"], cs), G[f"spatial729.c{cs}"], 1e-6)


@pytest.mark.parametrize("sample_type,seed", [("weighted_kmeans", 800), ("merge", 801)])
def test_whole_legacy_memory(sample_type, seed):
    x = INP["temporal"]
    torch.manual_seed(seed)
    random.seed(seed)
    init = torch.randperm(x.shape[0] - 1)[:3].tolist() if sample_type == "weighted_kmeans" else ()
    res = O.compress_temporal_features(x, sample_type=sample_type, ntm=ntm_params(700, 1152), mlp=mlp_params(701, 1152),
                                       init=init, randint=random.randint)
    assert list(res.shape) == list(G[f"temporal.{sample_type}.shape"])
    close(res[TEMPORAL_SAMPLE], G[f"temporal.{sample_type}.sample"], 2e-5)
    close(res.sum(axis=(1, 2)), G[f"temporal.{sample_type}.frame_sums"], 2e-4)


def test_host_side_of_the_drop_in():
    """What runs without a GPU: boundary lists, step replay, and the refusal to compute on CPU tensors."""
    from mavlm_b200 import legacy as L
    for key, want in META["uniform_segment"].items():
        T, d = map(int, key.split("_"))
        assert L.uniform_segment(T, d) == want, key
    # decisions (idx, idx+1) of a merge run rebuild the reference's step lists
    want = META["merge3.steps"]
    dec = []
    for before, after in zip(want[:-1], want[1:]):
        allg = before + [[max(max(g) for g in before) + 1]]
        idx = next(i for i in range(len(after)) if after[i] != allg[i])
        dec.append((idx, idx + 1))
    assert L._replay_steps(L.MERGE, INP["stream"].shape[0], 3, dec, None) == want
    # the other three modes: recover each step's decision from the reference's lists by search, then replay the log
    T = INP["stream"].shape[0]

    def apply(mode, groups, i, a, b):
        allg = [list(g) for g in groups] + [[i]]
        if mode in (L.DROP, L.K_DROP):
            del allg[a]
        else:
            allg[b] = allg[a] + allg[b]
            del allg[a]
        return allg

    for mode, key in ((L.DROP, "drop5"), (L.K_DROP, "kdrop5"), (L.K_MERGE, "kmerge5")):
        want = META[f"{key}.steps"]
        dec = []
        for n, (before, after) in enumerate(zip(want[:-1], want[1:])):
            found = [(a, b) for a in range(6) for b in range(6)
                     if (mode in (L.DROP, L.K_DROP) or a != b) and apply(mode, before, 5 + n, a, b) == after]
            assert found, (key, n)
            dec.append(found[0])
        coins_ = [1] * len(dec)                                  # k_drop with coin 1 drops `left` = the index we recovered
        assert L._replay_steps(mode, T, 5, dec, coins_) == want, key
    with pytest.raises(RuntimeError):
        L.segment(torch.from_numpy(INP["seg_feat"]))
    with pytest.raises(RuntimeError):
        L.merge_feature(torch.from_numpy(INP["stream"]), 3)
    f, s_, st = L.merge_feature(torch.from_numpy(INP["stream"][:2]), 3)          # short video: passes through untouched
    assert f.shape[0] == 2 and s_ is None and st == [[[0], [1]]]
