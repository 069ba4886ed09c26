"""SURVEY.md §8f-4 on the GPU: scene segmentation, streaming compression, k-means, Turing memory and the whole
legacy memory builder, through the C ABI, against oracle/legacy_memory_oracle.py and the reference's own outputs
(tests/golden/legacy_memory.npz).  Index / decision outputs must be identical; copies bit-exact; fp32 arithmetic
<= 1e-5 (GEMM-backed pieces 1e-4), bf16 <= 2e-2 of the tensor maximum."""
import json
import os
import random
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from gen_golden_legacy import TEMPORAL_SAMPLE, legacy_inputs, mlp_params, ntm_params, scene_frames  # noqa: E402
from oracle import legacy_memory_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(ROOT, "tests", "golden", "legacy_memory.npz"))
META = json.loads(bytes(G["meta"]).decode())
INP = legacy_inputs()


@pytest.fixture(scope="module")
def L():
    import mavlm_b200
    from mavlm_b200 import legacy
    return legacy


def cu(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda().to(dtype)


def err(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def coins(seed, n):
    random.seed(seed)
    return [random.randint(0, 1) for _ in range(n)]


def test_depth_scores_bit_exact(L):
    s = cu(INP["sim_rand"])
    assert np.array_equal(L.cal_depth_score(s).cpu().numpy(), G["depth_rand"])
    assert np.array_equal(L.cal_left_depth_score(s).cpu().numpy(), G["left_depth_rand"])
    g = np.random.default_rng(5)
    big = np.round(g.uniform(0, 1, 3000), 2).astype(np.float32)          # many ties and long plateaus
    assert np.array_equal(L.cal_depth_score(cu(big)).cpu().numpy(), O.cal_depth_score(big))
    assert np.array_equal(L.cal_left_depth_score(cu(big)).cpu().numpy(), O.cal_depth_score(big, left_only=True))


@pytest.mark.parametrize("name", ["seg_feat", "seg_feat_long"])
def test_scene_segmentation_matches_the_reference(L, name):
    f = cu(INP[name])
    b, d = L.segment(f, alpha=0.5)
    assert b == META[f"{name}.segment_a05"]
    assert err(d, G[f"{name}.depth"]) <= 2e-5
    assert L.segment(f, k=3)[0] == META[f"{name}.segment_k3"]
    assert L.adjusted_segment(f, alpha=0.5, min_distance=4, max_distance=8) == META[f"{name}.adjusted_4_8"]
    assert L.adjusted_segment(f) == META[f"{name}.adjusted_default"]
    assert L.adjusted_segment(f, k=5, min_distance=2, max_distance=30) == META[f"{name}.adjusted_k5"]
    assert L.segment_left(f, alpha=0.5) == META[f"{name}.segment_left"]


def test_uniform_segment_and_edge_cases(L):
    for key, want in META["uniform_segment"].items():
        T, d = map(int, key.split("_"))
        assert L.uniform_segment(T, d) == want
    one = cu(INP["seg_feat"][:1])
    assert L.segment(one)[0] == [0] and L.adjusted_segment(one) == [0]
    with pytest.raises(IndexError):
        L.segment(cu(INP["seg_feat"][:2]))                                # segment.py:31 on two frames
    with pytest.raises(RuntimeError):
        L.segment(torch.from_numpy(INP["seg_feat"]))                      # CPU tensor: no fallback


@pytest.mark.parametrize("name,n", [("scene_video", 8), ("scene_video", 20), ("scene_video", 45), ("scene_video", 50),
                                    ("scene_video_busy", 4), ("scene_video_busy", 12)])
def test_scene_priority_sampling(L, name, n):
    torch.manual_seed(100 + n)
    assert L.sample_scenes_priority(cu(INP[name]), sample_num=n) == META[f"{name}.sample_{n}"]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
def test_adjacent_cosine_and_frame_means_at_tower_size(L, dtype, tol):
    x = scene_frames(31, 6, 729, 1152, scenes=2)
    xt = cu(x, dtype)
    xr = xt.float().cpu().numpy()
    flat = xr.reshape(6, -1)
    assert err(L.adjacent_cosine(xt), O.cosine_rows(flat[:-1], flat[1:])) <= tol
    assert err(L.frame_means(xt), xr.astype(np.float64).mean(1)) <= tol


@pytest.mark.parametrize("T0", [3, 5])
def test_streaming_compression_matches_the_reference(L, T0):
    x = cu(INP["stream"])
    T = x.shape[0]
    random.seed(200 + T0)
    f, s, st = L.drop_feature(x, T0)
    assert np.array_equal(f.cpu().numpy(), G[f"drop{T0}.feat"]) and st == META[f"drop{T0}.steps"]
    assert err(s, G[f"drop{T0}.sim"]) <= 1e-5
    f, s, st = L.merge_feature(x, T0)
    assert np.array_equal(f.cpu().numpy(), G[f"merge{T0}.feat"]) and st == META[f"merge{T0}.steps"]
    assert err(s, G[f"merge{T0}.sim"]) <= 1e-5
    random.seed(300 + T0)
    f, s, st = L.k_drop_feature(x, T0)
    assert s is None and np.array_equal(f.cpu().numpy(), G[f"kdrop{T0}.feat"]) and st == META[f"kdrop{T0}.steps"]
    f, s, st = L.k_merge_feature(x, T0)
    assert np.array_equal(f.cpu().numpy(), G[f"kmerge{T0}.feat"]) and st == META[f"kmerge{T0}.steps"]
    assert err(s, G[f"kmerge{T0}.sim"]) <= 1e-5


def test_streaming_compression_short_video_passes_through(L):
    x = cu(INP["stream"][:3])
    f, s, st = L.merge_feature(x, 5)
    assert f is x and s is None and st == META["merge_identity.steps"]
    with pytest.raises(RuntimeError):
        L.stream_compress(cu(INP["stream"]), 65, L.MERGE)                  # keep > 64 / not enough frames
    with pytest.raises(RuntimeError):
        L.stream_compress(cu(scene_frames(3, 80, 2, 8)), 65, L.MERGE)      # keep > 64 with enough frames: C ABI refuses
    with pytest.raises(ValueError):
        L.stream_compress(cu(INP["stream"]), 3, L.DROP, [1, 0])            # too few coins


@pytest.mark.parametrize("mode", ["drop", "merge", "kdrop", "kmerge"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_streaming_compression_at_tower_size(L, mode, dtype):
    """40 frames of 729 x 1152 kept to 4 (64 partial-sum CTAs per pair, scratch rows, free-slot recycling)."""
    T, T0 = 40, 4
    x = cu(scene_frames(32, T, 729, 1152, scenes=7), dtype)
    xr = x.float().cpu().numpy()
    c = coins(77, T - T0)
    random.seed(77)
    if mode == "drop":
        got, want = L.drop_feature(x, T0), O.drop_feature(xr, T0, c)
    elif mode == "merge":
        got, want = L.merge_feature(x, T0), O.merge_feature(xr, T0)
    elif mode == "kdrop":
        got, want = L.k_drop_feature(x, T0), O.k_drop_feature(xr, T0, c)
    else:
        got, want = L.k_merge_feature(x, T0), O.k_merge_feature(xr, T0)
    assert got[2] == want[2], "streaming decisions differ"
    if mode in ("drop", "kdrop"):
        assert np.array_equal(got[0].float().cpu().numpy(), want[0])
    else:
        assert err(got[0], want[0]) <= (1e-6 if dtype == torch.float32 else 1e-2)
    if want[1] is not None:
        assert err(got[1], want[1]) <= (1e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("mode", ["drop", "merge", "kdrop", "kmerge"])
@pytest.mark.parametrize("T,T0", [(9, 1), (12, 2), (6, 5), (90, 64), (70, 33)])
def test_streaming_compression_edge_sizes(L, mode, T, T0):
    """keep = 1 (no similarities at all), keep = 2, one streamed frame only, the cap of 64 kept frames (2016 initial
    pairs, 129 row pairs per launch in k_merge) and an odd keep, against the oracle on the same fp32 frames."""
    x = scene_frames(50 + T0, T, 4, 16, scenes=min(6, T))
    c = coins(60 + T0, T - T0)
    xt = cu(x)
    if mode == "drop":
        got, want = L.stream_compress(xt, T0, L.DROP, c), O.drop_feature(x, T0, c)
    elif mode == "merge":
        got, want = L.stream_compress(xt, T0, L.MERGE), O.merge_feature(x, T0)
    elif mode == "kdrop":
        got, want = L.stream_compress(xt, T0, L.K_DROP, c), O.k_drop_feature(x, T0, c)
    else:
        got, want = L.stream_compress(xt, T0, L.K_MERGE), O.k_merge_feature(x, T0)
    assert got[2] == want[2], "streaming decisions differ"
    assert err(got[0], want[0]) <= 1e-6
    if mode in ("drop", "merge") and T0 > 1:
        assert err(got[1][:T0 - 1], want[1]) <= 1e-5
    if mode == "kmerge":
        assert err(got[1].view(T0, T0), want[1]) <= 1e-5


def test_streaming_compression_fp16(L):
    x = cu(scene_frames(70, 20, 8, 64, scenes=4), torch.float16)
    xr = x.float().cpu().numpy()
    f, s, st = L.merge_feature(x, 3)
    wf, ws, wst = O.merge_feature(xr, 3)
    assert st == wst and err(f, wf) <= 2e-3 and err(s, ws) <= 2e-3
    c = coins(71, 17)
    random.seed(71)
    f, _, st = L.k_drop_feature(x, 3)
    wf, _, wst = O.k_drop_feature(xr, 3, c)
    assert st == wst and np.array_equal(f.float().cpu().numpy(), wf)


@pytest.mark.parametrize("mode", ["drop", "merge", "kdrop", "kmerge"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_batched_streaming_equals_per_video_calls(L, mode, dtype):
    """Ragged batch (different lengths) in one launch sequence: same decisions and frames as one call per video."""
    m = {"drop": L.DROP, "merge": L.MERGE, "kdrop": L.K_DROP, "kmerge": L.K_MERGE}[mode]
    lens, T0 = (23, 9, 40, 6), 4
    vids = [cu(scene_frames(40 + i, t, 196, 512, scenes=4), dtype) for i, t in enumerate(lens)]
    cs = [coins(90 + i, t - T0) for i, t in enumerate(lens)]
    batched = L.stream_compress_batched(vids, T0, m, cs)
    for v, c, (bf, bs, bst) in zip(vids, cs, batched):
        f, s, st = L.stream_compress(v, T0, m, c)
        assert torch.equal(bf, f) and st == bst
        if m != L.K_DROP:
            n = T0 - 1 if m in (L.DROP, L.MERGE) else T0 * T0
            assert err(bs[:n], s[:n].cpu().numpy()) <= 1e-6            # partial-sum count per row pair depends on the batch
    with pytest.raises(RuntimeError):
        L.stream_compress_batched([vids[0], vids[1][:, :, :256].contiguous()], T0, m, cs[:2])


@pytest.mark.parametrize("T0", [3, 5])
def test_kmeans_matches_the_reference(L, T0):
    x = cu(INP["stream"])
    T = x.shape[0]
    torch.manual_seed(400 + T0)
    random.seed(400 + T0)
    f, s, st = L.kmeans_feature(x, T0)
    assert s is None and st == META[f"kmeans{T0}.steps"]
    assert err(f, G[f"kmeans{T0}.feat"]) <= 1e-5
    # weighted: the reference draws its start on x.device; replay the CPU draw of the golden run
    torch.manual_seed(500 + T0)
    random.seed(500 + T0)
    init = torch.randperm(T)[:T0]
    w = torch.linspace(0.5, 2.0, T).cuda()
    cent, labels, wsum, _ = L._kmeans(x.reshape(T, -1), T0, w, init)
    lab = labels.cpu().tolist()
    assert [[[j for j in range(T) if lab[j] == i] for i in range(T0)]] == META[f"wkmeans{T0}.steps"]
    assert err(cent.view(T0, *x.shape[1:]), G[f"wkmeans{T0}.feat"]) <= 1e-5
    assert err(wsum, G[f"wkmeans{T0}.weights"]) <= 1e-5
    f, s, st = L.weighted_kmeans_feature(x, T0, w)                         # public entry: device-side randperm
    assert f.shape == (T0,) + tuple(x.shape[1:]) and s.shape == (T0,) and sorted(sum(st[0], [])) == list(range(T))


def test_kmeans_reseeds_an_empty_cluster(L):
    x = cu(INP["stream"])
    T = x.shape[0]
    xr = INP["stream"]
    init = [0, 0, 5]                                                       # duplicate start -> cluster 1 is empty
    random.seed(9)
    cent, labels, wsum, _ = L._kmeans(x.reshape(T, -1), 3, None, torch.tensor(init))
    random.seed(9)
    want, _, steps = O.kmeans_feature(xr, 3, init, random.randint)
    lab = labels.cpu().tolist()
    assert [[[j for j in range(T) if lab[j] == i] for i in range(3)]] == steps
    assert err(cent.view(3, *x.shape[1:]), want) <= 1e-5


def _ntm(L, d, seed, dtype):
    m = L.NeuralTuringMachine(input_dim=d, output_dim=d).eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ntm_params(seed, d).items()})
    return m.cuda().to(dtype)


def test_turing_memory_matches_the_reference(L):
    ntm = _ntm(L, 16, 600, torch.float32)
    fr = cu(INP["ntm_frames"])
    a, b = fr[:3].reshape(-1, 16), fr[3:5].reshape(-1, 16)
    with torch.no_grad():
        assert err(ntm.get_weight(a, b), G["ntm.weight"]) <= 1e-5
        assert err(ntm(a, b), G["ntm.forward"]) <= 1e-5
        fn = lambda m, n, update_ratio: ntm.gated_update(m, n, update_ratio)   # noqa: E731
        assert err(L.attention_feature(fr, 3, fn, update_ratio=0.2)[0], G["ntm.attention_feature"]) <= 1e-5
        assert err(L.attention_feature(fr, 2, fn, update_ratio=0.5)[0], G["ntm.attention_feature_r05"]) <= 1e-5
        ntm.train()
        with pytest.raises(NotImplementedError):
            ntm(a, b)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2), (torch.float16, 5e-3)])
def test_turing_memory_at_tower_size(L, dtype, tol):
    """3 x 729 memory tokens folded over 5 more frames of 1152-d tokens (ragged last group)."""
    d = 1152
    p = ntm_params(700, d)
    ntm = _ntm(L, d, 700, dtype)
    x = cu(0.3 * scene_frames(33, 8, 729, d, scenes=3), dtype)
    xr = x.float().cpu().numpy()
    pr = {k: torch.from_numpy(v).to(dtype).float().numpy() for k, v in p.items()}
    want = O.attention_feature(xr, 3, pr["q_proj.weight"], pr["q_proj.bias"], pr["k_proj.weight"], pr["k_proj.bias"], 0.2)
    with torch.no_grad():
        got, _ = L.attention_feature(x, 3, lambda m, n, update_ratio: ntm.gated_update(m, n, update_ratio), 0.2)
    assert err(got, want) <= tol


def test_spatial_compression_matches_the_reference(L):
    h = L.MultimodalOpsMixin(types.SimpleNamespace())
    for cs in (1, 2, 3, 6):
        assert err(h.compress_spatial_features(cu(INP["spatial36"]), cs), G[f"spatial36.c{cs}"]) <= 1e-6
    for cs in (1, 4, 9):
        assert err(h.compress_spatial_features(cu(INP["spatial729"]), cs), G[f"spatial729.c{cs}"]) <= 1e-6
    x = cu(INP["spatial729"])
    assert h.compress_spatial_features(x, 27) is x
    with pytest.raises(AssertionError):
        h.compress_spatial_features(cu(INP["spatial36"][:, :35]), 2)
    bad = L.MultimodalOpsMixin(types.SimpleNamespace(compress_type="max"))
    with pytest.raises(NotImplementedError):
        bad.compress_spatial_features(cu(INP["spatial36"]), 2)


class _Holder:
    def get_model(self):
        return self


@pytest.mark.parametrize("sample_type,seed", [("weighted_kmeans", 800), ("merge", 801)])
def test_whole_legacy_memory_matches_the_reference(L, sample_type, seed):
    class H(L.MultimodalOpsMixin, _Holder):
        pass

    h = H(types.SimpleNamespace(video_sample_type=sample_type))
    h.attention_model = _ntm(L, 1152, 700, torch.float32)
    h.memory_mlp.load_state_dict({k: torch.from_numpy(v) for k, v in mlp_params(701, 1152).items()})
    h.memory_mlp.cuda()
    x = cu(INP["temporal"])
    torch.manual_seed(seed)
    random.seed(seed)
    if sample_type == "weighted_kmeans":                                    # replay the reference's CPU randperm
        init = torch.randperm(x.shape[0] - 1)[:3]
        orig = torch.randperm
        torch.randperm = lambda n, device=None: torch.tensor(init.tolist() + [i for i in range(n) if i not in init.tolist()])
    try:
        with torch.no_grad():
            res = h.compress_temporal_features([x, x], [0])
    finally:
        if sample_type == "weighted_kmeans":
            torch.randperm = orig
    assert res[1] is None
    res = res[0]
    assert list(res.shape) == list(G[f"temporal.{sample_type}.shape"])
    assert err(res.cpu().numpy()[TEMPORAL_SAMPLE], G[f"temporal.{sample_type}.sample"]) <= 1e-4
    assert err(res.double().sum(dim=(1, 2)), G[f"temporal.{sample_type}.frame_sums"]) <= 1e-3


def test_unknown_sample_type_raises(L):
    class H(L.MultimodalOpsMixin, _Holder):
        pass

    with pytest.raises(NotImplementedError):
        H(types.SimpleNamespace(video_sample_type="bogus")).compress_temporal_features([], [])
