"""Pin oracle/vismem_oracle.py against fixtures produced by the UNMODIFIED reference
(tools/gen_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import vismem_oracle as O

from conftest import GOLDEN


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    w = {k[3:]: z[k] for k in z.files if k.startswith("w::")}
    return z, w


def err(a, b):
    return O.normalized_max_error(a, b)


def test_rmt_three_chunks_formation_and_evolution():
    z, w = _load("rmt_small.npz")
    frames = z["frames"]
    cache, scores = [], []
    for i in range(3):
        cache, s = O.rmt_chunk(frames[2 * i:2 * i + 2], cache, w, want_scores=True)
        scores.append(s)
    assert len(cache) == 3
    for i in range(3):
        assert err(cache[i], z[f"state{i}"]) < 2e-5, i
        assert err(scores[i], z[f"score{i}"]) < 2e-5, i
        # frame scores sum to H*Lq/P = 64
        assert abs(scores[i].sum() - 64.0) < 1e-3


def test_rmt_float64_is_closer_than_float32_noise():
    z, w = _load("rmt_small.npz")
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    cache, _ = O.rmt_video(z["frames"].astype(np.float64), w64, chunk=2)
    assert err(cache[-1], z["state2"]) < 2e-5
    assert err(cache[-1], z["f64_state_last"]) < 1e-10


def test_rmt_stress_sharp_softmax():
    z, w = _load("rmt_small.npz")
    ws = {k: (v * 8.0 if "q_proj" in k else v) for k, v in w.items()}
    cache, _ = O.rmt_video(4.0 * z["frames"], ws, chunk=2)
    assert err(cache[0], z["stress_state_first"]) < 5e-5
    assert err(cache[-1], z["stress_state_last"]) < 1e-3          # fp32-vs-fp32 noise, amplified by the sharp softmax
    w64 = {k: v.astype(np.float64) for k, v in ws.items()}
    cache64, _ = O.rmt_video(4.0 * z["frames"].astype(np.float64), w64, chunk=2)
    assert err(cache64[-1], z["f64_stress_state_last"]) < 1e-10   # tight pin: float64 reference vs float64 oracle


def test_rmt_cache_cap_and_single_frame_chunks():
    z, w = _load("rmt_small.npz")
    cache, _ = O.rmt_video(z["frames12"], w, chunk=1)
    assert len(cache) == int(z["cap_len"]) == 10
    assert err(cache[0], z["cap_state_first"]) < 5e-5
    assert err(cache[-1], z["cap_state_last"]) < 5e-5
    sums = np.array([s.astype(np.float64).sum() for s in cache])
    assert np.allclose(sums, z["cap_state_sums"], atol=0.05)


def test_pool_modes():
    z, _ = _load("pool.npz")
    x = z["x"]
    assert err(O.get_2d_pool(x, 2, "bilinear"), z["bilinear"]) < 4e-6      # fp32 tap-weight rounding; fp64 below is tight
    assert err(O.get_2d_pool(x.astype(np.float64), 2, "bilinear"), z["bilinear_f64"]) < 1e-13
    assert err(O.get_2d_pool(x, 2, "average"), z["average"]) < 1e-6
    assert np.array_equal(O.get_2d_pool(x, 2, "max"), z["max"])
    assert O.get_2d_pool(x, 2, "average").shape == (2, 169, 8)
    assert err(O.get_2d_pool(x, 3, "bilinear"), z["bilinear_s3"]) < 4e-6
    with pytest.raises(ValueError):
        O.get_2d_pool(x, 2, "nearest")


def test_bilinear_taps_first_row():
    i0, i1, lam = O.bilinear_taps(27, 14)
    assert i0[0] == 0 and i1[0] == 1 and abs(lam[0] - 13 / 28) < 1e-12      # .5357/.4643
    assert i1[-1] == 26 and i0[-1] == 25


def test_temporal_pe():
    z, _ = _load("pe.npz")
    t32 = O.temporal_pe_table(600, 32)
    assert np.abs(t32 - z["table32"]).max() < 5e-5   # 1-ulp exp() differences x position 599 (fp32 table)
    for d in (32, 896, 3584):
        t = O.temporal_pe_table(600, d)
        assert np.abs(t[[0, 1, 7, 131, 599]] - z[f"rows{d}"]).max() < 2e-4, d   # fp32 angle 599*div: 1 ulp of div = 4e-5 rad
        assert abs(t.astype(np.float64).sum() - float(z[f"sum{d}"])) < 0.5, d
    y = O.add_temporal_pe(z["x"], z["idx"], z["table32"])
    assert np.array_equal(y, z["y"])
    assert np.array_equal(O.add_temporal_pe(z["x"], None, z["table32"]), z["y_default_idx"])
    with pytest.raises(ValueError):
        O.add_temporal_pe(z["x"], np.array([0, 1, 2, 3, 600]), z["table32"])
    with pytest.raises(ValueError):
        O.add_temporal_pe(z["x"], np.array([0, 1, 2, 3, -1]), z["table32"])
    with pytest.raises(ValueError):
        O.add_temporal_pe(z["x"][0], None, z["table32"])


def test_index_generation_matches_torch_linspace():
    with open(os.path.join(GOLDEN, "indices.json")) as fh:
        g = json.load(fh)
    for f, idx in g["sample"].items():
        assert O.sample_frame_indices(int(f)).tolist() == idx, f
    for n, idx in g["fine"].items():
        assert O.fine_frame_indices(int(n)).tolist() == idx, n
    for key, b in g["bounds"].items():
        t, d = map(int, key.split(","))
        assert O.uniform_segment_variant(t, d) == b, key


def test_projector_and_fusers():
    z, w = _load("projector_fuser.npz")
    assert err(O.mm_projector(z["proj_x"], w), z["proj_y"]) < 2e-6
    assert err(O.memory_fuser_mlp(z["fuser_x"], w), z["fuser_y"]) < 2e-6
    for nl in (1, 2):
        y = O.memory_fuser_encoder(z[f"enc{nl}_x"], w, prefix=f"enc{nl}.", num_layers=nl, heads=4)
        assert err(y, z[f"enc{nl}_y"]) < 1e-5, nl


def _embed(w, ids):
    tab = w["_emb.weight"]
    return tab[np.asarray(ids) % tab.shape[0]]


@pytest.mark.parametrize("name,raw", [("full_path.npz", 70), ("full_path_short.npz", 5)])
def test_full_path_against_reference_prepare_inputs(name, raw):
    zf, w = _load("full_path.npz")
    z = np.load(os.path.join(GOLDEN, name))
    video = z["video"]
    idx = O.sample_frame_indices(raw)
    tower = video[idx].reshape(len(idx), video.shape[1], 729).transpose(0, 2, 1)
    res = O.visual_memory_path(tower, idx, w, pe_table=w["positional_encoding.frame_embed"],
                               prompt_mem=_embed(w, O.MEMORY_PROMPT_IDS), prompt_frm=_embed(w, O.FRAME_PROMPT_IDS))
    ids = zf["input_ids"][0]
    seq = np.concatenate([_embed(w, ids[:2]), res["sequence"], _embed(w, ids[3:])], axis=0)
    ref = z["inputs_embeds"][0]
    assert seq.shape == ref.shape
    assert err(seq, ref) < 2e-5


# ---- the differentiable torch oracle (oracle/vismem_torch_oracle.py) against the reference's own autograd ----
def _load_grads(name):
    z = np.load(os.path.join(GOLDEN, name))
    w = {k[3:]: z[k] for k in z.files if k.startswith("w::")}
    g = {k[3:]: z[k] for k in z.files if k.startswith("g::")}
    return z, w, g


def test_torch_oracle_gradients_rmt_bptt():
    """2 chunks (formation x2 + evolution), loss = sum of mean(state^2): every RMT parameter gradient equals the
    float64 autograd of the unmodified reference module (tools/gen_golden.py::gen_rmt_grads)."""
    import torch
    from oracle import vismem_torch_oracle as T
    z, w, g = _load_grads("rmt_grads.npz")
    wt = T.leaf_weights(w, torch.float64)
    frames = torch.from_numpy(z["frames"])
    cache = T.rmt_video(frames, wt, chunk=2)
    loss = sum((s * s).mean() for s in cache)
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-12 * abs(float(z["loss"])) + 1e-15
    loss.backward()
    gmax = max(np.abs(v).max() for v in g.values())
    assert len(g) == 44
    for k, ref in g.items():
        got = wt[k].grad.numpy()
        assert np.abs(got - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-6 * gmax), k


def test_torch_oracle_gradients_whole_path():
    """loss = mean(sequence^2) through 3 chunks + fuser + type embeddings + newline + prompt rows
    (tools/gen_golden.py::gen_path_grads); the forward also equals the numpy oracle's."""
    import torch
    from oracle import vismem_torch_oracle as T
    z, w, g = _load_grads("path_grads.npz")
    rows = z["embed_rows"].tolist()
    tab = {r: w["embed_rows"][i] for i, r in enumerate(rows)}
    pm = np.stack([tab[i] for i in T.MEMORY_PROMPT_IDS])
    pf = np.stack([tab[i] for i in T.FRAME_PROMPT_IDS])
    wn = {k: v for k, v in w.items() if k != "embed_rows"}
    loss, grads, seqs = T.path_gradients(z["z"][None], wn, chunk=2, dtype=torch.float64, prompt_rows=(pm, pf))
    assert abs(loss - float(z["loss"])) < 1e-12 * abs(float(z["loss"]))
    assert err(seqs[0], z["sequence"]) < 1e-12
    gmax = max(np.abs(v).max() for v in g.values())
    for k, ref in g.items():
        if k == "embed_rows":
            continue
        assert np.abs(grads[k] - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-6 * gmax), k
    # prompt-row gradients: a token id that occurs in both prompts (279, 2766, 25) accumulates both
    acc = {r: np.zeros_like(pm[0]) for r in rows}
    for i, r in enumerate(T.MEMORY_PROMPT_IDS):
        acc[r] += grads["embed.prompt_mem"][i]
    for i, r in enumerate(T.FRAME_PROMPT_IDS):
        acc[r] += grads["embed.prompt_frm"][i]
    got = np.stack([acc[r] for r in rows])
    assert np.abs(got - g["embed_rows"]).max() <= 1e-9 * np.abs(g["embed_rows"]).max()
    # numpy oracle forward on the same inputs
    fused_seq = O.assemble_sequence(
        O.memory_fuser_mlp(np.concatenate(O.rmt_video(z["z"], wn, chunk=2)[0], axis=0), wn),
        z["z"][O.fine_frame_indices(6)], wn, prompt_mem=pm, prompt_frm=pf)
    assert err(seqs[0], fused_seq) < 1e-12
