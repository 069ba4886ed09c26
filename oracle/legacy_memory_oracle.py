"""CPU oracle (test infrastructure only) for SURVEY.md §8f-4: the Flash-VStream-style memories and the
scene-based segmentation that sit beside the recurrent memory in the reference tree.

Reference (never imported by the product path; restated here in numpy on index lists instead of tensor slicing):
  llava/model/memory_module/segment.py:3-25      cal_depth_score        (:210-223 left-only variant)
  llava/model/memory_module/segment.py:27-49     segment
  llava/model/memory_module/segment.py:52-128    adjusted_segment
  llava/model/memory_module/segment.py:131-167   uniform_segment
  llava/model/memory_module/segment.py:226-250   segment_left
  llava/model/memory_module/segment.py:252-337   sample_scenes_priority
  llava/model/memory_module/compress_functions.py:20-56    drop_feature
  llava/model/memory_module/compress_functions.py:59-91    merge_feature
  llava/model/memory_module/compress_functions.py:94-131   kmeans_feature
  llava/model/memory_module/compress_functions.py:134-173  weighted_kmeans_feature
  llava/model/memory_module/compress_functions.py:176-218  k_drop_feature
  llava/model/memory_module/compress_functions.py:221-264  k_merge_feature
  llava/model/memory_module/compress_functions.py:267-280  attention_feature
  llava/model/memory_module/memory_builder.py:8-39         NeuralTuringMachine (get_weight / forward)
  llava/model/memory_module/memory_builder.py:52-64        gated attention update of the Turing memory
  llava/model/memory_module/memory_builder.py:72-99        compress_spatial_features
  llava/model/memory_module/memory_builder.py:101-190      compress_temporal_features

The arithmetic is ATen's (torch.cosine_similarity: dot / (max(|a|, eps) * max(|b|, eps)); F.normalize eps 1e-12;
torch.std_mean unbiased; argmax = first maximum).  The reference draws its random decisions from Python's `random`
(one `randint(0, 1)` per streamed frame in drop / k_drop; `randint(0, T-1)` per empty k-means cluster) and from
`torch.randperm` (k-means initialisation, scene-sample padding); here they are explicit arguments so that a seeded
run of the reference can be replayed.  Pinned in tests/test_legacy_oracle.py against the reference executed in the
build container (tests/golden/legacy_memory.npz, made by tools/gen_golden_legacy.py).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

NEG = -100.0


# ------------------------------------------------------------------------------------------------ similarities
def cosine_rows(a: np.ndarray, b: np.ndarray, eps: float = 1e-8) -> np.ndarray:
    """torch.cosine_similarity(a, b, dim=-1, eps) on 2-D inputs, float64 accumulation, float32 result."""
    a64 = a.astype(np.float64)
    b64 = b.astype(np.float64)
    na = np.maximum(np.sqrt((a64 * a64).sum(-1)), eps)
    nb = np.maximum(np.sqrt((b64 * b64).sum(-1)), eps)
    return ((a64 * b64).sum(-1) / (na * nb)).astype(np.float32)


def _cos1(a: np.ndarray, b: np.ndarray, eps: float = 1e-8) -> np.float32:
    return cosine_rows(a.reshape(1, -1), b.reshape(1, -1), eps)[0]


# ------------------------------------------------------------------------------------------------ segmentation
def cal_depth_score(sim: np.ndarray, left_only: bool = False) -> np.ndarray:
    """segment.py:3-25 (:210-223 when left_only): climb to the nearest peak on each side while the similarity does
    not decrease; depth = lpeak + rpeak - 2 * sim (float32 arithmetic in that order)."""
    sim = sim.astype(np.float32)
    n = sim.shape[0]
    out = np.zeros(n, dtype=np.float32)
    for i in range(n):
        lpeak = sim[i]
        j = i - 1
        while j >= 0 and sim[j] >= lpeak:
            lpeak = sim[j]
            j -= 1
        if left_only:
            out[i] = np.float32(lpeak - sim[i])
            continue
        rpeak = sim[i]
        j = i + 1
        while j < n and sim[j] >= rpeak:
            rpeak = sim[j]
            j += 1
        out[i] = np.float32(np.float32(lpeak + rpeak) - np.float32(np.float32(2.0) * sim[i]))
    return out


def _threshold_boundaries(depth: np.ndarray, alpha: float, k: Optional[int], cap: Optional[int] = None) -> List[int]:
    if k is not None:
        order = np.argsort(-depth, kind="stable")[:k]
        return sorted(int(i) for i in order)
    mean = depth.astype(np.float64).mean()
    std = depth.astype(np.float64).std(ddof=1) if depth.shape[0] > 1 else float("nan")
    thresh = np.float32(np.float32(mean) + np.float32(alpha) * np.float32(std))
    picked = [int(i) for i in np.nonzero(depth > thresh)[0]]
    if cap is not None and len(picked) > cap:
        order = np.argsort(-depth, kind="stable")[:cap]
        picked = sorted(int(i) for i in order)
    return picked


def segment(features: np.ndarray, alpha: float = 0.5, k: Optional[int] = None) -> Tuple[List[int], np.ndarray]:
    """segment.py:27-49.  features [T, D] -> (sorted boundary list, depth scores [T-1])."""
    T = features.shape[0]
    if T == 1:
        return [0], np.zeros(1, dtype=np.float32)
    sim = cosine_rows(features[:-1], features[1:], eps=1e-2)
    sim[0] = sim[1]
    depth = cal_depth_score(sim)
    b = _threshold_boundaries(depth, alpha, k)
    if not b or b[-1] != T - 1:
        b.append(T)
    return sorted(set(b)), depth


def adjusted_segment(features: np.ndarray, alpha: float = 0.5, k: Optional[int] = None, min_distance: int = 32,
                     max_distance: int = 64) -> List[int]:
    """segment.py:52-128: candidate boundaries (at most 15 by threshold), then drop those closer than min_distance
    to the previous kept one and fill gaps longer than max_distance evenly."""
    T = features.shape[0]
    if T == 1:
        return [0]
    sim = cosine_rows(features[:-1], features[1:])
    depth = cal_depth_score(sim)
    b = _threshold_boundaries(depth, alpha, k, cap=15)
    if not b or b[-1] != T:
        b.append(T)
    if b[0] != 0:
        b.insert(0, 0)
    b = sorted(set(b))
    kept = [b[0]]
    for cand in b[1:-1]:
        gap = cand - kept[-1]
        if gap < min_distance:
            continue
        if gap > max_distance:
            extra = int(gap / max_distance)
            start = kept[-1]
            for i in range(1, extra + 1):
                nb = start + round(gap * i / (extra + 1))
                if kept[-1] < nb < cand:
                    kept.append(nb)
        kept.append(cand)
    gap = T - kept[-1]
    if gap >= min_distance or kept[-1] == 0:
        kept.append(T)
    else:
        kept[-1] = T
    return kept


def uniform_segment(T: int, d: int = 32) -> List[int]:
    """segment.py:131-167: the short chunk comes first."""
    if T <= d:
        return [0, T]
    first = T % d
    b = [0] + ([first] if first else [])
    cur = first
    while cur < T:
        cur = min(cur + d, T)
        b.append(cur)
    return b


def segment_left(features: np.ndarray, alpha: float = 0.5, k: Optional[int] = None) -> List[int]:
    """segment.py:226-250."""
    sim = cosine_rows(features[:-1], features[1:])
    depth = cal_depth_score(sim, left_only=True)
    b = _threshold_boundaries(depth, alpha, k)
    if not b:
        b.append(features.shape[0] - 1)
    return b


def _linspace_round(start: int, end: int, steps: int) -> List[int]:
    """torch.linspace(start, end, steps).round().long() in float32 (symmetric evaluation from both ends)."""
    if steps == 1:
        return [int(np.rint(np.float32(start)))]
    step = (np.float32(end) - np.float32(start)) / np.float32(steps - 1)
    half = steps // 2
    out = []
    for i in range(steps):
        v = np.float32(start) + step * np.float32(i) if i < half else np.float32(end) - step * np.float32(steps - 1 - i)
        out.append(int(np.rint(np.float32(v))))
    return out


def sample_scenes_priority(features: np.ndarray, sample_num: int = 32, alpha: float = 0.3, k: Optional[int] = None,
                           randperm: Optional[Callable[[int], Sequence[int]]] = None) -> List[int]:
    """segment.py:252-337.  features [T, P, D]; `randperm(n)` supplies the padding permutation."""
    T = features.shape[0]
    frame_features = features.astype(np.float64).mean(axis=1).astype(np.float32)
    bounds, depth = segment(frame_features, alpha=alpha, k=k)
    if 0 not in bounds:
        bounds = [0] + bounds
    if T not in bounds:
        bounds.append(T)
    bounds = sorted(set(bounds))
    n_scenes = len(bounds) - 1
    picked: List[int] = []
    if n_scenes <= sample_num:
        lengths = [bounds[i + 1] - bounds[i] for i in range(n_scenes)]
        total = sum(lengths)
        budget = [1 + int((sample_num - n_scenes) * ln / total) for ln in lengths]
        while sum(budget) < sample_num:
            budget[sum(budget) % n_scenes] += 1
        while sum(budget) > sample_num:
            budget[budget.index(max(budget))] -= 1
        for i in range(n_scenes):
            s, e = bounds[i], bounds[i + 1]
            if e - s <= budget[i]:
                picked.extend(range(s, e))
            else:
                picked.extend(_linspace_round(s, e - 1, budget[i]))
    else:
        scores = [0.0] + [float(depth[b - 1]) for b in bounds[1:-1]]
        ranked = sorted(range(len(scores)), key=lambda i: -scores[i])[:sample_num]
        for i in ranked:
            picked.append((bounds[i] + bounds[i + 1]) // 2)
    picked = sorted(set(picked))
    if len(picked) < sample_num:
        pool = sorted(set(range(T)) - set(picked))
        need = sample_num - len(picked)
        if len(pool) >= need:
            perm = list(randperm(len(pool)))[:need]
            picked.extend(pool[i] for i in perm)
        else:
            picked.extend(pool)
    return sorted(picked)[:sample_num]


# ------------------------------------------------------------------------------------------------ streaming compression
def _identity(x: np.ndarray, extra):
    T = x.shape[0]
    return x.copy(), extra, [[[i] for i in range(T)]]


def drop_feature(x: np.ndarray, T0: int, coins: Sequence[int]):
    """compress_functions.py:20-56.  Keeps T0 frames; each new frame joins, the frame at the most similar adjacent
    pair (its left or, on coin 1, its right member) leaves.  Returns (features [T0,P,D], sims [T0-1], step indices)."""
    T = x.shape[0]
    if T <= T0:
        return _identity(x, None)
    flat = x.reshape(T, -1)
    rows = list(range(T0))                                  # indices into x, logical order
    sim = [_cos1(flat[i], flat[i + 1]) for i in range(T0 - 1)]
    groups = [[i] for i in range(T0)]
    steps = [list(groups)]
    for n, i in enumerate(range(T0, T)):
        all_rows = rows + [i]
        all_sim = sim + [_cos1(flat[rows[-1]], flat[i])]
        all_groups = groups + [[i]]
        idx = int(np.argmax(np.asarray(all_sim, dtype=np.float32)))
        if coins[n] > 0:
            idx += 1
        rows = all_rows[:idx] + all_rows[idx + 1:]
        if idx == T0:
            sim = all_sim[:T0 - 1]
        elif idx == 0:
            sim = all_sim[1:]
        else:
            sim = all_sim[:idx] + all_sim[idx + 1:]
            sim[idx - 1] = _cos1(flat[all_rows[idx - 1]], flat[all_rows[idx + 1]])
        groups = all_groups[:idx] + all_groups[idx + 1:]
        steps.append(list(groups))
    return x[rows].copy(), np.asarray(sim, dtype=np.float32), steps


def merge_feature(x: np.ndarray, T0: int):
    """compress_functions.py:59-91: the most similar adjacent pair is averaged into its right member."""
    T = x.shape[0]
    if T <= T0:
        return _identity(x, None)
    dt = x.dtype
    flat = x.reshape(T, -1)
    cur = [flat[i].copy() for i in range(T0)]
    sim = [_cos1(cur[i], cur[i + 1]) for i in range(T0 - 1)]
    groups = [[i] for i in range(T0)]
    steps = [list(groups)]
    for i in range(T0, T):
        allf = cur + [flat[i].copy()]
        all_sim = sim + [_cos1(cur[-1], flat[i])]
        all_groups = groups + [[i]]
        idx = int(np.argmax(np.asarray(all_sim, dtype=np.float32)))
        allf[idx + 1] = ((allf[idx] + allf[idx + 1]).astype(dt) / dt.type(2.0)).astype(dt)
        all_groups[idx + 1] = all_groups[idx] + all_groups[idx + 1]
        cur = allf[:idx] + allf[idx + 1:]
        sim = all_sim[:idx] + all_sim[idx + 1:]
        groups = all_groups[:idx] + all_groups[idx + 1:]
        if idx > 0:
            sim[idx - 1] = _cos1(allf[idx - 1], allf[idx + 1])
        if idx + 1 < T0:
            sim[idx] = _cos1(allf[idx + 1], allf[idx + 2])
        steps.append(list(groups))
    return np.stack(cur).reshape((T0,) + x.shape[1:]), np.asarray(sim, dtype=np.float32), steps


def _normalize(v: np.ndarray) -> np.ndarray:
    v64 = v.astype(np.float64)
    return v64 / max(math.sqrt(float((v64 * v64).sum())), 1e-12)


def _pair_matrix(normed: List[np.ndarray]) -> np.ndarray:
    m = np.stack(normed)
    s = (m @ m.T).astype(np.float32)
    np.fill_diagonal(s, NEG)
    return s


def _grow_matrix(sim: np.ndarray, new_col: np.ndarray) -> np.ndarray:
    n = sim.shape[0]
    out = np.full((n + 1, n + 1), NEG, dtype=np.float32)
    out[:n, :n] = sim
    out[:n, n] = new_col
    out[n, :n] = new_col
    return out


def _shrink_matrix(sim: np.ndarray, idx: int) -> np.ndarray:
    keep = [i for i in range(sim.shape[0]) if i != idx]
    return sim[np.ix_(keep, keep)]


def k_drop_feature(x: np.ndarray, T0: int, coins: Sequence[int]):
    """compress_functions.py:176-218: all-pairs similarity; one member of the most similar pair (coin 1: the row
    index of the flat argmax, coin 0: the column index) leaves."""
    T = x.shape[0]
    if T <= T0:
        return _identity(x, None)
    flat = x.reshape(T, -1)
    rows = list(range(T0))
    normed = [_normalize(flat[i]) for i in rows]
    sim = _pair_matrix(normed)
    groups = [[i] for i in range(T0)]
    steps = [list(groups)]
    for n, i in enumerate(range(T0, T)):
        nn_ = _normalize(flat[i])
        col = np.asarray([float(v @ nn_) for v in normed], dtype=np.float32)
        all_sim = _grow_matrix(sim, col)
        flat_idx = int(np.argmax(all_sim))
        left, right = divmod(flat_idx, T0 + 1)
        idx = left if coins[n] > 0 else right
        all_rows = rows + [i]
        all_normed = normed + [nn_]
        all_groups = groups + [[i]]
        rows = all_rows[:idx] + all_rows[idx + 1:]
        normed = all_normed[:idx] + all_normed[idx + 1:]
        groups = all_groups[:idx] + all_groups[idx + 1:]
        sim = _shrink_matrix(all_sim, idx)
        steps.append(list(groups))
    return x[rows].copy(), None, steps


def k_merge_feature(x: np.ndarray, T0: int):
    """compress_functions.py:221-264: the most similar pair (left = row, right = column of the flat argmax) is
    averaged into `right`, `left` leaves, row/column `right` of the similarity matrix is recomputed."""
    T = x.shape[0]
    if T <= T0:
        return _identity(x, None)
    dt = x.dtype
    flat = x.reshape(T, -1)
    cur = [flat[i].copy() for i in range(T0)]
    normed = [_normalize(v) for v in cur]
    sim = _pair_matrix(normed)
    groups = [[i] for i in range(T0)]
    steps = [list(groups)]
    for i in range(T0, T):
        nn_ = _normalize(flat[i])
        col = np.asarray([float(v @ nn_) for v in normed], dtype=np.float32)
        all_sim = _grow_matrix(sim, col)
        allf = cur + [flat[i].copy()]
        all_normed = normed + [nn_]
        all_groups = groups + [[i]]
        left, right = divmod(int(np.argmax(all_sim)), T0 + 1)
        allf[right] = ((allf[left] + allf[right]).astype(dt) / dt.type(2.0)).astype(dt)
        all_normed[right] = _normalize(allf[right])
        all_groups[right] = all_groups[left] + all_groups[right]
        col = np.asarray([float(v @ all_normed[right]) for v in all_normed], dtype=np.float32)
        all_sim[right, :] = col
        all_sim[:, right] = col
        all_sim[right, right] = NEG
        cur = allf[:left] + allf[left + 1:]
        normed = all_normed[:left] + all_normed[left + 1:]
        groups = all_groups[:left] + all_groups[left + 1:]
        sim = _shrink_matrix(all_sim, left)
        steps.append(list(groups))
    return np.stack(cur).reshape((T0,) + x.shape[1:]), sim, steps


# ------------------------------------------------------------------------------------------------ k-means
def kmeans_feature(x: np.ndarray, T0: int, init: Sequence[int], randint: Callable[[int, int], int],
                   weights: Optional[np.ndarray] = None, weighted: bool = False, tol: float = 1e-4, max_iter: int = 10):
    """compress_functions.py:94-131 (plain) and :134-173 (weighted).  `init` = the first T0 entries of the
    reference's randperm; `randint(a, b)` re-seeds an empty cluster from a random frame.  As in the reference the
    centroids returned are those the last label assignment was measured against unless all iterations ran."""
    T = x.shape[0]
    if weighted and weights is None:
        weights = np.ones(T, dtype=x.dtype)
    if T <= T0:
        return _identity(x, weights if weighted else None)
    X = x.reshape(T, -1).astype(np.float64)
    cent = X[list(init)[:T0]].copy()
    w = weights.astype(np.float64) if weighted else np.ones(T)
    labels = np.zeros(T, dtype=np.int64)
    wsum = np.zeros(T0)
    for _ in range(max_iter):
        d = np.sqrt(((X[:, None, :] - cent[None, :, :]) ** 2).sum(-1))
        labels = d.argmin(1)
        new = np.zeros_like(cent)
        wsum = np.zeros(T0)
        empty = []
        for j in range(T0):
            m = labels == j
            if m.any():
                wsum[j] = w[m].sum()
                new[j] = (w[m, None] * X[m]).sum(0) / wsum[j] if weighted else X[m].mean(0)
            else:
                empty.append(j)
        for j in empty:
            new[j] = X[randint(0, T - 1)]
        diff = np.sqrt(((cent - new) ** 2).sum(1)).sum()
        if diff < tol:
            break
        cent = new
    groups = [[j for j in range(T) if labels[j] == i] for i in range(T0)]
    out = cent.astype(x.dtype).reshape((T0,) + x.shape[1:])
    return out, (wsum.astype(x.dtype) if weighted else None), [groups]


# ------------------------------------------------------------------------------------------------ Turing memory
def ntm_get_weight(mem: np.ndarray, new: np.ndarray, wq, bq, wk, bk) -> np.ndarray:
    """memory_builder.py:21-26: softmax(q_proj(mem) k_proj(new)^T / sqrt(output_dim)) over the new tokens."""
    q = mem.astype(np.float64) @ wq.astype(np.float64).T + bq.astype(np.float64)
    k = new.astype(np.float64) @ wk.astype(np.float64).T + bk.astype(np.float64)
    s = q @ k.T / math.sqrt(wq.shape[0])
    s -= s.max(-1, keepdims=True)
    e = np.exp(s)
    return e / e.sum(-1, keepdims=True)


def attention_update(mem: np.ndarray, new: np.ndarray, wq, bq, wk, bk, update_ratio: float = 0.2) -> np.ndarray:
    """memory_builder.py:52-64: memory <- memory * (1 - sum_j w_ij) + w @ new with w = ratio * softmax."""
    w = ntm_get_weight(mem, new, wq, bq, wk, bk) * update_ratio
    decay = w.sum(1, keepdims=True)
    return mem.astype(np.float64) * (1.0 - decay) + w @ new.astype(np.float64)


def attention_feature(x: np.ndarray, T0: int, wq, bq, wk, bk, update_ratio: float = 0.2) -> np.ndarray:
    """compress_functions.py:267-280: the first T0 frames are the memory, later frames are folded in T0 at a time."""
    T, P, D = x.shape
    if T <= T0:
        return x.copy()
    mem = x[:T0].reshape(T0 * P, D).astype(np.float64)
    for i in range(T0, T, T0):
        new = x[i:min(i + T0, T)].reshape(-1, D)
        mem = attention_update(mem, new, wq, bq, wk, bk, update_ratio)
    return mem.reshape(T0, P, D)


def ntm_forward(x: np.ndarray, y: np.ndarray, p: dict) -> np.ndarray:
    """memory_builder.py:28-39 in eval mode (dropouts off): LN(out_proj(softmax(q k^T / sqrt(d)) v)), eps 1e-12."""
    w = ntm_get_weight(x, y, p["q_proj.weight"], p["q_proj.bias"], p["k_proj.weight"], p["k_proj.bias"])
    v = y.astype(np.float64) @ p["v_proj.weight"].astype(np.float64).T + p["v_proj.bias"].astype(np.float64)
    o = (w @ v) @ p["out_proj.weight"].astype(np.float64).T + p["out_proj.bias"].astype(np.float64)
    mu = o.mean(-1, keepdims=True)
    var = ((o - mu) ** 2).mean(-1, keepdims=True)
    return (o - mu) / np.sqrt(var + 1e-12) * p["out_ln.weight"].astype(np.float64) + p["out_ln.bias"].astype(np.float64)


# ------------------------------------------------------------------------------------------------ spatial compression
def compress_spatial_features(x: np.ndarray, compress_size: int = 1) -> np.ndarray:
    """memory_builder.py:72-99 ('mean'): [T, s*s, D] -> [T, c*c, D] by avg_pool2d with window = stride = s // c
    (floor mode: trailing rows / columns that do not fill a window are ignored)."""
    T, N, D = x.shape
    s = round(math.sqrt(N))
    assert s * s == N
    if s == compress_size:
        return x
    if compress_size == 1:
        return x.astype(np.float64).mean(axis=1, keepdims=True).astype(x.dtype)
    k = s // compress_size
    o = (s - k) // k + 1
    g = x.reshape(T, s, s, D)[:, :o * k, :o * k].astype(np.float64).reshape(T, o, k, o, k, D)
    return g.mean(axis=(2, 4)).astype(x.dtype).reshape(-1, compress_size * compress_size, D)


def gelu_erf(x: np.ndarray) -> np.ndarray:
    from math import erf
    return 0.5 * x * (1.0 + np.vectorize(erf)(x / math.sqrt(2.0)))


def compress_temporal_features(x: np.ndarray, *, long_len: int = 3, turing_len: int = 3, cur_len: int = 1,
                               long_size: int = 27, turing_size: int = 27, update_ratio: float = 0.2,
                               sample_type: str = "weighted_kmeans", ntm: dict, mlp: dict, init: Sequence[int] = (),
                               randint: Callable[[int, int], int] = None, coins: Sequence[int] = ()) -> np.ndarray:
    """memory_builder.py:101-190 for one video x [T, 729, 1152]: Turing memory + long memory (+ 3 key frames nearest
    to the heaviest clusters) + current frames, then memory_mlp (Linear-GELU-Linear)."""
    T = x.shape[0]
    c = min(cur_len, T)
    cur, long_m = (x[:0], x) if c == 0 else (x[-c:], x[:-c])
    tur_m = long_m
    if long_size * long_size != long_m.shape[1]:
        long_m = compress_spatial_features(long_m, long_size)
    if turing_size * turing_size != tur_m.shape[1]:
        tur_m = compress_spatial_features(tur_m, turing_size)
    if long_len == 0 or long_m.shape[0] == 0:
        long_c = long_m[:0]
    else:
        if sample_type == "weighted_kmeans":
            long_c, weight, _ = kmeans_feature(long_m, long_len, init, randint, weighted=True)
        elif sample_type == "merge":
            long_c, weight, _ = merge_feature(long_m, long_len)
        elif sample_type == "drop":
            long_c, weight, _ = drop_feature(long_m, long_len, coins)
        else:
            raise NotImplementedError(sample_type)
        order = np.argsort(-weight.astype(np.float64), kind="stable")
        keys = long_m[order][:3].astype(np.float64)
        lm = long_m.astype(np.float64)
        d = np.sqrt(((lm[:, None] - keys[None]) ** 2).sum(axis=3).sum(axis=2))
        cur = np.concatenate([x[d.argmin(0)], cur], axis=0)
    if turing_len == 0 or tur_m.shape[0] == 0:
        tur_c = tur_m[:0]
    else:
        tur_c = attention_feature(tur_m, turing_len, ntm["q_proj.weight"], ntm["q_proj.bias"], ntm["k_proj.weight"],
                                  ntm["k_proj.bias"], update_ratio).astype(x.dtype)
    if long_c.shape[0] < long_len:
        long_c = long_m[:0]
    if tur_c.shape[0] < turing_len:
        tur_c = tur_m[:0]
    D = x.shape[-1]
    mem = np.concatenate([tur_c.reshape(-1, 729, D), long_c.reshape(-1, 729, D), cur.reshape(-1, 729, D)], axis=0)
    flat = mem.reshape(-1, D).astype(np.float64)
    h = gelu_erf(flat @ mlp["0.weight"].astype(np.float64).T + mlp["0.bias"].astype(np.float64))
    out = h @ mlp["2.weight"].astype(np.float64).T + mlp["2.bias"].astype(np.float64)
    return out.reshape(mem.shape)
