"""Host mirror of the reference's Flash-VStream-style memories and scene segmentation: same
function / class names, arguments, return tuples and error behaviour as

    llava/model/memory_module/segment.py               cal_depth_score, segment, adjusted_segment, uniform_segment,
                                                       cal_left_depth_score, segment_left, sample_scenes_priority
    llava/model/memory_module/compress_functions.py    drop_feature, merge_feature, kmeans_feature,
                                                       weighted_kmeans_feature, k_drop_feature, k_merge_feature,
                                                       attention_feature
    llava/model/memory_module/memory_builder.py        NeuralTuringMachine, MultimodalOpsMixin (attention, attention2,
                                                       compress_spatial_features, compress_temporal_features)

All arithmetic on frames runs in libmavlm.so (csrc/legacy_memory.cu; the Turing-memory contractions on the tcgen05
GEMM).  What stays on the host is what the reference also decides on the host: boundary lists from the (tiny)
depth-score vector, random draws (`random.randint`, `torch.randperm`, consumed in the reference's order so that a
seeded run makes the same decisions) and the k-means stopping test.  CPU tensors are an error (no fallback).
"""
from __future__ import annotations

import math
import random
from typing import List, Optional

import torch
from torch import nn

from . import _lib, ops
from ._device import on_tensor_device
from ._lib import ACT_GELU_ERF
from .autograd import gemm_ex
from .modules import uniform_segment_variant  # noqa: F401  (segment.py:169-192 lives with the scheduler)

DROP, MERGE, K_DROP, K_MERGE = 0, 1, 2, 3


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _rows(x: torch.Tensor) -> torch.Tensor:
    ops._need_cuda(x)
    x2 = x.reshape(x.shape[0], -1)
    return x2 if x2.is_contiguous() else x2.contiguous()


# ------------------------------------------------------------------------------------------------ segmentation
@on_tensor_device
def adjacent_cosine(features: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """torch.cosine_similarity(features[:-1], features[1:], eps=eps) over the flattened rows; fp32 [T - 1] holding
    values rounded through features.dtype."""
    x = _rows(features)
    T, L = x.shape
    sim = torch.empty(max(T - 1, 0), dtype=torch.float32, device=x.device)
    if T < 2:
        return sim
    lib = _lib.load()
    code = ops.dtype_code(x)
    ws = _ws(lib.mavlm_adjacent_cosine_workspace_bytes(T, L, code), x.device)
    _lib.check(lib.mavlm_adjacent_cosine_fwd(x.data_ptr(), T, L, x.stride(0), float(eps), sim.data_ptr(), ws.data_ptr(),
                                             ws.numel(), code, ops._stream()), "adjacent_cosine")
    return sim


@on_tensor_device
def _depth(sim: torch.Tensor, left_only: bool) -> torch.Tensor:
    ops._need_cuda(sim)
    s32 = sim if sim.dtype == torch.float32 else ops.cast(sim.contiguous(), torch.float32)
    s32 = s32.contiguous()
    out = torch.empty_like(s32)
    _lib.check(_lib.load().mavlm_depth_scores_fwd(s32.data_ptr(), out.data_ptr(), s32.numel(), int(left_only),
                                                  ops._stream()), "depth_scores")
    return out if sim.dtype == torch.float32 else ops.cast(out, sim.dtype)


def cal_depth_score(sim_scores: torch.Tensor) -> torch.Tensor:
    """segment.py:3-25."""
    return _depth(sim_scores, False)


def cal_left_depth_score(sim_scores: torch.Tensor) -> torch.Tensor:
    """segment.py:210-223."""
    return _depth(sim_scores, True)


@on_tensor_device
def frame_means(features: torch.Tensor) -> torch.Tensor:
    """features.mean(dim=1) for [T, P, D] (segment.py:266, llava_arch.py:528)."""
    ops._need_cuda(features)
    T, P, D = features.shape
    x = features.contiguous()
    out = torch.empty((T, D), dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().mavlm_frame_mean_fwd(x.data_ptr(), out.data_ptr(), T, P, D, ops.dtype_code(x), ops._stream()),
               "frame_mean")
    return out


def _pick_boundaries(depth: torch.Tensor, alpha: float, k: Optional[int], cap: Optional[int] = None) -> List[int]:
    d = depth.detach().float().cpu()                     # the reference syncs here too (.tolist())
    if k is not None:
        b = torch.topk(d, k).indices.sort()[0]
    else:
        std, mean = torch.std_mean(d)
        b = (d > mean + alpha * std).nonzero().squeeze(-1)
        if cap is not None and len(b) > cap:
            b = torch.topk(d, cap).indices.sort()[0]
    return b.tolist()


def segment(features: torch.Tensor, alpha: float = 0.5, k: Optional[int] = None):
    """segment.py:27-49: features [T, D] -> (boundaries, depth_scores)."""
    T = features.shape[0]
    if T == 1:
        return [0], torch.zeros(1)
    sim = adjacent_cosine(features, 1e-2)
    if sim.numel() < 2:
        raise IndexError("index 1 is out of bounds for dimension 0 with size 1")        # segment.py:31 on 2 frames
    sim[0:1].copy_(sim[1:2])
    depth = _depth(sim, False)
    boundaries = _pick_boundaries(depth, alpha, k)
    if boundaries == [] or boundaries[-1] != T - 1:
        boundaries.append(T)
    depth_out = depth if features.dtype == torch.float32 else ops.cast(depth, features.dtype)
    return sorted(set(boundaries)), depth_out


def adjusted_segment(features: torch.Tensor, alpha: float = 0.5, k: Optional[int] = None, min_distance: int = 32,
                     max_distance: int = 64) -> List[int]:
    """segment.py:52-128."""
    T = features.shape[0]
    if T == 1:
        return [0]
    depth = _depth(adjacent_cosine(features, 1e-8), False)
    b = _pick_boundaries(depth, alpha, k, cap=15)
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
    if T - kept[-1] >= min_distance or kept[-1] == 0:
        kept.append(T)
    else:
        kept[-1] = T
    return kept


def uniform_segment(features, d: int = 32) -> List[int]:
    """segment.py:131-167 (short chunk first); only shape[0] is read."""
    T = features if isinstance(features, int) else features.shape[0]
    if T <= d:
        return [0, T]
    first = T % d
    b = [0] + ([first] if first else [])
    cur = first
    while cur < T:
        cur = min(cur + d, T)
        b.append(cur)
    return b


def segment_left(features: torch.Tensor, alpha: float = 0.5, k: Optional[int] = None) -> List[int]:
    """segment.py:226-250."""
    depth = _depth(adjacent_cosine(features, 1e-8), True)
    b = _pick_boundaries(depth, alpha, k)
    if b == []:
        b.append(features.shape[0] - 1)
    return b


def sample_scenes_priority(features: torch.Tensor, sample_num: int = 32, alpha: float = 0.3, k: Optional[int] = None) -> List[int]:
    """segment.py:252-337: exactly `sample_num` distinct frame indices, scene aware."""
    T = features.shape[0]
    bounds, depth = segment(frame_means(features), alpha=alpha, k=k)
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
                picked.extend(torch.linspace(s, e - 1, steps=budget[i]).round().long().tolist())
    else:
        dh = depth.detach().float().cpu()
        scores = [0] + [dh[b - 1].item() for b in bounds[1:-1]]
        ranked = sorted(enumerate(scores), key=lambda it: -it[1])[:sample_num]
        for i, _ in ranked:
            picked.append((bounds[i] + bounds[i + 1]) // 2)
    picked = sorted(set(picked))
    if len(picked) < sample_num:
        pool = sorted(set(range(T)) - set(picked))
        need = sample_num - len(picked)
        if len(pool) >= need:
            picked.extend(pool[i] for i in torch.randperm(len(pool))[:need].tolist())
        else:
            picked.extend(pool)
    return sorted(picked)[:sample_num]


# ------------------------------------------------------------------------------------------------ streaming compression
def _replay_steps(mode: int, T: int, T0: int, decisions, coins) -> list:
    groups = [[i] for i in range(T0)]
    steps = [list(groups)]
    for n, i in enumerate(range(T0, T)):
        allg = groups + [[i]]
        a, b = int(decisions[n][0]), int(decisions[n][1])
        if mode == DROP:
            del allg[a]
        elif mode == MERGE:
            allg[a + 1] = allg[a] + allg[a + 1]
            del allg[a]
        elif mode == K_DROP:
            del allg[a if coins[n] > 0 else b]
        else:
            allg[b] = allg[a] + allg[b]
            del allg[a]
        groups = allg
        steps.append(list(groups))
    return steps


@on_tensor_device
def stream_compress(img_feature: torch.Tensor, video_max_frames: int, mode: int, coins=None, return_steps: bool = True):
    """One launch per streamed frame, decisions on the device (mavlm_stream_compress_fwd).
    Returns (features [T0, P, D], similarities fp32, step_indices or None)."""
    T, T0 = img_feature.shape[0], int(video_max_frames)
    x = _rows(img_feature)
    L = x.shape[1]
    if T <= T0:
        raise RuntimeError(f"mavlm: streaming compression needs more frames ({T}) than it keeps ({T0}); the "
                           "reference returns shorter videos unchanged (see drop_feature & co.)")
    lib = _lib.load()
    code = ops.dtype_code(x)
    dev = x.device
    out = torch.empty((T0,) + tuple(img_feature.shape[1:]), dtype=x.dtype, device=dev)
    n_sim = max(T0 - 1, 1) if mode in (DROP, MERGE) else T0 * T0
    sim = torch.empty(n_sim, dtype=torch.float32, device=dev)
    dec = torch.empty((T - T0, 2), dtype=torch.int32, device=dev)
    coin_t = None
    if mode in (DROP, K_DROP):
        if coins is None or len(coins) < T - T0:
            raise ValueError(f"mavlm: drop / k_drop need one coin per streamed frame ({T - T0}), got "
                             f"{0 if coins is None else len(coins)}")
        coin_t = torch.tensor(list(coins)[:T - T0], dtype=torch.uint8).to(dev)
    ws = _ws(lib.mavlm_stream_compress_workspace_bytes(L, T0, mode, code), dev)
    _lib.check(lib.mavlm_stream_compress_fwd(x.data_ptr(), T, L, T0, mode, ops._ptr(coin_t), out.data_ptr(), sim.data_ptr(),
                                             dec.data_ptr(), ws.data_ptr(), ws.numel(), code, ops._stream()),
               "stream_compress")
    steps = _replay_steps(mode, T, T0, dec.cpu().tolist(), coins) if return_steps else None
    return out, sim, steps


@on_tensor_device
def stream_compress_batched(videos, video_max_frames: int, mode: int, coins=None, return_steps: bool = True):
    """Several independent videos ([T_i, P, D], same P, D, dtype; every T_i > video_max_frames) compressed together:
    one launch per frame index with grid.z = video, so the per-frame latency chain is shared by the batch.
    `coins`: one list per video (drop / k_drop).  Returns a list of (features, similarities, step_indices) with the frames and
    decisions of per-video `stream_compress` calls."""
    T0 = int(video_max_frames)
    xs = [_rows(v) for v in videos]
    B = len(xs)
    L = xs[0].shape[1]
    dev, dt = xs[0].device, xs[0].dtype
    if any(x.shape[1] != L or x.dtype != dt or x.device != dev for x in xs):
        raise RuntimeError("mavlm: batched streaming compression needs videos of one token shape, dtype and device")
    lens = [int(x.shape[0]) for x in xs]
    stride = max(lens) - T0
    lib = _lib.load()
    code = ops.dtype_code(xs[0])
    tail = tuple(videos[0].shape[1:])
    out = torch.empty((B, T0) + tail, dtype=dt, device=dev)
    n_sim = max(T0 - 1, 1) if mode in (DROP, MERGE) else T0 * T0
    sim = torch.empty((B, n_sim), dtype=torch.float32, device=dev)
    dec = torch.empty((B, max(stride, 1), 2), dtype=torch.int32, device=dev)
    coin_t = None
    if mode in (DROP, K_DROP):
        if coins is None or len(coins) != B or any(len(c) < lens[i] - T0 for i, c in enumerate(coins)):
            raise ValueError("mavlm: drop / k_drop need one coin list per video with one coin per streamed frame")
        host = torch.zeros((B, max(stride, 1)), dtype=torch.uint8)
        for i, c in enumerate(coins):
            host[i, :lens[i] - T0] = torch.tensor(list(c)[:lens[i] - T0], dtype=torch.uint8)
        coin_t = host.to(dev)
    ptrs = torch.tensor([x.data_ptr() for x in xs], dtype=torch.int64).to(dev)
    import ctypes
    n_host = (ctypes.c_int64 * B)(*lens)
    ws = _ws(lib.mavlm_stream_compress_batched_workspace_bytes(B, L, T0, mode, code), dev)
    _lib.check(lib.mavlm_stream_compress_batched_fwd(ptrs.data_ptr(), ctypes.cast(n_host, ctypes.c_void_p), B, L, T0, mode,
                                                     ops._ptr(coin_t), max(stride, 1), out.data_ptr(), sim.data_ptr(),
                                                     dec.data_ptr(), ws.data_ptr(), ws.numel(), code, ops._stream()),
               "stream_compress_batched")
    dec_h = dec.cpu().tolist() if return_steps else None
    res = []
    for i in range(B):
        steps = _replay_steps(mode, lens[i], T0, dec_h[i], coins[i] if coins is not None else None) if return_steps else None
        res.append((out[i], sim[i], steps))
    return res


def _short(img_feature, extra):
    return img_feature, extra, [[[i] for i in range(img_feature.shape[0])]]


def _no_presim(img_similarity):
    if img_similarity is not None:
        raise NotImplementedError("mavlm: a precomputed img_similarity is not supported; pass None (the reference's "
                                  "memory path never passes one, memory_builder.py:150)")


def drop_feature(img_feature: torch.Tensor, video_max_frames: int, img_similarity=None):
    """compress_functions.py:20-56."""
    T = img_feature.shape[0]
    if T <= video_max_frames:
        return _short(img_feature, img_similarity)
    _no_presim(img_similarity)
    coins = [random.randint(0, 1) for _ in range(T - video_max_frames)]
    f, s, st = stream_compress(img_feature, video_max_frames, DROP, coins)
    return f, s[:video_max_frames - 1].to(img_feature.dtype), st


def merge_feature(img_feature: torch.Tensor, video_max_frames: int, img_similarity=None):
    """compress_functions.py:59-91."""
    T = img_feature.shape[0]
    if T <= video_max_frames:
        return _short(img_feature, img_similarity)
    _no_presim(img_similarity)
    f, s, st = stream_compress(img_feature, video_max_frames, MERGE)
    return f, s[:video_max_frames - 1].to(img_feature.dtype), st


def k_drop_feature(img_feature: torch.Tensor, video_max_frames: int, img_similarity=None):
    """compress_functions.py:176-218."""
    T = img_feature.shape[0]
    if T <= video_max_frames:
        return _short(img_feature, img_similarity)
    coins = [random.randint(0, 1) for _ in range(T - video_max_frames)]
    f, _, st = stream_compress(img_feature, video_max_frames, K_DROP, coins)
    return f, None, st


def k_merge_feature(img_feature: torch.Tensor, video_max_frames: int, img_similarity=None):
    """compress_functions.py:221-264."""
    T = img_feature.shape[0]
    if T <= video_max_frames:
        return _short(img_feature, img_similarity)
    f, s, st = stream_compress(img_feature, video_max_frames, K_MERGE)
    return f, s.view(video_max_frames, video_max_frames).to(img_feature.dtype), st


# ------------------------------------------------------------------------------------------------ k-means
@on_tensor_device
def _kmeans(X: torch.Tensor, K: int, weights: Optional[torch.Tensor], indices: torch.Tensor, tol: float = 1e-4,
            max_iter: int = 10):
    """Lloyd iterations over whole frames; the host only sees K weight sums + K centroid shifts per iteration."""
    lib = _lib.load()
    T, L = X.shape
    code = ops.dtype_code(X)
    dev = X.device
    cent = X[indices.to(dev)].contiguous()
    new = torch.empty_like(cent)
    labels = torch.empty(T, dtype=torch.int32, device=dev)
    stats = torch.empty(2 * K, dtype=torch.float32, device=dev)          # [wsum | diff]
    w32 = None if weights is None else weights.to(device=dev, dtype=torch.float32).contiguous()
    ws = _ws(lib.mavlm_kmeans_workspace_bytes(T, L, K, code), dev)
    wsum_h = None
    it = 0
    for it in range(max_iter):
        _lib.check(lib.mavlm_kmeans_iter_fwd(X.data_ptr(), ops._ptr(w32), cent.data_ptr(), new.data_ptr(), labels.data_ptr(),
                                             stats.data_ptr(), stats[K:].data_ptr(), None, T, L, K, ws.data_ptr(), ws.numel(),
                                             code, ops._stream()), "kmeans_iter")
        host = stats.cpu()
        wsum_h, diff = host[:K], host[K:].clone()
        empty = [j for j in range(K) if not wsum_h[j] > 0]
        if empty:                                           # "fix nan centroids": re-seed from random frames
            for j in empty:
                new[j].copy_(X[random.randint(0, T - 1)])
            sel = torch.tensor(empty, device=dev)
            d = torch.empty(len(empty), dtype=torch.float32, device=dev)
            a, b = cent[sel].contiguous(), new[sel].contiguous()
            _lib.check(lib.mavlm_row_distance_fwd(a.data_ptr(), b.data_ptr(), d.data_ptr(), len(empty), L, ws.data_ptr(),
                                                  ws.numel(), code, ops._stream()), "row_distance")
            diff[empty] = d.cpu()
        if float(diff.sum()) < tol:
            break
        cent, new = new, cent
    return cent, labels, wsum_h, it


def kmeans_feature(img_feature: torch.Tensor, video_max_frames: int, img_similarity=None):
    """compress_functions.py:94-131."""
    T, P, D = img_feature.shape
    T0 = video_max_frames
    if T <= T0:
        return _short(img_feature, img_similarity)
    X = _rows(img_feature)
    cent, labels, _, _ = _kmeans(X, T0, None, torch.randperm(T)[:T0])
    lab = labels.cpu().tolist()
    return cent.view(T0, P, D), img_similarity, [[[j for j in range(T) if lab[j] == i] for i in range(T0)]]


def weighted_kmeans_feature(img_feature: torch.Tensor, video_max_frames: int, weights: Optional[torch.Tensor] = None):
    """compress_functions.py:134-173."""
    if weights is None:
        weights = torch.ones(img_feature.size(0), dtype=img_feature.dtype, device=img_feature.device)
    T, P, D = img_feature.shape
    T0 = video_max_frames
    if T <= T0:
        return _short(img_feature, weights)
    X = _rows(img_feature)
    cent, labels, wsum, _ = _kmeans(X, T0, weights, torch.randperm(T, device=X.device)[:T0])
    lab = labels.cpu().tolist()
    return (cent.view(T0, P, D), wsum.to(device=X.device, dtype=X.dtype),
            [[[j for j in range(T) if lab[j] == i] for i in range(T0)]])


@on_tensor_device
def frame_distances(frames: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
    """fp32 [T, K] Frobenius distances between whole frames and key frames (memory_builder.py:157)."""
    lib = _lib.load()
    X, C = _rows(frames), _rows(keys)
    T, L = X.shape
    K = C.shape[0]
    code = ops.dtype_code(X)
    dist = torch.empty((T, K), dtype=torch.float32, device=X.device)
    labels = torch.empty(T, dtype=torch.int32, device=X.device)
    stats = torch.empty(2 * K, dtype=torch.float32, device=X.device)
    ws = _ws(lib.mavlm_kmeans_workspace_bytes(T, L, K, code), X.device)
    _lib.check(lib.mavlm_kmeans_iter_fwd(X.data_ptr(), None, C.data_ptr(), None, labels.data_ptr(), stats.data_ptr(),
                                         stats[K:].data_ptr(), dist.data_ptr(), T, L, K, ws.data_ptr(), ws.numel(), code,
                                         ops._stream()), "frame_distances")
    return dist


# ------------------------------------------------------------------------------------------------ Turing memory
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


@on_tensor_device
def _ntm_weight(q: torch.Tensor, k: torch.Tensor, scale: float, ratio: float, mem: Optional[torch.Tensor] = None):
    """w = ratio * softmax(q k^T * scale) as [M, pad8(n)] (zero padding) and, with mem, mem * (1 - rowsum(w))."""
    M, n = q.shape[0], k.shape[0]
    n_pad = _pad8(n)
    scores = torch.empty((M, n_pad), dtype=torch.float32, device=q.device)      # row stride padded for the GEMM's stores
    gemm_ex(q.contiguous(), False, k.contiguous(), True, M, n, q.shape[1], out=scores)
    w = torch.empty((M, n_pad), dtype=q.dtype, device=q.device)
    scaled = None if mem is None else torch.empty_like(mem)
    _lib.check(_lib.load().mavlm_ntm_softmax_fwd(scores.data_ptr(), scores.stride(0), M, n, float(scale), float(ratio),
                                                 w.data_ptr(), n_pad, n_pad, ops._ptr(mem), ops._ptr(scaled),
                                                 0 if mem is None else mem.shape[1], ops.dtype_code(q), ops._stream()),
               "ntm_softmax")
    return w, scaled


def _pad_rows(y: torch.Tensor, n_pad: int) -> torch.Tensor:
    if y.shape[0] == n_pad and y.is_contiguous():
        return y
    out = torch.zeros((n_pad, y.shape[1]), dtype=y.dtype, device=y.device)
    out[:y.shape[0]].copy_(y)
    return out


class NeuralTuringMachine(nn.Module):
    """memory_builder.py:8-39, same parameters / state_dict keys."""

    def __init__(self, input_dim=1152, output_dim=1152, attention_dropout=0.1):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.q_proj = nn.Linear(input_dim, output_dim)
        self.k_proj = nn.Linear(input_dim, output_dim)
        self.v_proj = nn.Linear(input_dim, output_dim)
        self.dropout = nn.Dropout(attention_dropout)
        self.out_proj = nn.Linear(output_dim, input_dim)
        self.out_dropout = nn.Dropout(attention_dropout)
        self.out_ln = nn.LayerNorm(input_dim, eps=1e-12)

    def _qk(self, x, y):
        return (ops.linear(x, self.q_proj.weight, self.q_proj.bias), ops.linear(y, self.k_proj.weight, self.k_proj.bias))

    def get_weight(self, x, y):
        q, k = self._qk(x, y)
        w, _ = _ntm_weight(q, k, 1.0 / math.sqrt(self.output_dim), 1.0)
        return w[:, :y.shape[0]]

    def gated_update(self, memory, new_feature, update_ratio=0.2):
        """memory * (1 - sum_j w_ij) + w @ new_feature with w = update_ratio * softmax (memory_builder.py:52-64)."""
        q, k = self._qk(memory, new_feature)
        w, out = _ntm_weight(q, k, 1.0 / math.sqrt(self.output_dim), update_ratio, memory.contiguous())
        new_pad = _pad_rows(new_feature, w.shape[1])
        gemm_ex(w, False, new_pad, False, memory.shape[0], memory.shape[1], w.shape[1], out=out, accumulate=True)
        return out

    def forward(self, x, y):
        if self.training and (self.dropout.p > 0 or self.out_dropout.p > 0):
            raise NotImplementedError("mavlm: NeuralTuringMachine.forward is inference-only (call .eval()); the "
                                      "reference marks this path deprecated (memory_builder.py:65)")
        q, k = self._qk(x, y)
        w, _ = _ntm_weight(q, k, 1.0 / math.sqrt(self.output_dim), 1.0)
        value = _pad_rows(ops.linear(y, self.v_proj.weight, self.v_proj.bias), w.shape[1])
        ctx = gemm_ex(w, False, value, False, x.shape[0], value.shape[1], w.shape[1])
        out = ops.linear(ctx, self.out_proj.weight, self.out_proj.bias)
        return ops.layernorm(out, self.out_ln.weight, self.out_ln.bias, self.out_ln.eps)


def attention_feature(img_feature: torch.Tensor, video_max_frames: int, attention_fn=None, update_ratio: float = 0.2):
    """compress_functions.py:267-280."""
    T, P, D = img_feature.shape
    T0 = video_max_frames
    if T <= T0:
        return img_feature, None
    memory = img_feature[:T0].reshape(T0 * P, D)
    for i in range(T0, T, T0):
        new_feature = img_feature[i:min(i + T0, T)].reshape(-1, D)
        memory = attention_fn(memory, new_feature, update_ratio=update_ratio)
    return memory.reshape(T0, P, D), None


class MultimodalOpsMixin:
    """memory_builder.py:41-190.  Mixed into the model class like the reference's (llava_arch.py:267); expects
    `self.get_model()` to expose `attention_model` (NeuralTuringMachine) and `memory_mlp`."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.memory_mlp = nn.Sequential(nn.Linear(1152, 1152), nn.GELU(), nn.Linear(1152, 1152))

    def attention(self, turing_memory, new_feature, update_ratio=0.2):
        T1, D1 = turing_memory.shape
        T2, D2 = new_feature.shape
        assert D1 == D2, f"Dimension mismatch: {D1} != {D2}"
        return self.get_model().attention_model.gated_update(turing_memory, new_feature, update_ratio)

    def attention2(self, turing_memory, new_feature, update_ratio=0.2):  # deprecated in the reference too
        T1, D1 = turing_memory.shape
        T2, D2 = new_feature.shape
        assert D1 == D2, f"dimmension not match, {D1} != {D2}"
        return self.get_model().attention_model.forward(turing_memory, new_feature)

    @on_tensor_device
    def compress_spatial_features(self, image_features, compress_size=1):
        compress_type = getattr(self.config, "compress_type", "mean")
        patch_size = round(math.sqrt(image_features.shape[1]))
        assert patch_size * patch_size == image_features.shape[1], (
            f"For ViT feature map, {patch_size}*{patch_size} != {image_features.shape[1]}")
        if patch_size == compress_size:
            return image_features
        if compress_type is not None:
            if "mean" not in compress_type:
                raise NotImplementedError(f"`compress_type` {self.config.compress_type} is not supported yet.")
            if compress_size == 1:
                return frame_means(image_features).unsqueeze(1)
            x = image_features.contiguous()
            ops._need_cuda(x)
            window = patch_size // compress_size
            o = (patch_size - window) // window + 1
            T, _, D = x.shape
            out = torch.empty((T, o * o, D), dtype=x.dtype, device=x.device)
            _lib.check(_lib.load().mavlm_avg_pool_fwd(x.data_ptr(), out.data_ptr(), T, patch_size, window, D,
                                                      ops.dtype_code(x), ops._stream()), "avg_pool")
            return out.view(-1, compress_size * compress_size, D)
        return image_features

    def compress_temporal_features(self, image_features, video_idx_in_batch, all_video=False):
        cfg = self.config
        long_len = getattr(cfg, "video_long_memory_length", 3)
        turing_len = getattr(cfg, "video_Turing_memory_length", 3)
        cur_len = getattr(cfg, "video_current_memory_length", 1)
        long_size = getattr(cfg, "compress_long_memory_size", 27)
        turing_size = getattr(cfg, "compress_Turing_memory_size", 27)
        ratio = getattr(cfg, "compress_Turing_update_ratio", 0.2)
        sample_type = getattr(cfg, "video_sample_type", "weighted_kmeans")
        fns = {"drop": drop_feature, "merge": merge_feature, "kmeans": kmeans_feature,
               "weighted_kmeans": weighted_kmeans_feature, "kdrop": k_drop_feature, "kmerge": k_merge_feature,
               "attention": attention_feature}
        if sample_type not in fns:
            raise NotImplementedError(f"video_sample_type {sample_type} is not supported.")
        compress_fn = fns[sample_type]
        if all_video:
            video_idx_in_batch = list(range(len(image_features)))
        mlp = self.get_model().memory_mlp
        result = []
        for idx, img_feature in enumerate(image_features):
            if idx not in video_idx_in_batch:
                result.append(None)
                continue
            cur_start = min(cur_len, img_feature.shape[0])
            if cur_start == 0:
                cur_memory, long_memory, turing_memory = img_feature[:0], img_feature, img_feature
            else:
                cur_memory, long_memory = img_feature[-cur_start:], img_feature[:-cur_start]
                turing_memory = long_memory
            if long_size * long_size != long_memory.shape[1]:
                long_memory = self.compress_spatial_features(long_memory, long_size)
            if turing_size * turing_size != turing_memory.shape[1]:
                turing_memory = self.compress_spatial_features(turing_memory, turing_size)
            if long_len == 0 or long_memory.shape[0] == 0:
                long_c = long_memory[:0]
            else:
                long_c, weight, _ = compress_fn(long_memory, long_len)
                order = torch.argsort(weight, descending=True)
                keys = long_memory[order][:3]                               # memory_builder.py:153-156 as written
                nearest = torch.argmin(frame_distances(long_memory, keys), dim=0)
                cur_memory = torch.cat([img_feature[nearest], cur_memory], dim=0)
            if turing_len == 0 or turing_memory.shape[0] == 0:
                turing_c = turing_memory[:0]
            else:
                turing_c, _ = attention_feature(turing_memory, turing_len, self.attention, update_ratio=ratio)
            if long_c.shape[0] < long_len:
                long_c = long_memory[:0]
            if turing_c.shape[0] < turing_len:
                turing_c = turing_memory[:0]
            mem = torch.cat([turing_c.reshape(-1, 729, 1152), long_c.reshape(-1, 729, 1152),
                             cur_memory.reshape(-1, 729, 1152)], dim=0)
            flat = mem.view(-1, mem.shape[-1])
            h = ops.linear(flat, mlp[0].weight, mlp[0].bias, act=ACT_GELU_ERF)
            result.append(ops.linear(h, mlp[2].weight, mlp[2].bias).view(*mem.shape))
        return result
