"""CPU oracle for the visual-memory path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a plain numpy restatement of the reference algorithm of
1023604540/Memory-Augmented-VLM for the path

    mm_projector -> bilinear 27x27->14x14 pool -> temporal PE -> chunk scheduler ->
    recurrent memory (formation + evolution) -> memory fuser -> token assembly

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker / the timed CPU
baseline.  The product package never imports anything under ``oracle/``.

Parity pinning: the reference ships no tests / golden vectors for this path,
so the oracle is pinned against outputs of the *reference modules themselves*, executed in
the build container by ``tools/gen_golden.py`` (path-import of the unmodified reference
files) and committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every
function below against those fixtures.

Every function cites the reference file:line it restates (paths relative to the
reference repo root).  All functions are functional (weights passed in a flat dict whose
keys are the reference ``state_dict`` keys) and work in the dtype of ``compute_dtype``
(float64 gives the "truth" the CUDA fp32 path is held to 1e-5 against).
"""
from __future__ import annotations

import math
import os
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

try:  # scipy is in the image; keep a slow fallback so the oracle never silently changes maths
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)

Array = np.ndarray
Weights = Dict[str, Array]

# Hard-coded prompt ids, llava/model/llava_arch.py:708 and :714
MEMORY_PROMPT_IDS = (1986, 374, 264, 1550, 11591, 12126, 315, 279, 2766, 25)
FRAME_PROMPT_IDS = (9485, 525, 48876, 9124, 14087, 504, 279, 2766, 25)


# ----------------------------------------------------------------------------------------
# small building blocks
# ----------------------------------------------------------------------------------------
def linear(x: Array, w: Array, b: Optional[Array]) -> Array:
    """torch.nn.Linear: y = x @ w.T + b, w is [out, in]  (MemoryController.py:23,37-39,64)."""
    y = x @ w.T
    if b is not None:
        y = y + b
    return y


def _rowwise_parallel(fn, x: Array) -> Array:
    """Apply a row-independent numpy function over blocks of the leading axis on all host threads
    (numpy ufuncs release the GIL).  Pure performance helper for the timed CPU baseline: the maths
    of `fn` is unchanged, and small inputs take the direct path."""
    n_thr = min(os.cpu_count() or 1, 64)
    if x.size < (1 << 20) or x.ndim < 2 or n_thr == 1:
        return fn(x)
    lead = x.shape[0]
    if lead < n_thr:                      # e.g. [heads, Lq, Lk]: split the second axis instead
        flat = x.reshape(-1, x.shape[-1])
        return _rowwise_parallel(fn, flat).reshape(x.shape)
    bounds = np.linspace(0, lead, n_thr + 1).astype(int)
    out = np.empty_like(x)

    def work(i):
        out[bounds[i]:bounds[i + 1]] = fn(x[bounds[i]:bounds[i + 1]])

    with ThreadPoolExecutor(max_workers=n_thr) as ex:
        list(ex.map(work, range(n_thr)))
    return out


def gelu_erf(x: Array) -> Array:
    """nn.GELU() default (erf form)  (multimodal_projector/builder.py:46, llava_arch.py:134)."""
    return _rowwise_parallel(lambda t: 0.5 * t * (1.0 + _erf(t / math.sqrt(2.0))), x)


def relu(x: Array) -> Array:
    """ACT2FN['relu']  (MemoryController.py:65 with llava_arch.py:120)."""
    return np.maximum(x, 0)


def layer_norm(x: Array, gamma: Array, beta: Array, eps: float) -> Array:
    """nn.LayerNorm over the last dim, biased variance  (MemoryController.py:24)."""
    mu = x.mean(axis=-1, keepdims=True)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True)
    return xc / np.sqrt(var + eps) * gamma + beta


def _softmax_block(s: Array) -> Array:
    m = s.max(axis=-1, keepdims=True)
    e = np.exp(s - m)
    return e / e.sum(axis=-1, keepdims=True)


def softmax_lastdim(s: Array) -> Array:
    return _rowwise_parallel(_softmax_block, s)


# ----------------------------------------------------------------------------------------
# a1: mm_projector (mlp2x_gelu)
# ----------------------------------------------------------------------------------------
def mm_projector(x: Array, w: Weights, prefix: str = "mm_projector.") -> Array:
    """Linear(Dv,D) -> GELU(erf) -> Linear(D,D)   (multimodal_projector/builder.py:41-48,
    called at llava_arch.py:302)."""
    h = gelu_erf(linear(x, w[prefix + "0.weight"], w[prefix + "0.bias"]))
    return linear(h, w[prefix + "2.weight"], w[prefix + "2.bias"])


# ----------------------------------------------------------------------------------------
# a2: get_2dPool  (bilinear, align_corners=False, no antialias)
# ----------------------------------------------------------------------------------------
def bilinear_taps(n_in: int, n_out: int) -> Tuple[Array, Array, Array]:
    """Source taps of F.interpolate(mode='bilinear', align_corners=False) along one axis:
    src = (o + 0.5) * n_in / n_out - 0.5, clamped at 0; i1 = min(i0 + 1, n_in - 1)
    (ATen upsample_bilinear2d semantics used by llava_arch.py:291)."""
    scale = n_in / n_out
    o = np.arange(n_out, dtype=np.float64)
    src = np.maximum((o + 0.5) * scale - 0.5, 0.0)
    i0 = np.floor(src).astype(np.int64)
    i0 = np.minimum(i0, n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    lam = src - i0
    return i0, i1, lam


def get_2d_pool(x: Array, stride: int = 2, mode: str = "bilinear", side: Optional[int] = None) -> Array:
    """[F, side*side, D] -> [F, ceil(side/stride)^2, D]   (llava_arch.py:277-297).
    'average' / 'max' are the reference's other two modes (kernel=stride, floor)."""
    f, n, d = x.shape
    if side is None:
        side = int(round(math.sqrt(n)))
    assert side * side == n
    g = x.reshape(f, side, side, d)
    if mode == "bilinear":
        out_side = math.ceil(side / stride)
        i0, i1, lam = bilinear_taps(side, out_side)
        lam = lam.astype(x.dtype)
        wy0 = (1 - lam)[None, :, None, None]
        wy1 = lam[None, :, None, None]
        rows = g[:, i0] * wy0 + g[:, i1] * wy1                       # [F, out, side, D]
        wx0 = (1 - lam)[None, None, :, None]
        wx1 = lam[None, None, :, None]
        out = rows[:, :, i0] * wx0 + rows[:, :, i1] * wx1            # [F, out, out, D]
    elif mode in ("average", "max"):
        out_side = side // stride
        g = g[:, : out_side * stride, : out_side * stride]
        g = g.reshape(f, out_side, stride, out_side, stride, d)
        out = g.mean(axis=(2, 4)) if mode == "average" else g.max(axis=(2, 4))
    else:
        raise ValueError(f"Unexpected mm_spatial_pool_mode: {mode}")      # llava_arch.py:294
    return out.reshape(f, -1, d)


# ----------------------------------------------------------------------------------------
# a3: TemporalPositionalEncoding (fixed sinusoid)
# ----------------------------------------------------------------------------------------
def temporal_pe_table(max_frames: int, dim: int) -> Array:
    """fp32 table, interleaved sin/cos  (position_encoding.py:29-36).  Computed in float32
    like the reference so that the table is bit-comparable."""
    pos = np.arange(max_frames, dtype=np.float32)[:, None]
    div = np.exp(np.arange(0, dim, 2, dtype=np.float32) * np.float32(-(math.log(10000.0) / dim))).astype(np.float32)
    pe = np.zeros((max_frames, dim), dtype=np.float32)
    ang = (pos * div).astype(np.float32)
    pe[:, 0::2] = np.sin(ang)
    pe[:, 1::2] = np.cos(ang)
    return pe


def add_temporal_pe(x: Array, frame_idx: Optional[Array], table: Array) -> Array:
    """x[T,N,C] + table[frame_idx][:,None,:] cast to x.dtype first  (position_encoding.py:57-64);
    index errors as position_encoding.py:73-76."""
    if x.ndim != 3:
        raise ValueError(f"Expected 3D input, got {x.ndim}D.")
    if frame_idx is None:
        frame_idx = np.arange(x.shape[0])
    frame_idx = np.asarray(frame_idx)
    if np.any(frame_idx >= table.shape[0]):
        raise ValueError(f"indices exceed max_frames: max {frame_idx.max()} vs limit {table.shape[0]}")
    if np.any(frame_idx < 0):
        raise ValueError(f"indices contains negative values: min {frame_idx.min()}")
    return x + table[frame_idx].astype(x.dtype)[:, None, :]


# ----------------------------------------------------------------------------------------
# a4 / a5: index generation (frame sampling, fine frames, chunk boundaries)
# ----------------------------------------------------------------------------------------
def _torch_linspace_f32(start: float, end: float, steps: int) -> Array:
    """torch.linspace(float32) evaluates symmetrically: first half start + i*step, second half
    end - (steps-1-i)*step, with step = (end-start)/(steps-1) in float32 (ATen RangeFactories)."""
    if steps == 1:
        return np.array([start], dtype=np.float32)
    start = np.float32(start)
    end = np.float32(end)
    step = np.float32((end - start) / np.float32(steps - 1))
    i = np.arange(steps)
    half = steps // 2
    lo = (start + step * i.astype(np.float32)).astype(np.float32)
    hi = (end - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < half, lo, hi).astype(np.float32)


def sample_frame_indices(num_frames: int) -> Array:
    """F<32 -> all; else F'=(F//32)*32 (at least 64); idx = linspace(0,F-1,F').long()
    (truncation)   (llava_arch.py:437-451)."""
    if num_frames < 32:
        n = num_frames
    else:
        n = (num_frames // 32) * 32
        if n < 64:
            n = 64
    return _torch_linspace_f32(0, num_frames - 1, n).astype(np.int64)


def fine_frame_indices(num_sampled: int, max_fine: int = 32) -> Array:
    """round(linspace(0,F'-1,min(32,F'))) clamped  (llava_arch.py:513-522); torch.round is
    round-half-to-even like numpy."""
    n = min(max_fine, num_sampled)
    idx = np.round(_torch_linspace_f32(0, num_sampled - 1, n)).astype(np.int64)
    return np.clip(idx, 0, num_sampled - 1)


def uniform_segment_variant(t: int, d: int = 32) -> List[int]:
    """[0, d, 2d, ..., T] with the remainder as the last chunk  (segment.py:169-192)."""
    bounds = [0]
    cur = 0
    while cur + d <= t:
        cur += d
        bounds.append(cur)
    if cur < t:
        bounds.append(t)
    return bounds


# ----------------------------------------------------------------------------------------
# a7-a11: recurrent memory transformer
# ----------------------------------------------------------------------------------------
def mha(xq: Array, xkv: Array, w: Weights, p: str, heads: int, return_probs: bool = False):
    """Attention.forward without the residual block  (MemoryController.py:47-54):
    q/k/v projections, split heads, softmax(q k^T / sqrt(dh)) v, merge heads.
    xq [Lq,D], xkv [Lk,D] (batch of 1 squeezed)."""
    d = xq.shape[-1]
    dh = d // heads
    q = linear(xq, w[p + "q_proj.weight"], w[p + "q_proj.bias"]).reshape(-1, heads, dh).transpose(1, 0, 2)
    k = linear(xkv, w[p + "k_proj.weight"], w[p + "k_proj.bias"]).reshape(-1, heads, dh).transpose(1, 0, 2)
    v = linear(xkv, w[p + "v_proj.weight"], w[p + "v_proj.bias"]).reshape(-1, heads, dh).transpose(1, 0, 2)
    scores = (q @ k.transpose(0, 2, 1)) / math.sqrt(dh)
    probs = softmax_lastdim(scores)                                # [H,Lq,Lk]
    ctx = (probs @ v).transpose(1, 0, 2).reshape(-1, d)
    return (ctx, probs) if return_probs else (ctx, None)


def residual_block(h: Array, x: Array, w: Weights, p: str, eps: float) -> Array:
    """Residual.forward: LayerNorm(dense(h) + x)  (MemoryController.py:26-29)."""
    y = linear(h, w[p + "dense.weight"], w[p + "dense.bias"])
    return layer_norm(y + x, w[p + "layernorm.weight"], w[p + "layernorm.bias"], eps)


def attention_block(xq: Array, xkv: Array, w: Weights, p: str, heads: int, eps: float, return_probs=False):
    """Attention.forward  (MemoryController.py:47-57)."""
    ctx, probs = mha(xq, xkv, w, p, heads, return_probs)
    return residual_block(ctx, xq, w, p + "residual.", eps), probs


def transformer_layer(mem: Array, frames: Array, w: Weights, p: str, heads: int, eps: float, return_probs=False):
    """TransformerLayer.forward  (MemoryController.py:69-72)."""
    a, probs = attention_block(mem, frames, w, p + "memory_segment_fusion_attention.", heads, eps, return_probs)
    up = relu(linear(a, w[p + "mlp.0.weight"], w[p + "mlp.0.bias"]))
    return residual_block(up, a, w, p + "residual.", eps), probs


def rmt_chunk(chunk: Array, cache: List[Array], w: Weights, *, prefix: str = "recurrent_memory_transformer.",
              heads: int = 8, depth: int = 2, eps: float = 1e-12, cache_cap: int = 10,
              want_scores: bool = False) -> Tuple[List[Array], Optional[Array]]:
    """TransformerProjector.forward for one chunk  (MemoryController.py:118-158).
    chunk [C,P,D]; cache = list of [M,P,D] states (mutated copy returned)."""
    c, p_, d = chunk.shape
    init = w[prefix + "initial_memory"] + w[prefix + "memory_pos_embed"]          # :123
    m_slots = init.shape[0]
    if cache:                                                                      # :125-127, :89-97
        mem = cache[-1].reshape(-1, d)
        kv = np.concatenate([s.reshape(-1, d) for s in cache], axis=0)
        mem, _ = attention_block(mem, kv, w, prefix + "memory_update_attention.", heads, eps)
    else:
        mem = init.astype(chunk.dtype).reshape(-1, d)
    frames = chunk.reshape(c * p_, d)                                              # :129
    probs = None
    for l in range(depth):                                                         # :132-139
        mem, probs = transformer_layer(mem, frames, w, f"{prefix}layers.{l}.", heads, eps,
                                       return_probs=want_scores and l == depth - 1)
    score = None
    if want_scores:                                                                # :135-137
        score = probs.sum(axis=0).sum(axis=0).reshape(c, p_).mean(axis=1)
    cache = list(cache) + [mem.reshape(m_slots, p_, d)]                            # :152
    if len(cache) > cache_cap:                                                     # :153-154
        cache = cache[-cache_cap:]
    return cache, score


def rmt_video(frames: Array, w: Weights, *, chunk: int = 32, **kw) -> Tuple[List[Array], List[Array]]:
    """Chunk scheduler: reset cache, walk uniform chunks sequentially  (llava_arch.py:528-537)."""
    bounds = uniform_segment_variant(frames.shape[0], chunk)
    cache: List[Array] = []
    scores: List[Array] = []
    for i in range(len(bounds) - 1):
        cache, s = rmt_chunk(frames[bounds[i]:bounds[i + 1]], cache, w, **kw)
        if s is not None:
            scores.append(s)
    return cache, scores


# ----------------------------------------------------------------------------------------
# a13 / a13': fusers
# ----------------------------------------------------------------------------------------
def memory_fuser_mlp(x: Array, w: Weights, prefix: str = "memory_fuser.") -> Array:
    """Linear(D,4D) -> GELU(erf) -> Linear(4D,D)  (llava_arch.py:132-136, called :546)."""
    h = gelu_erf(linear(x, w[prefix + "0.weight"], w[prefix + "0.bias"]))
    return linear(h, w[prefix + "2.weight"], w[prefix + "2.bias"])


def memory_fuser_encoder(x: Array, w: Weights, *, prefix: str = "", num_layers: int, heads: int = 4,
                         eps: float = 1e-5) -> Array:
    """MemoryFuser.forward in eval mode (dropout off)  (MemoryFuser.py:4-30): input_proj ->
    nn.TransformerEncoder(post-LN layers, self-attention over the sequence dim of each batch
    row, GELU FFN 4D) -> output_proj.  x is [B,S,D]; weights use nn.TransformerEncoderLayer
    key names (self_attn.in_proj_weight, linear1, linear2, norm1, norm2)."""
    b, s, d = x.shape
    dh = d // heads
    h = linear(x, w[prefix + "input_proj.weight"], w[prefix + "input_proj.bias"])
    for l in range(num_layers):
        p = f"{prefix}transformer_encoder.layers.{l}."
        qkv = linear(h, w[p + "self_attn.in_proj_weight"], w[p + "self_attn.in_proj_bias"])
        q, k, v = np.split(qkv, 3, axis=-1)
        def sh(t):
            return t.reshape(b, s, heads, dh).transpose(0, 2, 1, 3)
        q, k, v = sh(q), sh(k), sh(v)
        pr = softmax_lastdim((q @ k.transpose(0, 1, 3, 2)) / math.sqrt(dh))
        ctx = (pr @ v).transpose(0, 2, 1, 3).reshape(b, s, d)
        a = linear(ctx, w[p + "self_attn.out_proj.weight"], w[p + "self_attn.out_proj.bias"])
        h = layer_norm(h + a, w[p + "norm1.weight"], w[p + "norm1.bias"], eps)
        f = linear(gelu_erf(linear(h, w[p + "linear1.weight"], w[p + "linear1.bias"])),
                   w[p + "linear2.weight"], w[p + "linear2.bias"])
        h = layer_norm(h + f, w[p + "norm2.weight"], w[p + "norm2.bias"], eps)
    return linear(h, w[prefix + "output_proj.weight"], w[prefix + "output_proj.bias"])


# ----------------------------------------------------------------------------------------
# a14 / a15: type embeddings + token assembly
# ----------------------------------------------------------------------------------------
def assemble_sequence(mem: Array, fine: Array, w: Weights, *, prompt_mem: Array, prompt_frm: Array,
                      drop_frames: bool = False) -> Array:
    """mem [n*M,P,D] (fuser output), fine [Nf,P,D] (PE'd frames):
    + token_type_embedding[0/1] (llava_arch.py:548-554); flatten + image_newline each
    (:620-629); cat(prompt_mem, mem, prompt_frm, fine) or, when frames are dropped,
    cat(prompt_mem, mem)  (:705-731)."""
    e = w["token_type_embedding.weight"]
    nl = w["image_newline"][None, :]
    d = mem.shape[-1]
    mem_t = np.concatenate([(mem + e[0]).reshape(-1, d), nl], axis=0)
    fine_t = np.concatenate([(fine + e[1]).reshape(-1, d), nl], axis=0)
    if drop_frames:
        return np.concatenate([prompt_mem, mem_t], axis=0)
    return np.concatenate([prompt_mem, mem_t, prompt_frm, fine_t], axis=0)


# ----------------------------------------------------------------------------------------
# whole path
# ----------------------------------------------------------------------------------------
def visual_memory_path(tower_out: Array, frame_idx: Optional[Array], w: Weights, *, pe_table: Array,
                       prompt_mem: Array, prompt_frm: Array, chunk: int = 32, max_fine: int = 32,
                       heads: int = 8, depth: int = 2, eps: float = 1e-12, cache_cap: int = 10,
                       pooled_input: bool = False) -> Dict[str, Array]:
    """encode_images (projector) -> get_2dPool -> PE -> fine-frame pick -> RMT over chunks ->
    fuser -> type-embed -> assembly, for ONE video  (llava_arch.py:481-557, 613-629, 705-731).
    tower_out: [F',729,Dv] (or already pooled+projected [F',196,D] when pooled_input)."""
    if pooled_input:
        z = tower_out
    else:
        y = mm_projector(tower_out, w)
        z = get_2d_pool(y, 2, "bilinear")
    z = add_temporal_pe(z, frame_idx, pe_table)
    fine = z[fine_frame_indices(z.shape[0], max_fine)]
    cache, _ = rmt_video(z, w, chunk=chunk, heads=heads, depth=depth, eps=eps, cache_cap=cache_cap)
    cat = np.concatenate(cache, axis=0)
    fused = memory_fuser_mlp(cat, w)
    seq = assemble_sequence(fused, fine, w, prompt_mem=prompt_mem, prompt_frm=prompt_frm)
    return {"pooled_pe": z, "fine": fine, "states": cache, "fused": fused, "sequence": seq}


def normalized_max_error(a: Array, b: Array) -> float:
    """err(a,b) = max|a-b| / max|b|   (SURVEY.md §8d pass criterion)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
