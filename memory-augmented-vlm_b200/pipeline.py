"""Chunk scheduler + fused visual-memory pipeline (the hot path of
llava_arch.py:481-557, 613-629, 705-731 for a batch of equal-length videos).

    tower tokens [B, F, 729, Dv]
      -> mm_projector (2 GEMMs, GELU fused)                                  K1
      -> bilinear 27x27 -> 14x14 pool + temporal PE (one kernel)             K2+K3
      -> frame-side K/V of BOTH formation layers for ALL chunks in one GEMM  K5   (memory independent)
      -> for each chunk, sequentially:  evolution attention over the ring of cached states (their K/V
         projections cached, newest state projected once) -> 2 x [Q proj, fused attention, O proj +
         residual -> LN, MLP up (ReLU fused), MLP down + residual -> LN]    K5-K8, K10
      -> memory fuser (GEMM + GELU, GEMM + bias + type_emb[0]) written straight into the final
         sequence buffer; prompts / newline / fine frames + type_emb[1] by the assembly kernel  K11-K13

Differences from the reference that do not change the result: `image.mean(dim=1)` (K4) is never
computed (only its length is used, segment.py:180); cached states live in a ring buffer whose slot
order differs from `torch.cat(cache)` -- attention is invariant to key order, and the fuser output is
written to the reference's row positions; probs are never materialised.  Per-video results equal
B independent reference calls (the reference supports one video per rank, llava_arch.py:436).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import ops
from ._lib import ACT_GELU_ERF, ACT_NONE, ACT_RELU
from .modules import (Attention, MemoryFuser, MemoryFuserMLP, TemporalPositionalEncoding, TransformerProjector, VisionProjector,
                      fine_frame_indices, uniform_segment_variant)

MEMORY_PROMPT_IDS = (1986, 374, 264, 1550, 11591, 12126, 315, 279, 2766, 25)      # llava_arch.py:708
FRAME_PROMPT_IDS = (9485, 525, 48876, 9124, 14087, 504, 279, 2766, 25)            # llava_arch.py:714


class VisualMemoryPipeline(nn.Module):
    """Holds (references to) the drop-in modules and runs the whole path for B videos."""

    def __init__(self, *, mm_projector: Optional[VisionProjector], recurrent_memory_transformer: TransformerProjector,
                 memory_fuser: "MemoryFuserMLP | MemoryFuser", positional_encoding: TemporalPositionalEncoding,
                 token_type_embedding: nn.Embedding, image_newline: torch.Tensor, embed_tokens: nn.Embedding,
                 chunk_size: int = 32, max_fine_frames: int = 32, num_patches_per_side: int = 27,
                 pool_stride: int = 2, projector_frames_per_pass: int = 64, pool_before_w2: bool = True,
                 tail_fill: bool = False):
        super().__init__()
        self.mm_projector = mm_projector
        self.recurrent_memory_transformer = recurrent_memory_transformer
        self.memory_fuser = memory_fuser
        self.positional_encoding = positional_encoding
        self.token_type_embedding = token_type_embedding
        self.image_newline = image_newline
        self.embed_tokens = embed_tokens
        self.chunk_size = chunk_size
        self.max_fine_frames = max_fine_frames
        self.side = num_patches_per_side
        self.pool_stride = pool_stride
        self.projector_frames_per_pass = projector_frames_per_pass
        self.pool_before_w2 = pool_before_w2
        self.pool_mode = "bilinear"
        # one video, tensor-core tier: the frame-side K/V projection of chunk t+1 and the fuser MLP of finished states
        # are computed in the idle tile slots of chunk t's 1568-row GEMMs (ops.linear_fill) instead of in launches of
        # their own; same results bit for bit.  OFF by default: measured at parity on B200 (profiles/r2_fill_bench.json:
        # the fill tiles take L2 bandwidth the last wave's own tiles were using, and the forced 256 x 256 pair tile costs
        # the critical GEMM 8 us against the 256 x 192 one the heuristic picks) -- kept as a switch, tested bitwise
        self.tail_fill = tail_fill
        self._consts: Dict = {}

    # ------------------------------------------------------------------------------------------
    def _const_ids(self, device):
        key = str(device)
        if key not in self._consts:
            self._consts[key] = (torch.tensor(MEMORY_PROMPT_IDS, dtype=torch.int64, device=device),
                                 torch.tensor(FRAME_PROMPT_IDS, dtype=torch.int64, device=device))
        return self._consts[key]

    def prepare_constants(self, frames: int, device) -> None:
        """Fill the small index caches (prompt ids, fine-frame pick) for videos of `frames` frames on `device`, so that a
        later CUDA-graph capture of memory_forward at that length contains no host-to-device copy."""
        dev = torch.device(device)
        self._const_ids(dev)
        fkey = ("fine", frames, str(dev))
        if fkey not in self._consts:
            self._consts[fkey] = fine_frame_indices(frames, self.max_fine_frames).to(dev)

    def sequence_length(self, n_states: int, n_fine: int, drop_frames: bool = False) -> int:
        rmt = self.recurrent_memory_transformer
        lq = rmt.num_memory_tokens * rmt.patch_size
        n = len(MEMORY_PROMPT_IDS) + n_states * lq + 1
        if not drop_frames:
            n += len(FRAME_PROMPT_IDS) + n_fine * rmt.patch_size + 1
        return n

    def _formation_kv_weights(self):
        """[Wk0;Wv0;Wk1;Wv1;...] so the frame-side K/V of every layer come out of one GEMM."""
        rmt = self.recurrent_memory_transformer
        atts = [l.memory_segment_fusion_attention for l in rmt.layers]
        packs = [a.packed() for a in atts]
        key = tuple(a._pack_key for a in atts)                          # (data_ptr, version, ...) of the source parameters
        c = self._consts.get("fkv")
        if c is None or c[0] != key:
            w = torch.cat([p["wkv"] for p in packs], dim=0).contiguous()
            b = torch.cat([p["bkv"] for p in packs], dim=0).contiguous()
            self._consts["fkv"] = (key, w, b)
        return self._consts["fkv"][1], self._consts["fkv"][2]

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode_frames(self, tower_tokens: torch.Tensor, frame_idx: torch.Tensor, *, validate: bool = True,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[N, side*side, Dv] tower tokens (+ original frame indices [N]) -> pooled + PE'd [N, P, D]
        (encode_images -> get_2dPool -> positional_encoding, llava_arch.py:481-511).

        Bilinear pooling is a fixed convex combination of tokens, so it commutes with the projector's
        second (affine) layer: pool(h W2^T + b2) = pool(h) W2^T + b2 (tap weights sum to 1; verified to
        5e-16 in fp64).  With `pool_before_w2` (default) the second GEMM runs on 196 instead
        of 729 rows per frame (3.7x fewer); `pool_before_w2=False` keeps the reference's operation order.
        `out` [N, P, D]: destination (e.g. this rank's slot of an all-gather buffer, written by the last kernel)."""
        pe = self.positional_encoding
        if validate:
            pe.validate(frame_idx)
        frame_idx = frame_idx.to(tower_tokens.device)
        n = tower_tokens.shape[0]
        outs = []
        step = max(1, self.projector_frames_per_pass)
        table = pe.table()
        mp = self.mm_projector
        for i in range(0, n, step):
            x = tower_tokens[i:i + step]
            if self.pool_before_w2 and self.pool_mode == "bilinear":
                h = ops.linear(x, mp[0].weight, mp[0].bias, act=ACT_GELU_ERF)
                hp = ops.pool_pe(h, side=self.side, stride=self.pool_stride, mode="bilinear")
                outs.append(ops.linear_pe(hp, mp[2].weight, mp[2].bias, table, frame_idx[i:i + step],
                                          out=None if out is None else out[i:i + step]))
            else:
                y = mp(x)
                res = ops.pool_pe(y, side=self.side, stride=self.pool_stride, mode=self.pool_mode,
                                  pe_table=table, frame_idx=frame_idx[i:i + step])
                if out is not None:
                    out[i:i + step].copy_(res)
                outs.append(res)
        if out is not None:
            return out
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    @torch.no_grad()
    def memory_forward(self, z: torch.Tensor, *, seq_out: Optional[torch.Tensor] = None, drop_frames: bool = False,
                       return_states: bool = True, boundaries: Optional[Sequence[int]] = None,
                       piece_frames: Optional[int] = None, before_piece=None,
                       frames_assembled: bool = False) -> Dict[str, torch.Tensor]:
        """z: pooled + PE'd frames [B, F, P, D].  Runs the recurrence, the fuser and the assembly.
        `frames_assembled`: the caller has already run `assemble_frames(seq_out, z)` (every row of the sequence that
        does not hold a memory token: they depend on z only, so a host-streaming front end can send them back while the
        recurrence runs); only the memory-token rows are written here.
        `boundaries`: chunk boundaries [0, ..., F] replacing the uniform scheduler of llava_arch.py:528-534, e.g. the
        scene-based ones of `legacy.adjusted_segment(legacy.frame_means(z[0]))` (segment.py:52-128).
        `piece_frames` / `before_piece`: the frames arrive in pieces of `piece_frames` frames (a multiple of the chunk
        size; the frame-sharded pre-pass of dist.py all-gathers them piece by piece on a side stream): the frame-side
        K/V projection is then issued per piece, right before the piece's first chunk and after `before_piece(j)`
        (which makes the current stream wait for piece j), so the recurrence over piece j overlaps the arrival of
        piece j + 1.  Same arithmetic as the single projection GEMM (rows are independent)."""
        rmt = self.recurrent_memory_transformer
        b, f, p, d = z.shape
        dtype, dev = z.dtype, z.device
        m_slots, lq = rmt.num_memory_tokens, rmt.num_memory_tokens * p
        heads = rmt.layers[0].memory_segment_fusion_attention.num_attention_heads
        dh = d // heads
        scale = 1.0 / math.sqrt(dh)
        cap = rmt.cache_size
        z2 = z.reshape(b, f * p, d)

        # frame-side K/V for all chunks and layers (memory independent): one GEMM
        packs = [l.memory_segment_fusion_attention.packed() for l in rmt.layers]
        dhp = packs[0]["dhp"]
        hd = heads * dhp
        wf, bf = self._formation_kv_weights()
        fz = self.memory_fuser
        fill = (self.tail_fill and b == 1 and dtype != torch.float32 and piece_frames is None
                and not isinstance(fz, MemoryFuser) and d % 8 == 0)
        kv_work: List = []
        fillers: List = []                                              # GemmWork in the order they should be picked
        if fill:
            kvf = torch.empty((b, f * p, wf.shape[0]), dtype=dtype, device=dev)
        elif piece_frames is None:
            kvf = ops.linear(z2, wf, bf)                                # [B, F*P, depth*2*hd]
        else:
            if b != 1 or boundaries is not None or piece_frames % self.chunk_size != 0:
                raise ValueError("mavlm: piece-wise arrival needs one video, the uniform scheduler and pieces of whole chunks")
            kvf = torch.empty((b, f * p, wf.shape[0]), dtype=dtype, device=dev)

        if boundaries is None:
            bounds = uniform_segment_variant(f, self.chunk_size)
        else:
            bounds = [int(v) for v in boundaries]
            if len(bounds) < 2 or bounds[0] != 0 or bounds[-1] != f or any(b1 <= b0 for b0, b1 in zip(bounds, bounds[1:])):
                raise ValueError(f"mavlm: chunk boundaries must rise from 0 to the frame count {f}, got {bounds}")
        n_chunks = len(bounds) - 1
        n_keep = min(n_chunks, cap)
        first = n_chunks - n_keep
        if fill:
            kv_work = [ops.GemmWork(z2[0, bounds[t] * p: bounds[t + 1] * p], wf, bf, kvf[0, bounds[t] * p: bounds[t + 1] * p])
                       for t in range(n_chunks)]
            kv_work[0].run()
            n_fine = min(self.max_fine_frames, f)
            if seq_out is None:
                seq_out = torch.empty((b, self.sequence_length(n_keep, n_fine, drop_frames), d), dtype=dtype, device=dev)
            fuse_hidden = torch.empty((n_keep, lq, fz[0].weight.shape[0]), dtype=dtype, device=dev)
            fuse_work: List = []
        lin = ops.linear_fill if fill else (lambda *a, fillers=(), **k: ops.linear(*a, **k))
        evo = rmt.memory_update_attention
        evo_p = evo.packed()
        ring_states = torch.empty((b, cap, lq, d), dtype=dtype, device=dev)
        # evolution attention: a new state is projected ONCE, by one GEMM with [Wq; Wk; Wv] (N = 3*H*dh: 294 pair
        # tiles = 3.97 waves on 74 CTA pairs, where q and k|v separately cost 2 + 3 waves) into its ring slot
        # (q | k | v columns); q is read by the next chunk, k | v by every later chunk that still caches the state
        ring_qkv = torch.empty((b, cap * lq, 3 * hd), dtype=dtype, device=dev) if n_chunks > 1 else None
        if n_chunks > 1:
            ekey = evo._pack_key                                        # (data_ptr, version, ...) of the source parameters
            c = self._consts.get("evo_qkv")
            if c is None or c[0] != ekey:
                self._consts["evo_qkv"] = (ekey, torch.cat([evo_p["wq"], evo_p["wkv"]], dim=0).contiguous(),
                                           torch.cat([evo_p["bq"], evo_p["bkv"]], dim=0).contiguous())
            w_evo, b_evo = self._consts["evo_qkv"][1], self._consts["evo_qkv"][2]

        mem = rmt.initial_state(dtype).reshape(1, lq, d).expand(b, lq, d).contiguous()
        last_layer = len(rmt.layers) - 1
        for t in range(n_chunks):
            slot = t % cap
            if t > 0:
                # memory evolution: Q = newest state, K/V = every state still cached (incl. itself)
                n = min(t, cap)
                prev = (t - 1) % cap
                q = ring_qkv[:, prev * lq:(prev + 1) * lq, :hd]
                kv = ring_qkv[:, : n * lq]
                ctx, _, _ = ops.xattn(q, kv[..., hd:2 * hd], kv[..., 2 * hd:], heads, head_dim=dhp, scale=scale)
                mem = evo.residual(ctx, mem, weight=evo_p["wo"], fillers=fillers)
            if fill:
                # what may ride in this chunk's GEMM tails: what is left of THIS chunk's frame K/V (needed by the first
                # attention below), then the NEXT chunk's, then the fuser MLP of finished states (up; down once its up
                # is complete)
                fillers = kv_work[t:t + 2] + [w_ for w_ in fuse_work if not w_.done]
            r0, r1 = bounds[t] * p, bounds[t + 1] * p
            if piece_frames is not None and bounds[t] % piece_frames == 0:   # first chunk of a piece: its frames' K/V
                j = bounds[t] // piece_frames
                if before_piece is not None:
                    before_piece(j)
                e0, e1 = bounds[t] * p, min(f, bounds[t] + piece_frames) * p
                ops.linear(z2[0, e0:e1], wf, bf, out=kvf[0, e0:e1])
            for li, layer in enumerate(rmt.layers):
                pk = packs[li]
                att = layer.memory_segment_fusion_attention
                q = lin(mem, pk["wq"], pk["bq"], fillers=fillers)
                if fill and li == 0:
                    kv_work[t].run()                                    # this chunk's frame K/V: whatever the tails left
                kcol = li * 2 * hd
                ctx, _, _ = ops.xattn(q, kvf[:, r0:r1, kcol:kcol + hd], kvf[:, r0:r1, kcol + hd:kcol + 2 * hd], heads,
                                      head_dim=dhp, scale=scale)
                a = att.residual(ctx, mem, weight=pk["wo"], fillers=fillers)
                up = lin(a, layer.mlp[0].weight, layer.mlp[0].bias, act=layer._act, fillers=fillers)
                if li == last_layer and b == 1:                         # the new state is normalised straight into its ring slot
                    mem = layer.residual(up, a, out=ring_states[:, slot], fillers=fillers)
                else:
                    mem = layer.residual(up, a, fillers=fillers)
            if b != 1:
                ring_states[:, slot].copy_(mem)
            if fill and t >= first:
                # the state just written stays in its ring slot to the end (t + cap >= n_chunks): its fuser MLP
                # (llava_arch.py:545-546) becomes filler for the chunks that follow; rows land at their final position
                i_seq = t - first
                npm_ = len(MEMORY_PROMPT_IDS)
                up_w = ops.GemmWork(ring_states[0, slot], fz[0].weight, fz[0].bias, fuse_hidden[i_seq], act=ACT_GELU_ERF)
                dn_w = ops.GemmWork(fuse_hidden[i_seq], fz[2].weight, fz[2].bias,
                                    seq_out[0, npm_ + i_seq * lq: npm_ + (i_seq + 1) * lq],
                                    addvec=self.token_type_embedding.weight.detach()[0], after=up_w)
                fuse_work += [up_w, dn_w]
            if t + 1 < n_chunks:                                        # project the new state once for later chunks
                for bi in range(b):
                    lin(mem[bi], w_evo, b_evo, out=ring_qkv[bi, slot * lq:(slot + 1) * lq],
                        fillers=[w_ for w_ in fuse_work if not w_.done] if fill else ())

        # fuser + assembly: state written at chunk t sits in slot t % cap; reference order is oldest first
        n_fine = min(self.max_fine_frames, f)
        fkey = ("fine", f, str(dev))
        if fkey not in self._consts:                                    # cached: no H2D copy on the hot path / in graphs
            self._consts[fkey] = fine_frame_indices(f, self.max_fine_frames).to(dev)
        fine_idx = self._consts[fkey]
        seq_len = self.sequence_length(n_keep, n_fine, drop_frames)
        if seq_out is None:
            seq_out = torch.empty((b, seq_len, d), dtype=dtype, device=dev)
        pm_ids, pf_ids = self._const_ids(dev)
        emb = self.token_type_embedding.weight.detach()
        npm = len(MEMORY_PROMPT_IDS)
        if fill:
            # whatever the tails did not absorb: the remaining fuser tiles (each launch takes a ready filler along)
            for w_ in fuse_work:
                w_.run(fillers=[o_ for o_ in fuse_work if not o_.done])
            if not frames_assembled:
                ops.assemble(seq_out[0], None, n_keep * lq, z[0], fine_idx, p, emb, self.image_newline.detach(),
                             self.embed_tokens.weight.detach(), pm_ids, pf_ids, drop_frames)
            out = {"sequence": seq_out}
            if return_states:
                if first == 0:
                    out["states"] = ring_states[:, :n_keep]
                else:
                    order = [(first + i) % cap for i in range(n_keep)]
                    out["states"] = torch.cat([ring_states[:, s_:s_ + 1] for s_ in order], dim=1)
            return out
        if isinstance(fz, MemoryFuser) and frames_assembled:
            raise ValueError("mavlm: frames_assembled needs the MLP fuser (the encoder variant's tokens go through the assembly kernel)")
        if isinstance(fz, MemoryFuser):
            # encoder-variant fuser (MemoryFuser.py; the mode llava_arch.py:137-143 keeps commented out): self-attention
            # inside each 196-token memory slot, over the cached states oldest first (llava_arch.py:545-546)
            order = [(first + i) % cap for i in range(n_keep)]
            ordered = ring_states[:, :n_keep] if first == 0 else torch.cat([ring_states[:, s_:s_ + 1] for s_ in order], dim=1)
            tok = fz(ordered.reshape(b * n_keep * (lq // p), p, d)).reshape(b, n_keep * lq, d)
            for bi in range(b):
                ops.assemble(seq_out[bi], tok[bi].contiguous(), n_keep * lq, z[bi], fine_idx, p, emb,
                             self.image_newline.detach(), self.embed_tokens.weight.detach(), pm_ids, pf_ids, drop_frames)
            out = {"sequence": seq_out}
            if return_states:
                out["states"] = ordered
            return out
        hidden = ops.linear(ring_states[:, :n_keep].reshape(b * n_keep * lq, d), fz[0].weight, fz[0].bias,
                            act=ACT_GELU_ERF).reshape(b, n_keep, lq, 4 * d)
        # reference order is oldest state first: ring slots (first+i) % cap -- at most two contiguous runs of
        # slots, each ONE GEMM whose rows land in their final position in the sequence (M = run * Lq rows
        # quantises far better on 148 SMs than per-state GEMMs of 1568 rows)
        runs, i = [], 0
        while i < n_keep:
            s0 = (first + i) % cap
            ln = min(n_keep - i, cap - s0, n_keep - s0) if s0 < n_keep else 0
            if ln <= 0:
                raise RuntimeError("mavlm: inconsistent state ring")    # pragma: no cover
            runs.append((i, s0, ln))
            i += ln
        for bi in range(b):
            for (i0, s0, ln) in runs:
                dst = seq_out[bi, npm + i0 * lq: npm + (i0 + ln) * lq]
                ops.linear(hidden[bi, s0:s0 + ln].reshape(ln * lq, 4 * d), fz[2].weight, fz[2].bias, addvec=emb[0],
                           out=dst)
            if not frames_assembled:
                ops.assemble(seq_out[bi], None, n_keep * lq, z[bi], fine_idx, p, emb, self.image_newline.detach(),
                             self.embed_tokens.weight.detach(), pm_ids, pf_ids, drop_frames)
        out = {"sequence": seq_out}
        if return_states:
            if first == 0:
                out["states"] = ring_states[:, :n_keep]                 # [B, n_keep, Lq, D], oldest first
            else:
                order = [(first + i) % cap for i in range(n_keep)]
                out["states"] = torch.cat([ring_states[:, s_:s_ + 1] for s_ in order], dim=1)
        return out

    @torch.no_grad()
    def assemble_frames(self, seq_out: torch.Tensor, z: torch.Tensor, *, drop_frames: bool = False) -> int:
        """Every row of the assembled sequence [B, L, D] that is NOT a memory token -- both prompts, the newlines and
        the fine frames + token_type_embedding[1] (llava_arch.py:548-554, 613-629) -- from the pooled frames z
        [B, F, P, D] alone (uniform scheduler).  Returns the number of memory-token rows, which
        `memory_forward(..., seq_out=seq_out, frames_assembled=True)` fills in."""
        rmt = self.recurrent_memory_transformer
        b, f, p, d = z.shape
        lq = rmt.num_memory_tokens * p
        n_keep = min(len(uniform_segment_variant(f, self.chunk_size)) - 1, rmt.cache_size)
        self.prepare_constants(f, z.device)
        fine_idx = self._consts[("fine", f, str(z.device))]
        pm_ids, pf_ids = self._const_ids(z.device)
        emb = self.token_type_embedding.weight.detach()
        for bi in range(b):
            ops.assemble(seq_out[bi], None, n_keep * lq, z[bi], fine_idx, p, emb, self.image_newline.detach(),
                         self.embed_tokens.weight.detach(), pm_ids, pf_ids, drop_frames)
        return n_keep * lq

    @torch.no_grad()
    def forward(self, tower_tokens: torch.Tensor, frame_idx: torch.Tensor, *, validate: bool = True,
                **kw) -> Dict[str, torch.Tensor]:
        """tower_tokens [B, F, side*side, Dv], frame_idx [B, F] (original-video indices for the PE)."""
        b, f = tower_tokens.shape[:2]
        z = self.encode_frames(tower_tokens.reshape(b * f, *tower_tokens.shape[2:]), frame_idx.reshape(-1),
                               validate=validate)
        return self.memory_forward(z.reshape(b, f, *z.shape[1:]), **kw)

    def memory_forward_train(self, z: torch.Tensor, *, drop_frames: bool = False) -> Dict[str, torch.Tensor]:
        """Differentiable version of memory_forward for training (BPTT), BATCHED over the videos: every GEMM /
        attention runs on all B videos at once (M = B * 1568 rows instead of B launches of 1568), with the same
        per-video arithmetic as the reference's loop (MemoryController.py:118-158, llava_arch.py:528-557).  Every
        op is an autograd.Function backed by libmavlm.so.  z [B, F, P, D] are the (detached) pooled + PE'd frames;
        gradients reach the recurrent memory transformer, the fuser, token_type_embedding, image_newline and
        embed_tokens (prompt rows).  The state cache is a Python list like the reference's (cap 10, :152-154); the
        evolution K/V of a state are projected once and reused by later chunks (autograd sums their gradients)."""
        rmt = self.recurrent_memory_transformer
        b, f, p, d = z.shape
        dev, dtype = z.device, z.dtype
        z = z.detach()                                                  # llava_arch.py:302,481
        m_slots, lq = rmt.num_memory_tokens, rmt.num_memory_tokens * p
        heads = rmt.layers[0].memory_segment_fusion_attention.num_attention_heads
        scale = 1.0 / math.sqrt(d // heads)
        cap = rmt.cache_size
        pm_ids, pf_ids = self._const_ids(dev)
        self.prepare_constants(f, dev)                                  # cached: no H2D copy inside a captured step
        fine_idx = self._consts[("fine", f, str(dev))]
        emb = self.token_type_embedding.weight
        newline = self.image_newline
        fz = self.memory_fuser
        if isinstance(fz, MemoryFuser):
            raise NotImplementedError("mavlm: the encoder-variant MemoryFuser is inference-only on this path; train "
                                      "with the MLP fuser the reference uses (llava_arch.py:132-136)")
        bounds = uniform_segment_variant(f, self.chunk_size)
        n_chunks = len(bounds) - 1
        z2 = z.reshape(b, f * p, d)

        packs = [l.memory_segment_fusion_attention.packed() for l in rmt.layers]       # part of the autograd graph
        dhp = packs[0]["dhp"]
        hd = heads * dhp
        # frame-side K/V per chunk and layer (one GEMM of B*C*P rows each; the weight gradient accumulates over the
        # chunks in autograd -- row slices of a whole-video projection would cost a zero-filled full-size gradient each)
        evo = rmt.memory_update_attention
        evo_p = evo.packed() if n_chunks > 1 else None

        states: List[torch.Tensor] = []                                 # [B, Lq, D] each
        state_kv: List[torch.Tensor] = []                               # evolution (k | v) of states[i], projected once
        mem = rmt.initial_state(dtype).reshape(1, lq, d).expand(b, lq, d)
        for t in range(n_chunks):
            if t > 0:
                while len(state_kv) < len(states):
                    state_kv.append(ops.linear(states[len(state_kv)], evo_p["wkv"], evo_p["bkv"]))
                kv = state_kv[0] if len(state_kv) == 1 else torch.cat(state_kv, dim=1)
                q = ops.linear(mem, evo_p["wq"], evo_p["bq"])
                ctx = ops.xattn_kv(q, kv, heads, head_dim=dhp, scale=scale)
                mem = evo.residual(ctx, mem, weight=evo_p["wo"])
            r0, r1 = bounds[t] * p, bounds[t + 1] * p
            zc = z2[:, r0:r1] if n_chunks == 1 else z2[:, r0:r1].contiguous()   # one copy per chunk, shared by the layers
            for li, layer in enumerate(rmt.layers):
                pk = packs[li]
                q = ops.linear(mem, pk["wq"], pk["bq"])
                kvf = ops.linear(zc, pk["wkv"], pk["bkv"])
                ctx = ops.xattn_kv(q, kvf, heads, head_dim=dhp, scale=scale)
                a = layer.memory_segment_fusion_attention.residual(ctx, mem, weight=pk["wo"])
                up = ops.linear(a, layer.mlp[0].weight, layer.mlp[0].bias, act=layer._act)
                mem = layer.residual(up, a)
            states.append(mem)
            if len(states) > cap:                                       # MemoryController.py:153-154
                states = states[-cap:]
                state_kv = state_kv[-(cap - 1):] if cap > 1 else []
        cat = states[0] if len(states) == 1 else torch.cat(states, dim=1)               # [B, n*Lq, D]  llava_arch.py:545
        hid = ops.linear(cat, fz[0].weight, fz[0].bias, act=ACT_GELU_ERF)
        memtok = ops.linear(hid, fz[2].weight, fz[2].bias, addvec=emb[0])              # + token_type_embedding[0]
        from .autograd import AssembleFn
        seq = AssembleFn.apply(memtok, z, fine_idx, emb, newline, self.embed_tokens.weight, pm_ids, pf_ids,
                               self.embed_tokens(pm_ids), self.embed_tokens(pf_ids), bool(drop_frames))
        return {"sequence": seq,
                "states": torch.stack([s_.reshape(b, m_slots, p, d) for s_ in states], dim=1)}

    def graphed_train(self, batch: int, frames: int, loss_fn=None) -> "GraphedTrainStep":
        """CUDA-graph replay of one training step (forward + loss + backward) for a fixed (batch, frames)."""
        return GraphedTrainStep(self, batch, frames, loss_fn)

    def graphed(self, batch: int, frames: int, *, return_states: bool = False) -> "GraphedPipeline":
        """CUDA-graph replay of forward() for a fixed (batch, frames): the ~45 dependent launches of a step
        become one graph launch (the recurrence is launch-latency sensitive: ~25 kernels per chunk)."""
        key = ("graph", batch, frames, return_states)
        if key not in self._consts:
            self._consts[key] = GraphedPipeline(self, batch, frames, return_states=return_states)
        return self._consts[key]


class GraphedTrainStep:
    """One TRAINING step of the path -- memory_forward_train, the loss and the whole backward (BPTT through the chunks)
    -- captured into ONE CUDA graph (BASELINE config[3]; the reference runs this step under PyTorch autograd,
    train.py:1694-1728).  The ~600 launches of a step are issued by the autograd engine from Python otherwise, which
    leaves host gaps between kernels.  Static input buffer `z` [B, F, P, D] (pooled + PE'd frames, detached as in
    llava_arch.py:302); gradients land in the parameters' .grad (static tensors, overwritten by every replay);
    `loss_fn(sequence) -> scalar` defaults to the mean square the parity tests use."""

    def __init__(self, pipe: VisualMemoryPipeline, batch: int, frames: int, loss_fn=None, warmup: int = 3):
        rmt = pipe.recurrent_memory_transformer
        p0 = rmt.initial_memory
        dev = p0.device
        dtype = pipe.memory_fuser[0].weight.dtype
        self.pipe, self.dev = pipe, dev
        self.loss_fn = loss_fn or (lambda seq: (seq.float() ** 2).mean())
        self.z = torch.zeros((batch, frames, rmt.patch_size, rmt.hidden_size), dtype=dtype, device=dev)
        self.params = [p_ for p_ in pipe.parameters() if p_.requires_grad]
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._zero()
                    self.loss_fn(pipe.memory_forward_train(self.z)["sequence"]).backward()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self._zero()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                out = pipe.memory_forward_train(self.z)
                self.sequence = out["sequence"]
                self.loss = self.loss_fn(self.sequence)
                self.loss.backward()

    def _zero(self):
        for p_ in self.params:
            p_.grad = None

    def __call__(self, z: Optional[torch.Tensor] = None):
        """Replay on z (copied into the static buffer; None = keep its contents).  Returns (loss, sequence): device
        tensors that the next replay overwrites."""
        with torch.cuda.device(self.dev):
            if z is not None and z is not self.z:
                self.z.copy_(z.reshape(self.z.shape), non_blocking=True)
            self.graph.replay()
        return self.loss, self.sequence


class GraphedPipeline:
    """forward() captured once into a CUDA graph (static input / output buffers owned here)."""

    def __init__(self, pipe: VisualMemoryPipeline, batch: int, frames: int, *, return_states: bool = False):
        self.pipe = pipe
        self.return_states = return_states
        p0 = pipe.mm_projector[0].weight
        dev, dtype = p0.device, p0.dtype
        self.dev = dev
        self.x = torch.zeros((batch, frames, pipe.side * pipe.side, p0.shape[1]), dtype=dtype, device=dev)
        self.idx = torch.zeros((batch, frames), dtype=torch.int64, device=dev)
        self._capture()

    def _weights_key(self):
        """(data_ptr, version) of everything the captured kernels read: the graph bakes in pointers to packed /
        padded copies of the weights, so an in-place update (optimizer step, load_state_dict) or a re-allocated
        parameter after capture must trigger a re-capture instead of replaying stale weights."""
        pipe = self.pipe
        ts = [*pipe.parameters(), *pipe.buffers(), pipe.image_newline]
        return tuple((t.data_ptr(), t._version) for t in ts)

    def _capture(self):
        pipe, dev = self.pipe, self.dev
        with torch.cuda.device(dev):                                    # capture on the weights' device, whatever is current
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                               # warm-up: packs weights, fills constant caches
                for _ in range(2):
                    pipe.forward(self.x, self.idx, validate=False, return_states=self.return_states)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = pipe.forward(self.x, self.idx, validate=False, return_states=self.return_states)
        self._key = self._weights_key()

    @torch.no_grad()
    def __call__(self, tower_tokens: Optional[torch.Tensor], frame_idx: Optional[torch.Tensor]):
        """Copy the inputs into the static buffers (skipped when they ARE the static buffers / None) and
        replay.  The returned tensors are overwritten by the next call (and are NEW tensors after a re-capture,
        which happens when a weight changed since the capture)."""
        with torch.cuda.device(self.dev):
            if self._weights_key() != self._key:
                self._capture()
            if frame_idx is not None and frame_idx is not self.idx:
                self.pipe.positional_encoding.validate(frame_idx)       # host-side, like position_encoding.py:73-76
                self.idx.copy_(frame_idx.reshape(self.idx.shape), non_blocking=True)
            if tower_tokens is not None and tower_tokens is not self.x:
                self.x.copy_(tower_tokens.reshape(self.x.shape), non_blocking=True)
            self.graph.replay()
        return self.out


class _StreamSlot:
    """One of the two buffer sets of HostStreamEncoder: static input / pooled-frame / sequence buffers and the CUDA
    graphs that connect them -- one per input PIECE (projector + pool + PE of that piece's frames) and one for the
    recurrence + fuser -- so that ordinary stream events can sit between them (piece j's graph waits only for piece j's
    H2D copy; the frame rows of the sequence are assembled and on their way back before the recurrence starts)."""

    def __init__(self, pipe: VisualMemoryPipeline, batch: int, frames: int, pieces: int):
        p0 = pipe.mm_projector[0].weight
        dev, dtype = p0.device, p0.dtype
        rmt = pipe.recurrent_memory_transformer
        self.pipe, self.dev, self.batch, self.frames = pipe, dev, batch, frames
        n = batch * frames
        self.x = torch.zeros((n, pipe.side * pipe.side, p0.shape[1]), dtype=dtype, device=dev)
        self.idx = torch.zeros((n,), dtype=torch.int64, device=dev)
        self.z = torch.zeros((n, rmt.patch_size, rmt.hidden_size), dtype=dtype, device=dev)
        n_keep = min(len(uniform_segment_variant(frames, pipe.chunk_size)) - 1, rmt.cache_size)
        self.n_mem_rows = n_keep * rmt.num_memory_tokens * rmt.patch_size
        self.head_rows = len(MEMORY_PROMPT_IDS) + self.n_mem_rows       # memory prompt + memory tokens: known last
        seq_len = pipe.sequence_length(n_keep, min(pipe.max_fine_frames, frames))
        self.seq = torch.zeros((batch, seq_len, rmt.hidden_size), dtype=dtype, device=dev)
        step = -(-n // pieces)
        self.ranges = [(i, min(n, i + step)) for i in range(0, n, step)]
        self.capture()

    def weights_key(self):
        pipe = self.pipe
        ts = [*pipe.parameters(), *pipe.buffers(), pipe.image_newline]
        return tuple((t.data_ptr(), t._version) for t in ts)

    def _encode(self, r0: int, r1: int) -> None:
        self.pipe.encode_frames(self.x[r0:r1], self.idx[r0:r1], validate=False, out=self.z[r0:r1])

    def _z4(self) -> torch.Tensor:
        return self.z.reshape(self.batch, self.frames, *self.z.shape[1:])

    def assemble(self) -> None:                                         # one small kernel per video: launched eagerly
        self.pipe.assemble_frames(self.seq, self._z4())

    def _recur(self) -> None:
        self.pipe.memory_forward(self._z4(), seq_out=self.seq, return_states=False, frames_assembled=True)

    def capture(self) -> None:
        dev = self.dev
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                               # warm-up: packs weights, fills constant caches
                for _ in range(2):
                    for (r0, r1) in self.ranges:
                        self._encode(r0, r1)
                    self.assemble()
                    self._recur()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.g_piece = []
            for (r0, r1) in self.ranges:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._encode(r0, r1)
                self.g_piece.append(g)
            self.g_recur = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_recur):
                self._recur()
        self.key = self.weights_key()


class HostStreamEncoder:
    """Host-buffer front end: pinned host tower tokens in, pinned host sequence out, on three streams (copy-in /
    compute / copy-out) and two buffer sets that take turns, so that the H2D copy of video i+1 and the D2H copy of
    result i-1 overlap the compute of video i.  Inside one video the copies are pipelined as well: the input arrives in
    `pieces` pieces of frames and each piece's projector graph waits only for its own piece (the first kernel starts
    after half of the 107 MB, not all of it), and the rows of the sequence that depend on the frames alone -- prompts,
    newlines, fine frames: two thirds of it -- are assembled right after the projector and copied back WHILE the
    recurrence runs; only the memory-token rows leave after the last kernel.  Same kernels, same results bit for bit as
    `pipe(...)`.  Measured on B200 (tools/e2e_gap.py, OV-7B, 64 frames): one video host to host 7.4 ms serial -> 6.1 ms
    with 4 pieces; in a stream of videos 4 pieces cost the projector 0.14 ms per video (its second GEMM runs on 3136
    rows: 2.5 waves) and save 1.4 ms once per stream, 2 pieces cost 0.01 ms and save 0.95 ms: hence the default of 2.
    This is the public end-to-end call bench.py's `e2e` times."""

    def __init__(self, pipe: VisualMemoryPipeline, batch: int, frames: int, pieces: int = 2):
        if isinstance(pipe.memory_fuser, MemoryFuser):
            raise ValueError("mavlm: HostStreamEncoder needs the MLP fuser (llava_arch.py:132-136)")
        n = batch * frames
        pieces = max(1, min(int(pieces), n))
        self.slots = [_StreamSlot(pipe, batch, frames, pieces) for _ in range(2)]
        dev = self.slots[0].dev
        self.dev = dev
        self.pipe = pipe
        self.s_in, self.s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.ev_piece = [[torch.cuda.Event() for _ in sl.ranges] for sl in self.slots]
        self.ev_x_free = [torch.cuda.Event(), torch.cuda.Event()]       # the projector graphs have read the slot's input
        self.ev_frames = [torch.cuda.Event(), torch.cuda.Event()]       # frame rows of the sequence are assembled
        self.ev_done = [torch.cuda.Event(), torch.cuda.Event()]         # memory-token rows written: sequence complete
        self.ev_out_free = [torch.cuda.Event(), torch.cuda.Event()]     # both D2H copies of the slot's sequence finished
        self._used = [False, False]
        self._k = 0
        self._idx_set = False

    @torch.no_grad()
    def submit(self, host_tokens: torch.Tensor, frame_idx: Optional[torch.Tensor], host_out: torch.Tensor) -> None:
        """Enqueue one batch: host_tokens (pinned) -> device -> path -> host_out (pinned, [B, L, D]).  Asynchronous;
        call synchronize() before reading host_out.  frame_idx None = keep the indices of the previous call
        (initially 0 .. F-1 per video)."""
        with torch.cuda.device(self.dev):
            self._submit(host_tokens, frame_idx, host_out)

    def _submit(self, host_tokens: torch.Tensor, frame_idx: Optional[torch.Tensor], host_out: torch.Tensor) -> None:
        k = self._k
        sl = self.slots[k]
        cur = torch.cuda.current_stream(self.dev)
        if sl.weights_key() != sl.key:                                  # a weight changed since the capture
            torch.cuda.synchronize(self.dev)
            for s_ in self.slots:
                s_.capture()
        if frame_idx is not None:                                       # new indices go to both slots (stream-ordered
            self.pipe.positional_encoding.validate(frame_idx)           # on the compute stream, like every replay)
            for s_ in self.slots:
                s_.idx.copy_(frame_idx.reshape(s_.idx.shape), non_blocking=True)
            self._idx_set = True
        elif not self._idx_set:
            per_video = torch.arange(sl.frames, dtype=torch.int64).repeat(sl.batch)
            for s_ in self.slots:
                s_.idx.copy_(per_video, non_blocking=True)
            self._idx_set = True
        src = host_tokens.reshape(sl.x.shape)
        with torch.cuda.stream(self.s_in):
            if self._used[k]:
                self.s_in.wait_event(self.ev_x_free[k])                 # the slot's previous projector pass has read its input
            for j, (r0, r1) in enumerate(sl.ranges):
                sl.x[r0:r1].copy_(src[r0:r1], non_blocking=True)
                self.ev_piece[k][j].record(self.s_in)
        if self._used[k]:
            cur.wait_event(self.ev_out_free[k])                         # ... and its previous sequence has left
        for j, g in enumerate(sl.g_piece):
            cur.wait_event(self.ev_piece[k][j])
            g.replay()
        self.ev_x_free[k].record(cur)
        sl.assemble()
        self.ev_frames[k].record(cur)
        sl.g_recur.replay()
        self.ev_done[k].record(cur)
        dst = host_out.reshape(sl.seq.shape)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_frames[k])
            if sl.batch == 1:                                           # contiguous row ranges: two plain copies
                dst[0, sl.head_rows:].copy_(sl.seq[0, sl.head_rows:], non_blocking=True)
                self.s_out.wait_event(self.ev_done[k])
                dst[0, :sl.head_rows].copy_(sl.seq[0, :sl.head_rows], non_blocking=True)
            else:
                self.s_out.wait_event(self.ev_done[k])
                dst.copy_(sl.seq, non_blocking=True)
            self.ev_out_free[k].record(self.s_out)
        self._used[k] = True
        self._k = k ^ 1

    def synchronize(self) -> None:
        self.s_out.synchronize()
        torch.cuda.current_stream(self.dev).synchronize()
