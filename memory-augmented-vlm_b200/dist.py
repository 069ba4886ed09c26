"""Multi-GPU host logic: one process per GPU (torchrun), videos sharded over ranks.

The recurrence is sequential in time, so the only axes that shard are (i) independent videos --
each rank runs whole recurrences, weights replicated, NO data-path collective -- and (ii) the
memory-independent pre-pass of one long video (projector + pool + PE per frame), whose pooled tokens
are all-gathered once before the replicated recurrence.  NCCL (NVLink 5 / NVSwitch) is
used only for those gathers; `gloo` runs the same code on CPU for the world_size-2 tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous balanced split: the first n % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def all_gather_rows(local: torch.Tensor, counts: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Concatenate per-rank tensors along dim 0 (rank order).  Equal counts use one
    all_gather_into_tensor (NCCL ring / NVLS); ragged counts pad to the max and trim."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None:
        counts = [local.shape[0]] * world
    mx = max(counts)
    if all(c == mx for c in counts):
        out = local.new_empty((world * mx, *local.shape[1:]))
        if dist.get_backend() == "nccl":
            dist.all_gather_into_tensor(out, local.contiguous())
        else:
            parts = list(out.chunk(world, dim=0))
            dist.all_gather(parts, local.contiguous())
        return out
    padded = local.new_zeros((mx, *local.shape[1:]))
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def max_over_ranks(value: float, device=None) -> float:
    """Multi-GPU timings are reported as the max over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs, via the GPU's PCI bus id), so that the
    pinned host buffers it allocates afterwards are first-touched on that node.  With 8 ranks streaming 175 MB per
    step over PCIe each, buffers that all land on one socket make the host memory controller the bottleneck.
    Returns the node, or None when the topology cannot be read (then nothing is changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                      # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        base = f"/sys/bus/pci/devices/{bus}"
        with open(f"{base}/numa_node") as fh:
            node = int(fh.read().strip())
        with open(f"{base}/local_cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & os.sched_getaffinity(0)
        if node < 0 or not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:                                        # no nvml / sysfs (containers): keep the default placement
        return None


def synced_dropout_decision(prob: float = 0.5, device=None) -> bool:
    """One coin flip shared by all ranks (LlavaMetaForCausalLM.get_synced_dropout_decision, llava_arch.py:378-386):
    rank 0 draws, everybody receives the broadcast.  With `dropout_frames` (stage-1 training, finetune_short.sh:101)
    a True decision drops the fine-frame half of the video sequence (llava_arch.py:719-725); every rank must take
    the same branch or the sequence lengths -- and the collectives of the LLM step -- diverge."""
    if not dist.is_initialized():
        return bool(torch.rand(1).item() < prob)
    t = torch.zeros(1, device=device)
    if dist.get_rank() == 0:
        t.fill_(1.0 if torch.rand(1).item() < prob else 0.0)
    dist.broadcast(t, src=0)
    return bool(t.item())


@torch.no_grad()
def encode_videos_sharded(pipe, tower_tokens: torch.Tensor, frame_idx: torch.Tensor, *, gather: bool = True):
    """tower_tokens [V, F, 729, Dv] (same on every rank, or only the local shard is touched):
    rank r runs videos shard_range(V, r, world); returns assembled sequences [V, L, D] on every rank
    (gather=True) or the local shard."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    v = tower_tokens.shape[0]
    mine = shard_range(v, rank, world)
    dev = next(pipe.parameters()).device
    if len(mine):
        res = pipe(tower_tokens[mine.start:mine.stop].to(dev, non_blocking=True), frame_idx[mine.start:mine.stop],
                   return_states=False)["sequence"]
    else:
        res = None
    if not gather or world == 1:
        return res
    counts = [len(shard_range(v, r, world)) for r in range(world)]
    if res is None:
        n_chunks = -(-tower_tokens.shape[1] // pipe.chunk_size)
        seq_len = pipe.sequence_length(min(n_chunks, pipe.recurrent_memory_transformer.cache_size),
                                       min(pipe.max_fine_frames, tower_tokens.shape[1]))
        d = pipe.recurrent_memory_transformer.hidden_size
        res = torch.empty((0, seq_len, d), dtype=next(pipe.parameters()).dtype, device=dev)
    return all_gather_rows(res, counts)


def piece_schedule(n_frames: int, chunk: int, world: int):
    """Ownership of the frame-sharded pre-pass of ONE long video.  The video is cut into pieces of `world` chunks
    (`world * chunk` frames); inside piece j, rank r owns chunk j * world + r, i.e. frames
    [(j * world + r) * chunk, ... + chunk) clipped to the video.  So every all-gather collects ONE chunk from every
    rank -- piece j is complete after gather j, and the recurrence (sequential in time) can start on piece 0 while the
    later pieces are still being projected and gathered.  A contiguous split (rank r owns frames r*F/W ...) would make
    chunk 0 a broadcast from rank 0 and leave nothing to overlap.
    Returns (piece_frames, [[(start, stop) for rank in range(world)] for piece in range(n_pieces)])."""
    piece = world * chunk
    n_pieces = -(-n_frames // piece)
    sched = []
    for j in range(n_pieces):
        row = []
        for r in range(world):
            s0 = min(n_frames, (j * world + r) * chunk)
            row.append((s0, min(n_frames, s0 + chunk)))
        sched.append(row)
    return piece, sched


def gather_piece(piece_buf: torch.Tensor, slot: torch.Tensor) -> None:
    """All-gather one piece IN PLACE: `slot` is this rank's block of `piece_buf` ([world * n, ...]).  NCCL: one
    ncclAllGather whose send buffer is the receive buffer's rank-th block; gloo (CPU tests): list form."""
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(piece_buf, slot)
    else:
        parts = list(piece_buf.chunk(dist.get_world_size(), dim=0))
        dist.all_gather(parts, slot.clone())


class FrameShardedEncoder:
    """One long video [F, 729, Dv] on `world` GPUs: every rank projects + pools + PEs the chunks it owns
    (piece_schedule), the W2 GEMM's epilogue writes them straight into this rank's slot of the gather buffer, the pooled
    tokens (1.4 MB per OV-7B frame) are all-gathered piece by piece with NCCL on a SIDE stream, and every rank runs the
    replicated recurrence, whose first pieces overlap the projection / gather of the later ones.  The recurrence does
    not shard (state t needs state t-1)."""

    def __init__(self, pipe, n_frames: int):
        self.pipe = pipe
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.n_frames = n_frames
        self.piece, self.sched = piece_schedule(n_frames, pipe.chunk_size, self.world)
        p0 = pipe.mm_projector[0].weight
        self.dev = p0.device
        rmt = pipe.recurrent_memory_transformer
        n_pad = len(self.sched) * self.piece                          # whole pieces: equal-sized gathers
        self.z = torch.zeros((n_pad, rmt.patch_size, rmt.hidden_size), dtype=p0.dtype, device=self.dev)
        self.comm = torch.cuda.Stream(device=self.dev) if self.dev.type == "cuda" else None
        self.gather_events = []
        self.gathered_bytes = 0

    @torch.no_grad()
    def __call__(self, tower_tokens: torch.Tensor, frame_idx: torch.Tensor, *, overlap: bool = True, comm: bool = True,
                 return_states: bool = False, time_gathers: bool = False):
        """tower_tokens [F, 729, Dv] (device; only this rank's chunks are read), frame_idx [F] (host or device).
        overlap=False: every gather is waited for before the recurrence starts (the blocking baseline);
        comm=False: the gathers are skipped (the buffer keeps the previous call's data): times the compute alone."""
        pipe, chunk, r, w = self.pipe, self.pipe.chunk_size, self.rank, self.world
        pipe.positional_encoding.validate(frame_idx)
        frame_idx = frame_idx.to(self.dev)
        cur = torch.cuda.current_stream(self.dev)
        done_pre, done_gather = [], []
        self.gather_events = []
        self.gathered_bytes = 0
        for j, row in enumerate(self.sched):
            s0, s1 = row[r]
            slot = self.z[j * self.piece + r * chunk: j * self.piece + (r + 1) * chunk]
            if s1 > s0:
                pipe.encode_frames(tower_tokens[s0:s1], frame_idx[s0:s1], validate=False, out=slot[: s1 - s0])
            ev = torch.cuda.Event()
            ev.record(cur)
            done_pre.append(ev)
            if w > 1 and comm:
                with torch.cuda.stream(self.comm):
                    self.comm.wait_event(ev)
                    piece_buf = self.z[j * self.piece:(j + 1) * self.piece]
                    if time_gathers:
                        t0 = torch.cuda.Event(enable_timing=True)
                        t0.record(self.comm)
                    gather_piece(piece_buf, slot)                      # in place: slot IS piece_buf's rank-th block
                    ge = torch.cuda.Event(enable_timing=time_gathers)
                    ge.record(self.comm)
                    if time_gathers:
                        self.gather_events.append((t0, ge))
                    done_gather.append(ge)
                    self.gathered_bytes += (w - 1) * slot.numel() * slot.element_size()
        if w > 1 and comm and not overlap:
            for ge in done_gather:
                cur.wait_event(ge)
            done_gather = []

        def before_piece(j):
            if j < len(done_gather):
                cur.wait_event(done_gather[j])

        z = self.z[: self.n_frames][None]
        return pipe.memory_forward(z, piece_frames=self.piece, before_piece=before_piece, return_states=return_states)

    def gather_ms(self) -> float:
        """Sum of the gathers' durations on the comm stream (after a synchronised call with time_gathers=True)."""
        return float(sum(a.elapsed_time(b) for a, b in self.gather_events))


@torch.no_grad()
def encode_long_video_frame_sharded(pipe, tower_tokens: torch.Tensor, frame_idx: torch.Tensor, **kw):
    """One long video [F, 729, Dv]: frame-sharded pre-pass + piece-wise all-gather overlapped with the replicated
    recurrence (FrameShardedEncoder).  Single process: the plain path."""
    dev = next(pipe.parameters()).device
    enc = FrameShardedEncoder(pipe, tower_tokens.shape[0])
    return enc(tower_tokens.to(dev, non_blocking=True), frame_idx, **kw)
