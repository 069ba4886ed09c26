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


@torch.no_grad()
def encode_long_video_frame_sharded(pipe, tower_tokens: torch.Tensor, frame_idx: torch.Tensor):
    """One long video [F, 729, Dv]: each rank projects + pools + PEs F/world frames, the pooled tokens
    are all-gathered (1.4 MB per 7B frame), then every rank runs the (replicated) recurrence."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    f = tower_tokens.shape[0]
    mine = shard_range(f, rank, world)
    dev = next(pipe.parameters()).device
    z_local = pipe.encode_frames(tower_tokens[mine.start:mine.stop].to(dev, non_blocking=True),
                                 frame_idx[mine.start:mine.stop])
    counts = [len(shard_range(f, r, world)) for r in range(world)]
    z = all_gather_rows(z_local, counts)
    return pipe.memory_forward(z[None])
