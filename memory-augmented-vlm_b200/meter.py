"""Per-kernel device timing of a captured step (measurement plumbing, used by bench.py and tools/sweep.py).

The step is captured into a CUDA graph with an external timing event recorded before and after every library call
(event-record nodes on the capture stream), so the intervals are device times of back-to-back kernels exactly as a
replayed step runs them: no host launch gaps, no allocator stalls.  Every call is metered with its ALGORITHMIC work
(2 M N K flop for a GEMM, 4 B Lq Lk H dh for attention) or its I/O bytes (LayerNorm, pool + PE, assembly), which is
what the roofline fractions are computed from.
"""
from __future__ import annotations

import statistics
from typing import Callable, Dict, List

import torch

from . import ops

KERNEL_NAMES = {"gemm": "gemm_tc_kernel (tcgen05 bf16 GEMM, all epilogues)",
                "xattn": "attn_tc_kernel (tcgen05 flash cross-attention)",
                "layernorm": "layernorm_kernel", "pool_pe": "pool_pe_kernel", "add_pe": "add_pe_kernel",
                "assemble": "assemble_kernel"}


def _nbytes(t):
    return 0 if t is None else t.numel() * t.element_size()


def _m_linear(y, x, w, b=None, **kw):
    m = x.numel() // x.shape[-1]
    return "gemm", 2.0 * m * w.shape[0] * w.shape[1], 0, f"{m}x{w.shape[0]}x{w.shape[1]}"


def _m_linear_pe(y, x, w, b, table, fidx, **kw):
    m = x.numel() // x.shape[-1]
    return "gemm", 2.0 * m * w.shape[0] * w.shape[1], 0, f"{m}x{w.shape[0]}x{w.shape[1]}"


def _m_xattn(y, q, k, v, heads, **kw):
    bq, lq, hd = q.shape
    lk = k.shape[1]
    return "xattn", 4.0 * bq * lq * lk * hd, 0, f"B{bq} Lq{lq} Lk{lk} dh{kw.get('head_dim') or hd // heads}"


def _m_layernorm(y, x, *a, **kw):
    out = kw.get("out")
    return ("layernorm", 0.0, _nbytes(x) + (_nbytes(out) if out is not None else _nbytes(y)),
            f"{x.numel() // x.shape[-1]}x{x.shape[-1]}")


def _m_pool(y, x, **kw):
    return "pool_pe", 0.0, _nbytes(x) + _nbytes(y), f"{x.shape[0]} frames"


def _m_add_pe(y, x, *a, **kw):
    return "add_pe", 0.0, 2 * _nbytes(x), ""


def _m_assemble(y, seq, mem, n_mem_rows, frames, fine_idx, tokens, *a, **kw):
    rows = seq.shape[0] - (0 if mem is not None else n_mem_rows)          # rows this launch writes (each read once too)
    return "assemble", 0.0, 2 * rows * seq.shape[-1] * seq.element_size(), f"{rows} rows"


_METERS = {"linear": _m_linear, "linear_pe": _m_linear_pe, "xattn": _m_xattn, "layernorm": _m_layernorm,
           "pool_pe": _m_pool, "add_pe": _m_add_pe, "add_rows": _m_add_pe, "assemble": _m_assemble}


class KernelMeter:
    """with KernelMeter() as km: graph = <capture the step>;  then km.collect(replay, n) and km.summary(...)."""

    def __init__(self):
        self.recs: List = []                        # (family, flops, bytes, e0, e1, tag)
        self._orig: Dict[str, Callable] = {}
        self.rec_ms: List[float] = []
        self.step_ms = 0.0
        self.replays = 0
        self._null = None                           # (e0, e1) around an 8-element cast: the per-launch floor
        self.launch_floor_ms = 0.0

    def __enter__(self):
        for name, meter in _METERS.items():
            self._orig[name] = getattr(ops, name)
            setattr(ops, name, self._wrap(self._orig[name], meter))
        return self

    def __exit__(self, *exc):
        for name, fn in self._orig.items():
            setattr(ops, name, fn)
        return False

    def _wrap(self, fn0, meter):
        def fn(*a, **kw):
            if not torch.cuda.is_current_stream_capturing():
                return fn0(*a, **kw)
            if self._null is None:
                # calibration: the same event -> launch -> event pattern around a kernel with nothing to do (8 elements).
                # Its duration is what the instrumentation adds to EVERY metered launch (launch latency that a plain
                # graph overlaps with the previous kernel's tail, plus the event nodes)
                dev = next(t.device for t in list(a) + list(kw.values()) if isinstance(t, torch.Tensor))
                src = torch.zeros(8, device=dev, dtype=torch.float32)
                n0 = torch.cuda.Event(enable_timing=True, external=True)
                n1 = torch.cuda.Event(enable_timing=True, external=True)
                n0.record()
                self._null_out = ops.cast(src, torch.bfloat16)
                n1.record()
                self._null = (n0, n1)
            e0 = torch.cuda.Event(enable_timing=True, external=True)
            e1 = torch.cuda.Event(enable_timing=True, external=True)
            e0.record()
            y = fn0(*a, **kw)
            e1.record()
            fam, flops, nbytes, tag = meter(y, *a, **kw)
            self.recs.append((fam, flops, nbytes, e0, e1, tag))
            return y
        return fn

    def collect(self, replay: Callable[[], None], n: int, burst: int = 0, pause_s: float = 0.0, on_burst=None) -> None:
        """Replay the metered graph n times; keep the MEDIAN duration of every launch and of the whole step.
        `burst` > 0: the replays run in bursts of that many with `pause_s` seconds of idle time in between, so that a
        pass of >= 100 replays stays in the clock regime of a short timed region instead of running into the power cap
        half-way; `on_burst()` is called at the end of every burst (e.g. to read the SM clock while still loaded)."""
        import time
        per = [[] for _ in self.recs]
        steps = []
        floor = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(n):
            if burst > 0 and it > 0 and it % burst == 0 and pause_s > 0:
                time.sleep(pause_s)
            e0.record()
            replay()
            e1.record()
            if on_burst is not None and burst > 0 and (it + 1) % burst == 0:
                on_burst()                                              # the burst's last replay is still running
            torch.cuda.synchronize()
            steps.append(e0.elapsed_time(e1))
            for i, (_, _, _, a, b, _) in enumerate(self.recs):
                per[i].append(a.elapsed_time(b))
            if self._null is not None:
                floor.append(self._null[0].elapsed_time(self._null[1]))
        self.rec_ms = [statistics.median(v) if v else 0.0 for v in per]
        self.step_ms = statistics.median(steps) if steps else 0.0
        self.launch_floor_ms = statistics.median(floor) if floor else 0.0
        self.replays = n

    def families(self) -> Dict[str, Dict[str, float]]:
        fams: Dict[str, Dict[str, float]] = {}
        for (fam, flops, nbytes, _, _, _), ms in zip(self.recs, self.rec_ms):
            f = fams.setdefault(fam, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            f["launches"] += 1
            f["ms"] += ms
            f["flops"] += flops
            f["bytes"] += nbytes
        return fams

    def roofline_kernels(self, tensor_peak: float, peaks: Dict[str, float]) -> List[Dict]:
        """One entry per kernel family: achieved / peak / frac against the tensor peak in force (burst or sustained)
        or the measured HBM copy bandwidth, plus both tensor fractions."""
        out = []
        for fam, f in self.families().items():
            if f["ms"] <= 0:
                continue
            ent = {"kernel": KERNEL_NAMES.get(fam, fam), "launches_per_step": f["launches"], "ms_per_step": f["ms"],
                   "share_of_step": f["ms"] / self.step_ms if self.step_ms > 0 else None}
            if f["flops"] > 0:
                ach = f["flops"] / (f["ms"] * 1e-3) / 1e12
                ent.update({"bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s",
                            "frac": ach / tensor_peak, "frac_of_burst": ach / peaks["tflops_burst"],
                            "frac_of_sustained": ach / peaks["tflops_sustained"], "gflop_per_step": f["flops"] / 1e9})
            else:
                ach = f["bytes"] / (f["ms"] * 1e-3) / 1e9
                ent.update({"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": ach / peaks["hbm_gbs"], "mbytes_per_step": f["bytes"] / 1e6})
            net_ms = f["ms"] - f["launches"] * self.launch_floor_ms
            if self.launch_floor_ms > 0 and net_ms > 0:
                # the same rate with the instrumentation's per-launch floor (KernelMeter.launch_floor_ms) taken out:
                # what the kernel does between its first and last instruction, as ncu's gpu__time_duration sees it
                ent["ms_per_step_net_of_launch_floor"] = net_ms
                ent["frac_net_of_launch_floor"] = ent["frac"] * f["ms"] / net_ms
            out.append(ent)
        out.sort(key=lambda e: -e["ms_per_step"])
        return out

    def shapes(self, family: str) -> Dict[str, Dict[str, float]]:
        by = {}
        for (fam, flops, _, _, _, tag), ms in zip(self.recs, self.rec_ms):
            if fam != family:
                continue
            t = by.setdefault(tag, [0, 0.0, flops])
            t[0] += 1
            t[1] += ms
        return {tag: {"launches_per_step": c, "us": 1e3 * ms / c, "tflops": fl / (ms / c * 1e-3) / 1e12 if ms > 0 else 0.0}
                for tag, (c, ms, fl) in by.items()}
