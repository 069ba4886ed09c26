#!/usr/bin/env python
"""Headline benchmark of the visual-memory path (BASELINE.json: frames/sec, OV-7B, 196 tok/frame).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one video per GPU: BASELINE config[1] (OV-7B dims, bf16,
1 video x 64 sampled frames x 729 SigLIP tokens): mm_projector -> bilinear pool + PE -> recurrent
memory (2 chunks of 32 frames: formation x2, evolution x1) -> fuser -> token assembly.
N > 1: one process per GPU, each rank owns one video (weak scaling, no data-path collective; the
recurrence is sequential in time so videos are the only sharding axis).

Prints ONE JSON line on rank 0.  `value` is device-resident throughput, `e2e` is the same metric
through the public API with pinned HOST buffers (H2D of the tower tokens and D2H of the assembled
sequence inside the timed region), `roofline` is the dominant kernel (tcgen05 GEMM) measured live
with CUDA events in an instrumented pass, `cpu_baseline` is the numpy oracle (port of the reference's
CPU path) on the host cores.  --impl reference times that oracle as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec (memory update + fuser, OV-7B, 196 tok/frame)"
HIDDEN, VISION, FRAMES, CHUNK = 3584, 1152, 64, 32


def algorithmic_gflop(frames: int, chunk: int, d: int = HIDDEN, dv: int = VISION, lq: int = 1568, p: int = 196,
                      depth: int = 2, cap: int = 10) -> float:
    """SURVEY.md §8d formulas (MAC = 2 flops; no recomputation; evolution K/V cached; pool-after projector)."""
    n_chunks = -(-frames // chunk)
    fl = 2.0 * 729 * frames * (dv * d + d * d)                                  # projector as the reference computes it
    for t in range(n_chunks):
        c = min(chunk, frames - t * chunk)
        fl += depth * (4.0 * lq * d * d + 4.0 * c * p * d * d + 4.0 * lq * c * p * d + 16.0 * lq * d * d)
        if t > 0:
            n = min(t, cap)
            fl += 8.0 * lq * d * d + 4.0 * lq * (n * lq) * d
    fl += 16.0 * min(n_chunks, cap) * lq * d * d                                # fuser
    return fl / 1e9


_REAL_STDOUT = None


def isolate_stdout():
    """stdout carries exactly ONE JSON line (the driver parses it).  Libraries write there too -- NCCL prints its
    "NCCL version ..." banner to fd 1 at communicator creation -- so point fd 1 at stderr for the run and keep the
    real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_line(text, flush=True):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(text + "\n")
    out.flush()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1590.0))),
                "hbm_gbs": float(p.get("hbm_gbs", 6650.0)), "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"tflops": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.lines = []
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(device_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is not None:
            time.sleep(0.08)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.lines:
            if self.t0 is None or not (self.t0 - 0.03 <= t <= (self.t1 or t) + 0.08):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def oracle_pass(frames: int, chunk: int, weights32, x32, pe_table):
    """One pass of the numpy oracle (fp32, OpenBLAS threads = host cores) over `frames` frames."""
    import numpy as np
    from oracle import vismem_oracle as O
    idx = np.arange(frames)
    t0 = time.perf_counter()
    res = O.visual_memory_path(x32[:frames], idx, weights32, pe_table=pe_table,
                               prompt_mem=weights32["embed_tokens.weight"][list(O.MEMORY_PROMPT_IDS)],
                               prompt_frm=weights32["embed_tokens.weight"][list(O.FRAME_PROMPT_IDS)], chunk=chunk)
    dt = time.perf_counter() - t0
    return dt, res


def host_weights_fp32(seed=0):
    """Same random-init weights as the GPU arm, as fp32 numpy (built on the CPU; no GPU needed)."""
    import numpy as np
    from mavlm_b200 import synthetic
    _, w = synthetic.build_pipeline(HIDDEN, VISION, dtype=__import__("torch").float32, seed=seed, device="cpu",
                                    vocab=50000)
    return {k: v.astype(np.float32) for k, v in w.items()}


def use_all_host_cores() -> None:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host core."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n, user_api="blas")
        threadpool_limits(limits=n, user_api="openmp")
    except Exception:
        pass


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 0) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    use_all_host_cores()
    steps = args.steps if args.steps is not None else 3
    warmup = args.warmup if args.warmup is not None else 1
    w32 = host_weights_fp32()
    from mavlm_b200 import synthetic
    import torch
    x32 = synthetic.synthetic_tower_tokens(1, CHUNK, dtype=torch.float32)[0].numpy()
    pe = w32["positional_encoding.frame_embed"]
    # bounded sample: one full 32-frame chunk per step; shrink if K+W would take more than ~4 minutes
    frames = CHUNK
    t_probe, _ = oracle_pass(2, CHUNK, w32, x32, pe)          # also warms the BLAS thread pool
    est_full = t_probe * 6.0                                   # 2-frame pass ~ 1/6 of a 32-frame pass (per-chunk part dominates)
    budget = 240.0 / max(1, steps + warmup)
    while frames > 1 and est_full * (0.45 + 0.55 * frames / CHUNK) > budget:
        frames //= 2
    for _ in range(warmup):
        oracle_pass(frames, CHUNK, w32, x32, pe)
    t = 0.0
    for _ in range(steps):
        dt, _ = oracle_pass(frames, CHUNK, w32, x32, pe)
        t += dt
    fps = frames * steps / t
    cores = blas_threads()
    sample = (f"numpy/OpenBLAS fp32 port of the reference path, 1 video x {frames} frames (one chunk of the OV-7B "
              f"workload incl. projector/pool/PE/fuser/assembly) per step, {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(int(os.environ.get("WORLD_SIZE", "1"))), "gpu_launches": 0,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(json.dumps(line), flush=True)


def workload_config(world):
    return {"workload": "OV-7B dims (D=3584, SigLIP 729x1152 tokens/frame -> 196 pooled), 1 video x 64 frames per GPU, "
                        "chunks of 32: mm_projector + bilinear pool + PE + recurrent memory (formation x2, evolution "
                        "x1) + fuser MLP + token assembly",
            "frames_per_video": FRAMES, "videos_per_gpu": 1, "chunk": CHUNK, "hidden": HIDDEN, "parallelism":
            f"videos sharded over {world} GPU(s), weights replicated, no data-path collective",
            "l2_policy": "working set per step (0.97 GB weights + 107 MB input) exceeds the 126 MB L2; no flush needed"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import mavlm_b200 as M
    from mavlm_b200 import _lib, ops, synthetic

    steps = args.steps if args.steps is not None else 50
    warmup = max(3, args.warmup if args.warmup is not None else 5)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    numa = M.dist.bind_to_gpu_numa_node(local) if world > 1 else None      # before any pinned buffer is allocated

    pipe, weights = synthetic.build_pipeline(HIDDEN, VISION, dtype=torch.bfloat16, chunk_size=CHUNK, device=dev)
    x_host = synthetic.synthetic_tower_tokens(1, FRAMES, seed=1234 + rank, pin=True)
    x_dev = x_host.to(dev)
    idx = torch.arange(FRAMES)[None]
    n_chunks = -(-FRAMES // CHUNK)
    seq_len = pipe.sequence_length(min(n_chunks, 10), min(32, FRAMES))
    out_host = torch.empty((1, seq_len, HIDDEN), dtype=torch.bfloat16, pin_memory=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        if after is not None:                               # drain side streams: e1 must follow the last D2H
            after()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    graphed = pipe.graphed(1, FRAMES)                       # CUDA-graph replay of the whole step
    graphed(x_dev, idx)                                     # loads the static input / index buffers once
    streamer = M.HostStreamEncoder(pipe, 1, FRAMES)

    def step_device():                                      # inputs resident in HBM (the graph's static buffer)
        graphed(None, None)

    def step_e2e():                                         # pinned host in -> pinned host out, copies overlapped
        streamer.submit(x_host, None, out_host)

    torch.cuda.synchronize()
    l0 = lib.mavlm_launch_count()
    pipe(x_dev, idx, return_states=False)                   # one eager step: counts the kernels a step launches
    torch.cuda.synchronize()
    launches_per_step = lib.mavlm_launch_count() - l0

    for _ in range(warmup):
        step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.15)
    if sampler:
        sampler.mark_start()
    ms_total = timed(step_device, steps)
    if sampler:
        sampler.mark_end()
    launches = launches_per_step * steps                    # kernels replayed inside the timed region
    clocks = sampler.stop() if sampler else None

    if args.timed_only:
        if rank == 0:
            emit_line(json.dumps({"metric": METRIC, "value": world * FRAMES * steps / (ms_total * 1e-3), "unit": "frames/s",
                              "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps,
                              "note": "--timed-only (profiling run)"}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    for _ in range(3):
        step_e2e()
    streamer.synchronize()
    ms_e2e = timed(lambda: step_e2e(), steps, after=streamer.synchronize)

    def step_eager():
        pipe(x_dev, idx, return_states=False)

    # ---- instrumented pass: per-launch CUDA-event timing of every op, INSIDE a CUDA graph of the step ----
    # The step is captured a second time with an external timing event recorded before and after every library
    # call (event-record nodes on the capture stream), so the intervals are device times of back-to-back
    # kernels exactly as the timed graph replays run them: no host launch gaps, no allocator stalls.
    recs = []
    other = {}
    orig = {name: getattr(ops, name) for name in ("linear", "xattn", "layernorm", "pool_pe", "add_pe", "add_rows",
                                                  "assemble")}

    def _events():
        return (torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True))

    def timed_linear(x, w, b=None, **kw):
        if not torch.cuda.is_current_stream_capturing():
            return orig["linear"](x, w, b, **kw)
        e0, e1 = _events()
        e0.record()
        y = orig["linear"](x, w, b, **kw)
        e1.record()
        recs.append((x.numel() // x.shape[-1], w.shape[0], w.shape[1], e0, e1))
        return y

    def timed_other(name):
        def fn(*a, **kw):
            if not torch.cuda.is_current_stream_capturing():
                return orig[name](*a, **kw)
            e0, e1 = _events()
            e0.record()
            y = orig[name](*a, **kw)
            e1.record()
            other.setdefault(name, []).append((e0, e1))
            return y
        return fn

    orig_linear_pe = ops.linear_pe

    def timed_linear_pe(x, w, b, table, fidx):                  # projector W2 + fused PE: a GEMM launch like the others
        if not torch.cuda.is_current_stream_capturing():
            return orig_linear_pe(x, w, b, table, fidx)
        e0, e1 = _events()
        e0.record()
        y = orig_linear_pe(x, w, b, table, fidx)
        e1.record()
        recs.append((x.numel() // x.shape[-1], w.shape[0], w.shape[1], e0, e1))
        return y

    ops.linear = timed_linear
    ops.linear_pe = timed_linear_pe
    for name in orig:
        if name != "linear":
            setattr(ops, name, timed_other(name))
    try:
        prof_graph = M.GraphedPipeline(pipe, 1, FRAMES)
    finally:
        for name, fn in orig.items():
            setattr(ops, name, fn)
        ops.linear_pe = orig_linear_pe
    prof_graph(x_dev, idx)
    torch.cuda.synchronize()
    prof_steps = min(steps, 10)
    gemm_t = [0.0] * len(recs)
    other_t = {name: 0.0 for name in other}
    prof_ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(prof_steps):
        e0.record()
        prof_graph(None, None)
        e1.record()
        torch.cuda.synchronize()
        prof_ms += e0.elapsed_time(e1)
        for i, (_, _, _, a, b) in enumerate(recs):
            gemm_t[i] += a.elapsed_time(b)
        for name, evs in other.items():
            other_t[name] += sum(a.elapsed_time(b) for (a, b) in evs)
    breakdown = {"gemm": sum(gemm_t) / prof_steps}
    for name in other:
        breakdown[name] = other_t[name] / prof_steps
    breakdown = {k: {"ms_per_step": v, "share": v * prof_steps / prof_ms} for k, v in breakdown.items()}
    by_shape = {}
    for i, (m, n, k, _, _) in enumerate(recs):
        t = by_shape.setdefault((m, n, k), [0, 0.0])
        t[0] += 1
        t[1] += gemm_t[i] / prof_steps
    gemm_shapes = {f"{m}x{n}x{k}": {"launches_per_step": c, "us": 1e3 * ms / c,
                                    "tflops": 2.0 * m * n * k / (ms / c * 1e-3) / 1e12}
                   for (m, n, k), (c, ms) in by_shape.items()}
    gemm_ms = sum(gemm_t)
    gemm_fl = sum(2.0 * m * n * k for (m, n, k, _, _) in recs) * prof_steps
    n_gemm = len(recs) * prof_steps
    peaks = measured_peaks()
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tr_path):
        try:
            with open(tr_path) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 bf16 GEMM, all epilogues)", "achieved": achieved,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic,
                "peak_source": peaks["source"], "launches_per_step": n_gemm / prof_steps,
                "gflop_per_launch_avg": gemm_fl / max(1, n_gemm) / 1e9, "us_per_launch_avg": 1e3 * gemm_ms / max(1, n_gemm),
                "gemm_share_of_step": gemm_ms / prof_ms if prof_ms > 0 else None}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: the oracle (port of the reference's CPU path) on this box's host cores ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import numpy as np
        use_all_host_cores()
        w32 = {k: v.astype(np.float32) for k, v in weights.items()}
        x32 = x_host[0].float().numpy()
        oracle_pass(2, CHUNK, w32, x32, w32["positional_encoding.frame_embed"])   # BLAS warm-up
        dt, _ = oracle_pass(FRAMES, CHUNK, w32, x32, w32["positional_encoding.frame_embed"])
        cores = blas_threads()
        cpu = {"value": FRAMES / dt, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"the full workload once (1 video x {FRAMES} frames, fp32 numpy/OpenBLAS, {cores} threads, "
                         f"{dt:.1f} s)"}

    fps = world * FRAMES * steps / (ms_total * 1e-3)
    fps_e2e = world * FRAMES * steps / (ms_e2e * 1e-3)
    gflop_step = algorithmic_gflop(FRAMES, CHUNK)
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": x_host.numel() * x_host.element_size(),
                    "d2h_bytes_per_step": out_host.numel() * out_host.element_size(), "ms_per_step": ms_e2e / steps},
            "host_numa_node_rank0": numa, "gpu_launches": int(launches), "launch_mode": "one CUDA graph replay per step (kernels counted from an "
            "eager step)", "roofline": roofline, "cpu_baseline": cpu,
            "kernel_breakdown": breakdown, "gemm_shapes": gemm_shapes,
            "algorithmic_gflop_per_step": gflop_step,
            "path_tflops": gflop_step * 1e9 / (ms_total / steps * 1e-3) / 1e12,
            "path_frac_of_peak": gflop_step * 1e9 / (ms_total / steps * 1e-3) / 1e12 / peaks["tflops"]}
    emit_line(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_legacy(args):
    """`--workload legacy`: the Flash-VStream-style memories / scene segmentation of SURVEY.md 8f-4 on one GPU
    (tools/legacy_bench.py), with the numpy oracle timed beside it.  Headline of this line: merge_feature, 512 frames of
    729 x 1152 bf16 tokens kept to 3; the other ops are under "ops".  One step = one whole video."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import legacy_bench
    from oracle import legacy_memory_oracle as O
    frames, keep = 512, 3
    if args.impl == "reference":
        import numpy as np
        use_all_host_cores()
        steps = args.steps if args.steps is not None else 3
        warmup = args.warmup if args.warmup is not None else 1
        sample_frames = 24
        x = np.random.default_rng(0).standard_normal((sample_frames, 729, 1152)).astype(np.float32)
        for _ in range(warmup):
            O.merge_feature(x, keep)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.merge_feature(x, keep)
        dt = time.perf_counter() - t0
        fps = sample_frames * steps / dt
        cb = {"value": fps, "unit": "frames/s", "cores": blas_threads(), "kind": "port",
              "sample": f"merge_feature on the first {sample_frames} frames per step (numpy oracle)"}
        emit_line(json.dumps({"impl": "reference", "metric": "legacy_merge_feature_frames_per_s", "value": fps, "unit": "frames/s",
                          "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"legacy merge_feature, {frames} frames x 729 x 1152, keep {keep}"},
                          "gpu_launches": 0, "cpu_baseline": cb,
                          "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    import torch
    sampler = ClockSampler(0)
    sampler.mark_start()
    res = legacy_bench.run(frames, keep, cpu_frames=24, oracle=None if args.no_cpu_baseline else O, quiet=True)
    sampler.mark_end()
    clocks = sampler.stop()
    op = res["ops"]["stream_merge"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "legacy_stream_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")
    line = {"metric": "legacy_merge_feature_frames_per_s", "value": op["frames_per_s"], "unit": "frames/s", "n_gpus": 1,
            "steps": 3, "warmup": 1, "ms_per_step": op["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"legacy merge_feature (SURVEY 8f-4), {frames} frames x 729 x 1152 bf16 kept to {keep}; "
                                   "inputs (860 MB) larger than L2"},
            "clocks": clocks, "gpu_launches": op["launches"],
            "roofline": {"bound": "hbm", "achieved": op["achieved_gbps"], "peak": res["hbm_peak_gbps"], "unit": "GB/s",
                         "frac": op["frac_of_hbm_peak"], "traffic": traffic,
                         "note": "one launch per streamed frame; 6 rows of 1.68 MB per launch (5 read, 1 written); latency-chain bound"},
            "ops": res["ops"]}
    if "cpu_oracle" in res:
        c = res["cpu_oracle"]
        line["cpu_baseline"] = {"value": c["ops"]["stream_merge"]["frames_per_s"], "unit": "frames/s", "cores": c["cores"],
                                "kind": "port", "sample": "merge_feature, " + c["sample"], "ops": c["ops"]}
    emit_line(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="path", choices=["path", "legacy"],
                    help="path: the headline visual-memory path (BASELINE.json); legacy: SURVEY 8f-4 memories")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timed-only", action="store_true",
                    help="only the device-timed graph replays (no e2e / breakdown / CPU passes): the short run that is "
                         "put under `ncu` for the per-launch list in profiles/")
    args = ap.parse_args()
    isolate_stdout()
    if args.workload == "legacy":
        run_legacy(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
