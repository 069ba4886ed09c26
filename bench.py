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
sequence inside the timed region; `e2e.copy_only` times the same copies without the compute),
`roofline` is the dominant kernel (tcgen05 GEMM) and `roofline_kernels` every kernel family of the
step (medians over >= 100 instrumented graph replays, CUDA events between the kernels: RAW numbers,
with `instrumentation_launch_floor_us` -- what the event nodes add to every launch -- reported beside
them, `frac_net_of_launch_floor` per kernel, and the bandwidth-bound kernels also timed `isolated`),
`e2e.single_video_host_to_host_ms` one submit + synchronize on an idle GPU, `sustained`
is a >= 5 s replay loop with its own clock record, `config3` is BASELINE config[2] (8 videos x 256
frames, 16-frame chunks, sharded 8/N per rank, NCCL all-gather of the assembled sequences inside the
timed region), `frame_sharded` (N > 1) one 1024-frame video with the frame-sharded pre-pass and the
piece-wise all-gather overlapped with the recurrence, `cpu_baseline` the UNMODIFIED reference modules
(baseline/_ref, see baseline/ref_arm.py) on the host cores.  --impl reference times those modules as
the reference arm (the numpy oracle port only when the reference install is absent).
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
    """Roofline denominators: MEASURED_PEAKS.json (driver-written: copy bandwidth, cuBLAS bf16 burst and sustained), else
    the fallback B200_PROFILING.md states."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        burst = float(p.get("bf16_tflops", 1590.0))
        return {"tflops_burst": burst, "tflops_sustained": float(p.get("bf16_tflops_sustained", burst)),
                "hbm_gbs": float(p.get("hbm_gbs", 6650.0)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    Two sources feed the same sample list: an in-process NVML poll every 5 ms (the counters nvidia-smi itself prints;
    it cannot miss a timed region of a few tens of milliseconds) and `nvidia-smi -lms 50` (the recipe's clocks line),
    whose start-up can take longer than a short timed region on a fresh box."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, device_index: int):
        self.samples = []          # (time, sm_mhz, sm_max_mhz, [reasons], source)
        self.proc = None
        self.t0 = self.t1 = None
        self._run = True
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = device_index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if device_index < len(ids) and ids[device_index].isdigit():
                    phys = int(ids[device_index])
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = {"hw_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            self._nvml = (pynvml, h, mx, bits)
            threading.Thread(target=self._poll, daemon=True).start()
        except Exception:
            self._nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(device_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        pynvml, h, mx, bits = self._nvml
        while self._run:
            try:
                sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.samples.append((time.time(), sm, mx, [n for n, b in bits.items() if mask & b], "nvml"))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 8:
                continue
            try:
                sm, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            self.samples.append((time.time(), sm, mx,
                                 [n for n, v in zip(self.NAMES, f[4:8]) if v.lower().startswith("active")], "nvidia-smi"))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is not None:
            time.sleep(0.08)
            self.proc.terminate()
        self._run = False
        sm, mx, reasons, src = [], [], set(), set()
        for t, s, m, r, source in list(self.samples):
            in_region = self.t0 is not None and self.t0 <= t <= (self.t1 or t)
            # nvidia-smi reports the previous 50 ms window: accept its line up to one period after the region
            late_smi = source == "nvidia-smi" and self.t0 is not None and self.t0 <= t <= (self.t1 or t) + 0.08
            if not (in_region or late_smi):
                continue
            sm.append(s)
            mx.append(m)
            reasons.update(r)
            src.add(source)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "source": "+".join(sorted(src))}


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


def use_all_host_cores() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host core."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n, user_api="blas")
        threadpool_limits(limits=n, user_api="openmp")
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    return n


def blas_threads() -> int:
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 0) for i in threadpool_info() if i.get("user_api") == "blas"]
        if n:
            return max(n)
    except Exception:
        pass
    return os.cpu_count() or 1


class ReferenceCpu:
    """The reference's own CPU implementation of the path (baseline/ref_arm.py): the UNMODIFIED modules from
    baseline/_ref (or /root/reference in the build container) behind prepare_inputs_labels_for_multimodal, fp32, all
    host threads; falls back to the numpy oracle port (kind "port") only when no reference install is present."""

    def __init__(self, weights32, frames_host_f32):
        import torch
        self.cores = use_all_host_cores()
        self.kind = "port"
        self.w = weights32
        self.x = frames_host_f32                                   # torch fp32 [F, 729, Dv]
        self.where = None
        try:
            from baseline import ref_arm
            root = ref_arm.find_reference_root()
            if root is not None:
                arch, pb = ref_arm.load_reference(root)
                self.model = ref_arm.build_reference_model(arch, pb, weights32, HIDDEN, VISION)
                self.arch, self.ref_arm = arch, ref_arm
                self.kind = "reference"
                self.where = os.path.relpath(root, ROOT) if root.startswith(ROOT) else root
        except Exception as e:                                      # a broken install must not kill the GPU arm
            print(f"bench.py: reference modules unavailable ({type(e).__name__}: {e}); timing the numpy port",
                  file=sys.stderr)
            self.kind = "port"
        if self.kind == "reference":
            self.cores = torch.get_num_threads()
        else:
            self.cores = blas_threads()

    def run(self, frames: int):
        """One pass over the first `frames` frames; returns (seconds, visual token sequence as numpy [L, D])."""
        if self.kind == "reference":
            video = self.ref_arm.tokens_as_video(self.x[:frames])
            t0 = time.perf_counter()
            seq = self.ref_arm.reference_pass(self.model, self.arch, video)
            return time.perf_counter() - t0, seq.numpy()
        dt, res = oracle_pass(frames, CHUNK, self.w, self.x.numpy(), self.w["positional_encoding.frame_embed"])
        return dt, res["sequence"]

    def describe(self, frames: int) -> str:
        if self.kind == "reference":
            return (f"UNMODIFIED reference modules ({self.where}: llava_arch.prepare_inputs_labels_for_multimodal -> "
                    f"mm_projector, get_2dPool, TemporalPositionalEncoding, TransformerProjector, memory_fuser, splice), "
                    f"torch CPU fp32, {self.cores} threads, 1 video x {frames} frames per pass")
        return (f"numpy/OpenBLAS fp32 port of the reference path (no reference install found), {self.cores} threads, "
                f"1 video x {frames} frames per pass")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    steps = args.steps if args.steps is not None else 3
    warmup = args.warmup if args.warmup is not None else 1
    w32 = host_weights_fp32()
    from mavlm_b200 import synthetic
    x = synthetic.synthetic_tower_tokens(1, FRAMES, dtype=torch.float32)[0]
    ref = ReferenceCpu(w32, x)
    # the full 64-frame workload per step; a bounded sample (one 32-frame chunk) only if K + W full passes would
    # not end within a few minutes on this box -- and then `config` says so
    t_first, _ = ref.run(FRAMES)                                    # also warms the thread pools
    frames = FRAMES
    if t_first * (steps + max(0, warmup - 1)) > 270.0:
        frames = CHUNK
    for _ in range(max(0, warmup - 1)):
        ref.run(frames)
    t = 0.0
    for _ in range(steps):
        dt, _ = ref.run(frames)
        t += dt
    fps = frames * steps / t
    sample = ref.describe(frames)
    cfg = workload_config(int(os.environ.get("WORLD_SIZE", "1")))
    if frames != FRAMES:
        cfg["workload"] += f" [reference arm: bounded sample, the first {frames} frames (one chunk) per step]"
        cfg["frames_per_video"] = frames
    cfg["parallelism"] = "one CPU process (rank 0), all host threads"
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "gpu_launches": 0,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit_line(json.dumps(line), flush=True)


def workload_config(world):
    return {"workload": "OV-7B dims (D=3584, SigLIP 729x1152 tokens/frame -> 196 pooled), 1 video x 64 frames per GPU, "
                        "chunks of 32: mm_projector + bilinear pool + PE + recurrent memory (formation x2, evolution "
                        "x1) + fuser MLP + token assembly",
            "frames_per_video": FRAMES, "videos_per_gpu": 1, "chunk": CHUNK, "hidden": HIDDEN, "parallelism":
            f"videos sharded over {world} GPU(s), weights replicated, no data-path collective",
            "l2_policy": "working set per step (0.97 GB weights + 107 MB input) exceeds the 126 MB L2; no flush needed"}


def _median(xs):
    return statistics.median(xs) if xs else 0.0


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
    # ---- instrumented pass: per-launch CUDA-event timing of every op, INSIDE a CUDA graph of the step ----
    # (mavlm_b200.meter: the step is captured a second time with an external timing event around every library call;
    # medians over >= 100 replays)
    from mavlm_b200.meter import KERNEL_NAMES, KernelMeter
    with KernelMeter() as km:
        prof_graph = M.GraphedPipeline(pipe, 1, FRAMES)
    prof_graph(x_dev, idx)
    torch.cuda.synchronize()
    prof_steps = max(100, min(steps, 200))
    # 100 replays back to back (0.5 s) run into the power cap half-way, i.e. into another clock regime than the 0.08 s
    # timed region: the pass runs in bursts of 10 with idle time in between, and reads the SM clock at the end of each
    burst_clocks = []

    def read_clock():
        try:
            burst_clocks.append(float(torch.cuda.clock_rate(local)))
        except Exception:
            pass

    km.collect(lambda: prof_graph(None, None), prof_steps, burst=10, pause_s=0.3, on_burst=read_clock if rank == 0 else None)
    clocks_prof = {"sm_mhz_at_burst_ends": burst_clocks, "sm_mhz": (statistics.median(burst_clocks) if burst_clocks else None),
                   "bursts": "10 replays per burst, 0.3 s idle between bursts"}
    step_ms_prof = km.step_ms
    peaks = measured_peaks()
    timed_region_s = ms_total * 1e-3
    burst_applies = timed_region_s < 1.0                    # a sub-second timed region runs at burst clocks
    tpeak = peaks["tflops_burst"] if burst_applies else peaks["tflops_sustained"]
    tpeak_name = "burst" if burst_applies else "sustained"
    fams = km.families()
    kernel_names = KERNEL_NAMES
    roofline_kernels = km.roofline_kernels(tpeak, peaks)
    breakdown = {fam: {"ms_per_step": f["ms"], "share": f["ms"] / step_ms_prof} for fam, f in fams.items()}
    gemm_shapes = km.shapes("gemm")
    attn_shapes = km.shapes("xattn")
    # ---- the bandwidth-bound kernels once more, ISOLATED: 20 back-to-back launches of the kernel alone in a plain CUDA
    # graph (no event nodes, launches overlap as in the un-instrumented step), at the step's shapes.  A 6 us kernel cannot
    # be timed between two event nodes (the launch floor above is as long as the kernel).
    if not args.no_extras:
        iso = isolated_hbm_kernels(torch, ops, dev, peaks["hbm_gbs"])
        for ent in roofline_kernels:
            for fam, rec in iso.items():
                if ent["kernel"] == kernel_names.get(fam):
                    ent["isolated"] = rec
    for _ in range(3):
        step_e2e()
    streamer.synchronize()
    ms_e2e = timed(lambda: step_e2e(), steps, after=streamer.synchronize)

    # ---- copy-only pass: the SAME per-step H2D + D2H copies on the streamer's copy streams, no compute.  When this
    # alone takes about as long as the e2e step, e2e is bound by the host side (pinned-memory / PCIe), not the GPU.
    g0 = streamer.slots[0]

    def step_copy_only():
        with torch.cuda.stream(streamer.s_in):
            g0.x.copy_(x_host.reshape(g0.x.shape), non_blocking=True)
        with torch.cuda.stream(streamer.s_out):
            out_host.copy_(g0.seq, non_blocking=True)

    def drain_copies():
        cur = torch.cuda.current_stream(dev)
        cur.wait_stream(streamer.s_in)
        cur.wait_stream(streamer.s_out)

    for _ in range(3):
        step_copy_only()
    drain_copies()
    ms_copy = timed(step_copy_only, steps, after=drain_copies)

    # ---- one video alone, host to host: submit -> synchronize on an idle GPU (what a serving request waits for)
    lat = []
    for _ in range(7):
        torch.cuda.synchronize()
        time.sleep(0.02)
        t0 = time.perf_counter()
        step_e2e()
        streamer.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    single_ms = _median(lat)

    g = fams.get("gemm", {"launches": 0, "ms": 0.0, "flops": 0.0})
    achieved = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tr_path):
        try:
            with open(tr_path) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": kernel_names["gemm"], "achieved": achieved,
                "peak": tpeak, "unit": "TFLOP/s", "frac": achieved / tpeak, "traffic": traffic,
                "peak_source": f"{peaks['source']}: bf16 {tpeak_name} peak (the timed region is {timed_region_s:.2f} s"
                               f"{' < 1 s, i.e. burst clocks' if burst_applies else ''})",
                "frac_of_burst": achieved / peaks["tflops_burst"], "frac_of_sustained": achieved / peaks["tflops_sustained"],
                "launches_per_step": g["launches"], "gflop_per_launch_avg": g["flops"] / max(1, g["launches"]) / 1e9,
                "us_per_launch_avg": 1e3 * g["ms"] / max(1, g["launches"]),
                "share_of_step": g["ms"] / step_ms_prof if step_ms_prof > 0 else None,
                "method": f"median over {prof_steps} instrumented graph replays (CUDA events between the kernels) in bursts of "
                          "10, right after the timed region; the event nodes cost the kernels their programmatic-dependent-"
                          "launch overlap, so the instrumented step is slower than the timed one",
                "achieved_scaled_to_timed_step": achieved * step_ms_prof / (ms_total / steps) if ms_total > 0 else None,
                "achieved_scaled_note": "achieved x instrumented_step_ms / timed_step_ms: the GEMM rate inside the timed "
                                        "region if every kernel family sped up alike (an estimate, not a measurement)",
                "instrumented_step_ms": step_ms_prof, "timed_step_ms": ms_total / steps,
                "instrumentation_launch_floor_us": 1e3 * km.launch_floor_ms,
                "instrumentation_launch_floor_note": "event -> 8-element cast kernel -> event inside the same instrumented "
                                                     "graph: what the event nodes and the un-overlapped launch add to every "
                                                     "metered launch; `achieved` / `frac` are RAW (floor included), "
                                                     "roofline_kernels[*].frac_net_of_launch_floor takes it out",
                "achieved_net_of_launch_floor": (g["flops"] / ((g["ms"] - g["launches"] * km.launch_floor_ms) * 1e-3) / 1e12
                                                 if g["ms"] > g["launches"] * km.launch_floor_ms else None),
                "clocks_during_instrumented_pass": clocks_prof}

    # ---- sustained: >= 5 s of back-to-back replays with their own clock record (what the path holds under the power cap)
    sustained = None
    if not args.no_extras:
        n_sus = max(steps, int(5200.0 / max(ms_total / steps, 1e-3)))
        sampler2 = ClockSampler(local) if rank == 0 else None
        time.sleep(0.1)
        if sampler2:
            sampler2.mark_start()
        ms_sus = timed(step_device, n_sus)
        if sampler2:
            sampler2.mark_end()
        clocks2 = sampler2.stop() if sampler2 else None
        gf_exec = sum(f["flops"] for f in fams.values()) / 1e9
        sustained = {"seconds": ms_sus * 1e-3, "steps": n_sus, "ms_per_step": ms_sus / n_sus,
                     "value": world * FRAMES * n_sus / (ms_sus * 1e-3), "unit": "frames/s", "clocks": clocks2,
                     "executed_tflops": gf_exec * 1e9 / (ms_sus / n_sus * 1e-3) / 1e12,
                     "frac_of_sustained_peak": gf_exec * 1e9 / (ms_sus / n_sus * 1e-3) / 1e12 / peaks["tflops_sustained"]}

    # ---- BASELINE config[2]: 8 videos x 256 frames, 16-frame chunks, videos sharded 8/N per rank, NCCL all-gather of
    # the assembled sequences inside the timed region (strong scaling: the total work is fixed)
    config3 = frame_sharded = None
    if not args.no_extras:
        config3 = bench_config3(torch, dist, M, synthetic, dev, rank, world, barrier)
        if world > 1:
            frame_sharded = bench_frame_sharded(torch, dist, M, synthetic, dev, rank, world, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: the UNMODIFIED reference modules on this box's host cores, the full workload once; its output
    # is also compared with the GPU's (the same weights are loaded into the reference modules)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import numpy as np
        w32 = {k: v.astype(np.float32) for k, v in weights.items()}
        ref = ReferenceCpu(w32, x_host[0].float())
        ref.run(8)                                           # thread-pool warm-up (8 frames)
        dt, seq_ref = ref.run(FRAMES)
        got = graphed(x_dev, idx)["sequence"][0].float().cpu().numpy()
        par = float(np.abs(got - seq_ref).max() / np.abs(seq_ref).max())
        cpu = {"value": FRAMES / dt, "unit": "frames/s", "cores": ref.cores, "kind": ref.kind,
               "sample": f"the full workload once ({dt:.1f} s): " + ref.describe(FRAMES),
               "gpu_vs_cpu_sequence_err": par,
               "gpu_vs_cpu_note": "max|gpu - cpu| / max|cpu| over the assembled sequence; bf16 GPU path vs the fp32 CPU run "
                                  "on UNROUNDED weights, so it includes the bf16 weight rounding (tolerance tests round first)"}

    fps = world * FRAMES * steps / (ms_total * 1e-3)
    fps_e2e = world * FRAMES * steps / (ms_e2e * 1e-3)
    gflop_step = algorithmic_gflop(FRAMES, CHUNK)
    h2d = x_host.numel() * x_host.element_size()
    d2h = out_host.numel() * out_host.element_size()
    copy_gbs = world * (h2d + d2h) * steps / (ms_copy * 1e-3) / 1e9
    e2e_bound = ("host copies (pinned memory / PCIe): the copies alone take "
                 f"{ms_copy / ms_e2e:.0%} of the e2e step" if ms_copy >= 0.85 * ms_e2e else
                 f"GPU compute: the copies alone take {ms_copy / ms_e2e:.0%} of the e2e step and overlap it")
    line = {"metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(world), "clocks": clocks,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / steps, "bound": e2e_bound,
                    "single_video_host_to_host_ms": single_ms,
                    "single_video_note": "median wall time of one submit() + synchronize() on an idle GPU; the input is "
                                         f"copied in {len(g0.ranges)} pieces that the projector consumes as they land and "
                                         "the frame rows of the sequence travel back while the recurrence runs (the serial "
                                         "sum would be H2D + compute + D2H)",
                    "copy_only": {"ms_per_step": ms_copy / steps, "host_gbs_all_ranks": copy_gbs,
                                  "note": "the same H2D + D2H copies per step on the copy streams with no compute, all "
                                          "ranks at once (max over ranks)"}},
            "host_numa_node_rank0": numa, "gpu_launches": int(launches), "launch_mode": "one CUDA graph replay per step (kernels counted from an "
            "eager step)", "roofline": roofline, "roofline_kernels": roofline_kernels, "cpu_baseline": cpu,
            "sustained": sustained, "config3": config3, "frame_sharded": frame_sharded,
            "kernel_breakdown": breakdown, "gemm_shapes": gemm_shapes, "attn_shapes": attn_shapes,
            "algorithmic_gflop_per_step": gflop_step,
            "executed_gflop_per_step": sum(f["flops"] for f in fams.values()) / 1e9,
            "path_tflops": gflop_step * 1e9 / (ms_total / steps * 1e-3) / 1e12,
            "path_frac_of_peak": gflop_step * 1e9 / (ms_total / steps * 1e-3) / 1e12 / tpeak,
            "path_frac_note": f"SURVEY 8d algorithmic count (projector at 729 rows, reference order) over the bf16 {tpeak_name} "
                              "peak; executed_gflop_per_step is what the kernels actually run (W2 after the pool)"}
    emit_line(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def isolated_hbm_kernels(torch, ops, dev, hbm_gbs, reps=20, replays=9):
    """LayerNorm / pool + PE / assembly at the shapes of one bench step, each as `reps` back-to-back launches in its own
    CUDA graph; median over `replays` of (graph time / reps).  LayerNorm reads what the GEMM ahead of it has just written
    (L2-resident in the step as in this loop); pool and assembly stream tensors larger than half the L2."""
    lq, d, p = 1568, HIDDEN, 196
    out = {}

    def graph_us(fn):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(replays):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
            time.sleep(0.05)
        return _median(ts)

    def rec(us, nbytes, what):
        gbs = nbytes / us / 1e3
        return {"us_per_launch": us, "achieved": gbs, "unit": "GB/s", "peak": hbm_gbs, "frac": gbs / hbm_gbs,
                "bytes_per_launch": nbytes, "what": what,
                "method": f"{reps} back-to-back launches in a plain CUDA graph, median of {replays} replays (CUDA events)"}

    bf = torch.bfloat16
    x = torch.randn(lq, d, device=dev)
    g_ = torch.randn(d, device=dev).to(bf)
    b_ = torch.randn(d, device=dev).to(bf)
    y = torch.empty(lq, d, device=dev, dtype=bf)
    out["layernorm"] = rec(graph_us(lambda: ops.layernorm(x, g_, b_, 1e-12, out_dtype=bf, out=y)), lq * d * 6,
                           f"{lq} x {d} fp32 pre-LN sum -> bf16")
    xt = torch.randn(FRAMES, 729, d, device=dev).to(bf)
    table = torch.randn(600, d, device=dev)
    fidx = torch.arange(FRAMES, device=dev)
    yp = ops.pool_pe(xt, side=27, pe_table=table, frame_idx=fidx)
    out["pool_pe"] = rec(graph_us(lambda: ops.pool_pe(xt, side=27, pe_table=table, frame_idx=fidx)),
                         xt.numel() * 2 + yp.numel() * 2, f"{FRAMES} frames 27x27 -> 14x14 + PE")
    fine = torch.arange(0, FRAMES, max(1, FRAMES // 32), device=dev)[:32]
    emb = torch.randn(2, d, device=dev).to(bf)
    nl = torch.randn(d, device=dev).to(bf)
    tab = torch.randn(1000, d, device=dev).to(bf)
    pm = torch.arange(10, device=dev)
    pf = torch.arange(9, device=dev)
    n_mem = 2 * lq
    rows = 10 + 1 + 9 + len(fine) * p + 1
    seq = torch.zeros(rows + n_mem, d, device=dev, dtype=bf)
    out["assemble"] = rec(graph_us(lambda: ops.assemble(seq, None, n_mem, yp, fine, p, emb, nl, tab, pm, pf)),
                          2 * rows * d * 2, f"{rows} rows (memory tokens already in place)")
    return out


def bench_config3(torch, dist, M, synthetic, dev, rank, world, barrier, videos=8, frames=256, chunk=16, n_timed=5):
    """BASELINE config[2] measured in this process: `videos` x `frames` frames in `chunk`-frame chunks (16 chunks > the
    10-deep cache), sharded videos/world per rank, then ONE ncclAllGather of the assembled sequences [V, L, D] so every rank
    holds all results -- inside the timed region.  Device-timed, max over ranks.  At world > 1 rank 0 also times the whole
    batch alone, so the line carries its own strong-scaling efficiency."""
    if videos % world != 0:
        return {"skipped": f"{videos} videos do not split evenly over {world} ranks"}
    pipe, _ = synthetic.build_pipeline(HIDDEN, VISION, dtype=torch.bfloat16, chunk_size=chunk, device=dev)
    per = videos // world
    mine = M.dist.shard_range(videos, rank, world)

    def video_tokens(v):                                     # synthetic tower tokens generated ON the device (seeded per video)
        gen = torch.Generator(device=dev).manual_seed(1234 + v)
        return torch.randn((frames, 729, VISION), generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)

    idx = torch.arange(frames)[None].expand(per, frames)
    g = pipe.graphed(per, frames)
    g(torch.stack([video_tokens(v) for v in mine]), idx)
    seq = g.out["sequence"]
    full = torch.empty((videos,) + tuple(seq.shape[1:]), dtype=seq.dtype, device=dev) if world > 1 else None

    def step(gather=True):
        g(None, None)
        if world > 1 and gather:
            dist.all_gather_into_tensor(full, g.out["sequence"])

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / n

    for _ in range(3):
        step()
    ms = timed(step, n_timed)
    out = {"workload": f"OV-7B dims, bf16, {videos} videos x {frames} frames, chunks of {chunk} ({frames // chunk} chunks: the "
                       f"10-deep state ring wraps), {per} video(s) per rank batched, ncclAllGather of the assembled sequences "
                       f"[{videos}, {seq.shape[1]}, {seq.shape[2]}] inside the timed region",
           "scaling": "strong", "n_gpus": world, "videos_per_rank": per, "ms_per_step": ms,
           "value": videos * frames / (ms * 1e-3), "unit": "frames/s", "steps": n_timed, "warmup": 3}
    if world > 1:
        ms_nogather = timed(lambda: step(False), n_timed)

        def gather_only():
            dist.all_gather_into_tensor(full, g.out["sequence"])

        for _ in range(2):
            gather_only()
        ms_gather = timed(gather_only, n_timed)
        recv = (world - 1) * seq.numel() * seq.element_size()
        out.update({"compute_ms": ms_nogather, "gather_ms": ms_gather, "gather_exposed_ms": ms - ms_nogather,
                    "gather_bytes_received_per_rank": recv, "gather_gbs_per_rank": recv / (ms_gather * 1e-3) / 1e9,
                    "collective": "ncclAllGather (torch.distributed all_gather_into_tensor) on the compute stream after "
                                  "the graph replay; 770 GB/s per direction is the measured NVLink peer-copy reference"})
        # strong-scaling base: the whole batch on rank 0 alone
        barrier()
        ms1 = None
        if rank == 0:
            idx1 = torch.arange(frames)[None].expand(videos, frames)
            g1 = pipe.graphed(videos, frames)
            g1(torch.stack([video_tokens(v) for v in range(videos)]), idx1)
            for _ in range(2):
                g1(None, None)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                g1(None, None)
            e1.record()
            torch.cuda.synchronize()
            ms1 = e0.elapsed_time(e1) / 3
            del g1
        barrier()
        if ms1 is not None:
            out.update({"one_gpu_ms_per_step": ms1, "strong_scaling_efficiency": ms1 / (world * ms)})
    del g, pipe
    torch.cuda.empty_cache()
    return out


def bench_frame_sharded(torch, dist, M, synthetic, dev, rank, world, barrier, frames=1024, chunk=32, n_timed=3):
    """ONE 1024-frame OV-7B video on `world` GPUs: frame-sharded projector + pool + PE, piece-wise
    ncclAllGather of the pooled tokens on a side stream overlapped with the replicated recurrence
    (dist.FrameShardedEncoder).  Reports the step with the gathers overlapped, with every gather waited for before the
    recurrence (blocking), with the gathers skipped (compute alone) and the replicated single-GPU step."""
    pipe, _ = synthetic.build_pipeline(HIDDEN, VISION, dtype=torch.bfloat16, chunk_size=chunk, device=dev, max_frames=frames)
    gen = torch.Generator(device=dev).manual_seed(777)
    x = torch.randn((frames, 729, VISION), generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    idx = torch.arange(frames)
    enc = M.dist.FrameShardedEncoder(pipe, frames)

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / n

    res = {}
    for name, kw in (("overlapped", dict(overlap=True)), ("blocking", dict(overlap=False)), ("no_comm", dict(comm=False))):
        for _ in range(2):
            enc(x, idx, **kw)
        res[name] = timed(lambda: enc(x, idx, **kw), n_timed)
    enc(x, idx, overlap=True, time_gathers=True)
    torch.cuda.synchronize()
    gather_ms = enc.gather_ms()
    gbytes = enc.gathered_bytes
    t = torch.tensor([gather_ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gather_ms = float(t.item())
    # parity of the sharded path with the plain single-GPU path on the same video (bitwise: same kernels, same rows)
    seq_sharded = enc(x, idx)["sequence"].clone()
    seq_plain = pipe(x[None], idx[None], return_states=False)["sequence"]
    same = bool(torch.equal(seq_sharded, seq_plain))
    for _ in range(1):
        pipe(x[None], idx[None], return_states=False)
    ms_plain = timed(lambda: pipe(x[None], idx[None], return_states=False), n_timed)
    exposed = max(0.0, res["overlapped"] - res["no_comm"])
    out = {"workload": f"OV-7B dims, bf16, ONE video x {frames} frames, chunks of {chunk}: each rank projects / pools / PEs the "
                       f"chunks it owns (one chunk per rank per piece of {world} chunks), pooled tokens all-gathered piece by piece "
                       "on a side stream, recurrence + fuser + assembly replicated on every rank (eager launches)",
           "n_gpus": world, "ms_per_step_overlapped": res["overlapped"], "ms_per_step_blocking_gathers": res["blocking"],
           "ms_per_step_no_comm": res["no_comm"], "ms_per_step_single_gpu_unsharded": ms_plain,
           "value": frames / (res["overlapped"] * 1e-3), "unit": "frames/s",
           "speedup_vs_unsharded": ms_plain / res["overlapped"],
           "gather_ms_total": gather_ms, "gather_ms_exposed": exposed, "gather_ms_hidden": max(0.0, gather_ms - exposed),
           "gather_bytes_received_per_rank": gbytes, "gather_gbs_per_rank": gbytes / (gather_ms * 1e-3) / 1e9 if gather_ms > 0 else None,
           "collective": f"{len(enc.sched)} x ncclAllGather of [{world} x {chunk} frames x 196 x {HIDDEN}] bf16 on a side stream",
           "sharded_equals_unsharded_bitwise": same,
           "limit": "only the projector + pool + PE (11 of ~75 executed GFLOP per frame) shard; the recurrence is sequential in time"}
    del enc, pipe
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sustained loop, config3 and frame_sharded objects (quick developer runs)")
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
