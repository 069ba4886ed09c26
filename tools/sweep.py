#!/usr/bin/env python
"""Sweeps for BASELINE configs 3 and 5 (parity-test / reporting cases, not the headline bench line):

  config 3: OV-7B, 256 frames in 16-frame chunks, 8 videos per GPU batch (sharded over ranks by bench.py-style
            weak scaling when run under torchrun), pooled-token input.
  config 5: OV-7B, 1024 frames, memory slots M in {8, 32, 64} (Lq = 196*M; 128/256 with --big) x chunk in
            {8, 16, 32}, pooled-token input, max_frames 1024.

    python tools/sweep.py [config3] [config5] [--big]      -> gpurun_out/sweep.json (+ stdout table)
Per point: ms, frames/s, algorithmic TFLOP/s (SURVEY.md §8d formulas, memory path only) and its fraction of
the measured sustained bf16 peak (a point runs for 0.1-60 s: the sustained peak applies), and -- as config 5 asks --
the roofline fraction PER KERNEL (GEMM / attention against the tensor peak, LayerNorm / assembly against the HBM
copy bandwidth), from a metered CUDA-graph capture of the point (mavlm_b200.meter).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mavlm_b200 import synthetic  # noqa: E402


def algo_gflop(frames, chunk, lq, d=3584, p=196, depth=2, cap=10):
    n_chunks = -(-frames // chunk)
    fl = 0.0
    for t in range(n_chunks):
        c = min(chunk, frames - t * chunk)
        fl += depth * (4.0 * lq * d * d + 4.0 * c * p * d * d + 4.0 * lq * c * p * d + 16.0 * lq * d * d)
        if t > 0:
            fl += 8.0 * lq * d * d + 4.0 * lq * (min(t, cap) * lq) * d
    fl += 16.0 * min(n_chunks, cap) * lq * d * d
    return fl / 1e9


def run_point(batch, frames, chunk, slots, iters=3, peaks=None):
    from mavlm_b200.meter import KernelMeter
    pipe, _ = synthetic.build_pipeline(3584, 1152, dtype=torch.bfloat16, chunk_size=chunk, device="cuda:0",
                                       num_memory_tokens=slots, max_frames=max(600, frames))
    g = torch.Generator(device="cuda:0").manual_seed(1234)
    z = torch.randn(batch, frames, 196, 3584, device="cuda:0", generator=g, dtype=torch.float32).bfloat16()
    # warm-up on two chunks (weight packs, function attributes, constant caches), then ONE metered capture of the point
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        pipe.memory_forward(z[:, : 2 * chunk], return_states=False)
        pipe.memory_forward(z[:, : 2 * chunk], return_states=False)
    torch.cuda.current_stream().wait_stream(side)
    pipe.prepare_constants(frames, z.device)
    torch.cuda.synchronize()
    with KernelMeter() as km:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = pipe.memory_forward(z, return_states=False)
    graph.replay()                                                  # first replay: page-in / clocks
    torch.cuda.synchronize()
    km.collect(graph.replay, iters)
    ms = km.step_ms
    gf = batch * algo_gflop(frames, chunk, 196 * slots)
    kernels = km.roofline_kernels(peaks["tflops_sustained"], peaks) if peaks else []
    res = {"batch": batch, "frames": frames, "chunk": chunk, "slots": slots, "lq": 196 * slots, "ms": ms,
           "frames_per_s": batch * frames / ms * 1e3, "algorithmic_tflops": gf / ms, "gflop": gf,
           "launches": len(km.recs), "instrumentation_launch_floor_us": 1e3 * km.launch_floor_ms,
           "kernels": [{"kernel": k["kernel"].split(" ")[0], "bound": k["bound"], "share": k["share_of_step"],
                        "achieved": k["achieved"], "unit": k["unit"], "frac": k["frac"],
                        "frac_net_of_launch_floor": k.get("frac_net_of_launch_floor")} for k in kernels]}
    del pipe, z, graph, out, km
    torch.cuda.empty_cache()
    return res


def main():
    peaks = {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        mp = json.load(open(pk))
        peaks = {"tflops_burst": mp.get("bf16_tflops", 1590.0), "tflops_sustained": mp.get("bf16_tflops_sustained", 1400.0),
                 "hbm_gbs": mp.get("hbm_gbs", 6650.0)}
    peak = peaks["tflops_sustained"]
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["config3", "config5"]
    out = []
    if "config3" in which:
        for b in (1, 8):
            out.append(dict(run_point(b, 256, 16, 8, peaks=peaks), config=3))
            print(json.dumps(out[-1]), flush=True)
    if "config5" in which:
        slots = (8, 32, 64) + ((128, 256) if "--big" in sys.argv else ())
        for m in slots:
            for c in (8, 16, 32):
                out.append(dict(run_point(1, 1024, c, m, iters=1 if m >= 32 else 3, peaks=peaks), config=5))
                print(json.dumps(out[-1]), flush=True)
    for r in out:
        r["frac_of_peak"] = r["algorithmic_tflops"] / peak
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w") as fh:
        json.dump({"peaks": peaks, "points": out}, fh, indent=1)
    print(f"{'cfg':>3} {'B':>2} {'F':>5} {'C':>3} {'M':>4} {'ms':>9} {'frames/s':>10} {'TFLOP/s':>8} {'frac':>5}")
    for r in out:
        per = "  ".join(f"{k['kernel'].replace('_kernel', '')}:{k['frac']:.2f}/{(k['frac_net_of_launch_floor'] or 0):.2f}({k['share']:.0%})"
                        for k in r["kernels"])   # raw / net of the instrumentation's per-launch floor (share of the step)
        print(f"{r['config']:>3} {r['batch']:>2} {r['frames']:>5} {r['chunk']:>3} {r['slots']:>4} {r['ms']:>9.2f} "
              f"{r['frames_per_s']:>10.0f} {r['algorithmic_tflops']:>8.0f} {r['frac_of_peak']:>5.2f}   {per}")


if __name__ == "__main__":
    main()
