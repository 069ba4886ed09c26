#!/usr/bin/env python
"""Graph-timed head_dim-448 attention at the bench / config shapes (20 launches per graph, best of 5 replays)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mavlm_b200 import _lib, ops  # noqa: E402

lib = _lib.load()

dev = "cuda"
h, dh = 8, 448
for (bsz, lq, lk) in ((1, 1568, 6272), (8, 1568, 6272), (1, 1568, 15680), (8, 1568, 3136), (1, 6272, 6272), (1, 50176, 6272)):
    q = torch.randn(bsz, lq, h * dh, device=dev).bfloat16()
    k = torch.randn(bsz, lk, h * dh, device=dev).bfloat16()
    v = torch.randn(bsz, lk, h * dh, device=dev).bfloat16()
    reps = 20 if lq < 20000 else 3
    graphs = {}
    for name, flags in (("default", 0),):                           # add (name, mavlm_debug_set_flags bits) pairs for an A/B
        lib.mavlm_debug_set_flags(flags)
        ops.xattn(q, k, v, h)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                o = ops.xattn(q, k, v, h)
        graphs[name] = g
    lib.mavlm_debug_set_flags(0)
    best = {n: 1e9 for n in graphs}
    for rnd in range(6):
        for name in (list(graphs) if rnd % 2 == 0 else list(graphs)[::-1]):
            time.sleep(0.3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graphs[name].replay()
            e1.record()
            torch.cuda.synchronize()
            best[name] = min(best[name], e0.elapsed_time(e1) / reps)
    fl = 4.0 * bsz * h * lq * lk * dh
    print(f"xattn B{bsz} {lq}x{lk} dh{dh}: " + "   ".join(f"{n}: {b * 1e3:.1f} us {fl / b / 1e9:.0f} TFLOP/s" for n, b in best.items()), flush=True)
