#!/usr/bin/env python
"""Time one GEMM shape under every tile the kernel offers (development): python tools/tile_probe.py 1568x3584x3584 ..."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mavlm_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(1568, 3584, 3584), (1568, 3584, 14336),
                                                                         (1568, 14336, 3584), (3136, 3584, 14336)]
for m, n, k in shapes:
    x = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.02).bfloat16()
    b = torch.randn(n, device="cuda").bfloat16()
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    line = f"{m}x{n}x{k}:"
    for tile in (0, 1256, 1192, 1128, 256, 192, 128):
        lib.mavlm_debug_force_gemm_bn(tile)
        f = lambda: ops.linear(x, w, b, out=out)  # noqa: E731
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                f()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        line += f"  {tile if tile else 'auto'}: {us:6.1f} us"
    lib.mavlm_debug_force_gemm_bn(0)
    print(line)
