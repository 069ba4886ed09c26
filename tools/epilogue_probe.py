#!/usr/bin/env python
"""What the GEMM epilogue costs on the short-K projector shape (development): the same 46656 x 3584 x 1152 GEMM with
bias + GELU, bias only, and no epilogue work, timed back to back in one CUDA graph each."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mavlm_b200 import ops  # noqa: E402

shapes = [(46656, 3584, 1152), (12544, 3584, 3584), (1568, 14336, 3584)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for m, n, k in shapes:
    x = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.02).bfloat16()
    b = torch.randn(n, device="cuda").bfloat16()
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    for name, kw in (("bias+gelu", dict(bias=b, act=ops.ACT_GELU_ERF)), ("bias", dict(bias=b)), ("plain", dict()),
                     ("bias+relu", dict(bias=b, act=ops.ACT_RELU))):
        bias = kw.pop("bias", None)
        f = lambda: ops.linear(x, w, bias, out=out, **kw)  # noqa: E731
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
        print(f"{m}x{n}x{k} {name:10s}: {us:7.1f} us  {2 * m * n * k / us / 1e6:6.0f} TFLOP/s")
