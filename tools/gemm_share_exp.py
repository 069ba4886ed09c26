import sys, torch
sys.path.insert(0, "/root/repo")
from mavlm_b200 import ops, _lib
lib = _lib.load()
dev = "cuda"
def graph_time(fn, n=10, reps=3):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (n * reps) * 1e3
lib.mavlm_debug_force_gemm_bn(1256)
K = 3584
for (mt, nt) in ((1, 74), (2, 37), (7, 14), (8, 37), (37, 2), (74, 1), (37, 8), (14, 7), (49, 56)):
    m, n = 256 * mt, 256 * nt
    a = torch.randn(m, K, device=dev).bfloat16(); w = torch.randn(n, K, device=dev).bfloat16()
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    us = graph_time(lambda: ops.linear(a, w, None, out=out))
    waves = -(-(mt * nt) // 74)
    print(f"pair tiles {mt:3d} x {nt:3d} = {mt*nt:5d} ({waves} waves)  {us:8.1f} us  {us/waves:7.1f} us/wave  {2*m*n*K/us/1e6:6.0f} TF  per-wave MMA eff {waves*56*512/1.9e3/us*100:5.1f}% @1.9GHz")
