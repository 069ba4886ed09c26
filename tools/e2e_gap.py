#!/usr/bin/env python
"""Where the e2e step's extra time over the device-resident step goes: (a) one-graph replay, (b) the streamer's graphs
(pieces + assembly + recurrence) with no copies, (c) copies only, (d) the full HostStreamEncoder step -- order rotated,
idle time before every measurement (the power cap bites ~0.1 s into a burst)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import mavlm_b200 as M  # noqa: E402
from mavlm_b200 import synthetic  # noqa: E402

dev = torch.device("cuda:0")
F = 64
pipe, _ = synthetic.build_pipeline(3584, 1152, dtype=torch.bfloat16, chunk_size=32, device=dev)
x_host = synthetic.synthetic_tower_tokens(1, F, pin=True)
idx = torch.arange(F)[None]
g = pipe.graphed(1, F)
g(x_host.to(dev), idx)
pieces = int(sys.argv[1]) if len(sys.argv) > 1 else 4
enc = M.HostStreamEncoder(pipe, 1, F, pieces=pieces)
sl = enc.slots[0]
out_host = torch.empty(sl.seq.shape, dtype=torch.bfloat16, pin_memory=True)


def dev_step():
    g(None, None)


def slot_step():
    for gp in sl.g_piece:
        gp.replay()
    sl.assemble()
    sl.g_recur.replay()


def copy_step():
    with torch.cuda.stream(enc.s_in):
        sl.x.copy_(x_host.reshape(sl.x.shape), non_blocking=True)
    with torch.cuda.stream(enc.s_out):
        out_host.copy_(sl.seq, non_blocking=True)


def e2e_step():
    enc.submit(x_host, None, out_host)


def drain():
    enc.synchronize()
    enc.s_in.synchronize()


def timed(fn, n=40):
    for _ in range(3):
        fn()
    drain()
    torch.cuda.synchronize()
    time.sleep(0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    drain()
    e1.record()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / n


def slot_h2d_step():                                                 # compute + a concurrent H2D into a SPARE buffer
    with torch.cuda.stream(enc.s_in):
        spare_x.copy_(x_host.reshape(spare_x.shape), non_blocking=True)
    slot_step()


def slot_d2h_step():                                                 # compute + a concurrent D2H from a SPARE buffer
    with torch.cuda.stream(enc.s_out):
        out_host.copy_(spare_seq, non_blocking=True)
    slot_step()


spare_x = torch.zeros_like(sl.x)
spare_seq = torch.zeros_like(sl.seq)
variants = [("one graph", dev_step), ("slot graphs, no copies", slot_step), ("copies only", copy_step), ("e2e", e2e_step),
            ("compute + H2D", slot_h2d_step), ("compute + D2H", slot_d2h_step)]
for rnd in range(6):
    order = variants[rnd % 6:] + variants[:rnd % 6]
    res = {name: timed(fn) for name, fn in order}
    print(f"round {rnd}: " + "   ".join(f"{name}: {res[name]:.3f} ms" for name, _ in variants), flush=True)
