#!/usr/bin/env python
"""Kernel time table of one OV-7B training step (forward + backward, batch 8 x 32 frames) via torch.profiler."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from mavlm_b200 import synthetic  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32
pipe, _ = synthetic.build_pipeline(3584, 1152, dtype=torch.bfloat16, chunk_size=32, device="cuda:0")
z = torch.randn(8, frames, 196, 3584, device="cuda:0").bfloat16()


def step():
    for p_ in pipe.parameters():
        p_.grad = None
    (pipe.memory_forward_train(z)["sequence"].float() ** 2).mean().backward()


step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total)
        for e in prof.key_averages()]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"total device time {tot / 1e3:.2f} ms")
for k, c, t in rows[:28]:
    print(f"{t / 1e3:8.2f} ms {100 * t / tot:5.1f}%  n={c:4d}  {k[:110]}")
