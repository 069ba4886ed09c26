#!/usr/bin/env python
"""Event timeline of the first CTA pair of the pair attention kernel (development).  python tools/attn_pair_trace.py [B]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mavlm_b200 import _lib, ops  # noqa: E402
lib = _lib.load()
b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, DH, lq, lk = 8, 448, 1568, 6272
q = torch.randn(b, lq, H * DH, device="cuda").bfloat16()
k = torch.randn(b, lk, H * DH, device="cuda").bfloat16()
v = torch.randn(b, lk, H * DH, device="cuda").bfloat16()
lib.mavlm_debug_set_flags(64)
for _ in range(2):
    ops.xattn(q, k, v, H)
buf = torch.zeros(2 * 3 * 256, dtype=torch.int64, device="cuda")
lib.mavlm_debug_attn_trace(buf.data_ptr())
ops.xattn(q, k, v, H)
torch.cuda.synchronize()
lib.mavlm_debug_attn_trace(None)
lib.mavlm_debug_set_flags(0)
t = buf.cpu().view(2, 3, 256)
names = {0: {1: "tma K", 2: "tma V"}, 1: {1: "QK issue", 2: "QK issued", 3: "PVown wait", 4: "PVpeer wait", 5: "PV go", 6: "PV issued"},
         2: {1: "sm wait S", 2: "sm S ready", 3: "sm ref sent", 4: "sm P computed", 5: "sm psend ok", 6: "sm P published", 7: "sm wait ref", 8: "sm ref got"}}
ev = []
for c in range(2):
    for r in range(3):
        for i in range(256):
            x = int(t[c, r, i])
            if x == 0:
                continue
            ev.append((x >> 8, c, r, x & 255))
ev.sort()
t0 = ev[0][0]
for ts, c, r, code in ev[:400]:
    print(f"{ts - t0:9d}  cta{c}  {'   ' * (3 * c + r)}{names[r].get(code, code)}")
