#!/usr/bin/env python
"""ncu target: the OV-7B chunk attention alone (B x 8 heads, 1568 queries x 6272 keys, head_dim 448).
  ncu --set full --import-source on -k regex:attn -s 2 -c 1 -o gpurun_out/attn python tools/ncu_attn_target.py [B] [flags]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mavlm_b200 import _lib, ops  # noqa: E402

bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 1
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
_lib.load().mavlm_debug_set_flags(flags)
h, dh, lq, lk = 8, 448, 1568, 6272
q = torch.randn(bsz, lq, h * dh, device="cuda").bfloat16()
k = torch.randn(bsz, lk, h * dh, device="cuda").bfloat16()
v = torch.randn(bsz, lk, h * dh, device="cuda").bfloat16()
for _ in range(4):
    ops.xattn(q, k, v, h)
torch.cuda.synchronize()
print("ok")
