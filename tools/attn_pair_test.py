#!/usr/bin/env python
"""Developer harness for the CTA-pair attention kernel (attn_pair.cu): pair vs single-CTA kernel vs an fp32 torch
reference on the GPU, growing shapes; then timing.   python tools/attn_pair_test.py [stage]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mavlm_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
dev = "cuda"
H, DH = 8, 448


def ref_attn(q, k, v, h):
    b, lq, hd = q.shape
    dh = hd // h
    qh = q.float().view(b, lq, h, dh).transpose(1, 2)
    kh = k.float().view(b, -1, h, dh).transpose(1, 2)
    vh = v.float().view(b, -1, h, dh).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / dh ** 0.5
    o = (s.softmax(-1) @ vh).transpose(1, 2).reshape(b, lq, hd)
    return o, torch.logsumexp(s, -1)


def run(b, lq, lk, h=H, qscale=1.0, seed=0, groups=0):
    torch.manual_seed(seed)
    q = (torch.randn(b, lq, h * DH, device=dev) * qscale).bfloat16()
    k = torch.randn(b, lk, h * DH, device=dev).bfloat16()
    v = torch.randn(b, lk, h * DH, device=dev).bfloat16()
    if lk > 3:
        k[0, -3] = q[0, min(7, lq - 1)] * 0.5
    lib.mavlm_debug_force_attn_groups(groups)
    lib.mavlm_debug_set_flags(0)
    o_old, lse_old, _ = ops.xattn(q, k, v, h, want_lse=True)
    lib.mavlm_debug_set_flags(64)
    o_new, lse_new, _ = ops.xattn(q, k, v, h, want_lse=True)
    lib.mavlm_debug_set_flags(0)
    torch.cuda.synchronize()
    lib.mavlm_debug_force_attn_groups(0)
    o_ref, lse_ref = ref_attn(q, k, v, h)
    sc = o_ref.abs().max()
    e_new = float((o_new.float() - o_ref).abs().max() / sc)
    e_old = float((o_old.float() - o_ref).abs().max() / sc)
    l_new = float((lse_new - lse_ref).abs().max())
    print(f"B{b} H{h} Lq{lq} Lk{lk} qs{qscale} g{groups}: pair err {e_new:.3e} (single {e_old:.3e}) lse err {l_new:.2e}", flush=True)
    if not (e_new < 2e-2 and l_new < 1e-3):
        bad = (o_new.float() - o_ref).abs().amax(dim=-1)[0]
        rows = torch.nonzero(bad > 2e-2 * sc).flatten()[:10].tolist()
        cols = (o_new.float() - o_ref).abs().amax(dim=(0, 1))
        print("   bad rows:", rows, " bad col blocks:", [int(c) for c in torch.nonzero(cols.view(-1, 64).amax(-1) > 2e-2 * sc).flatten()[:16]])
        return False
    return True


def timeit(b, lq, lk, flags, n=20):
    q = torch.randn(b, lq, H * DH, device=dev).bfloat16()
    k = torch.randn(b, lk, H * DH, device=dev).bfloat16()
    v = torch.randn(b, lk, H * DH, device=dev).bfloat16()
    lib.mavlm_debug_set_flags(flags)
    for _ in range(3):
        ops.xattn(q, k, v, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ops.xattn(q, k, v, H)
    e1.record()
    torch.cuda.synchronize()
    lib.mavlm_debug_set_flags(0)
    us = e0.elapsed_time(e1) * 1e3 / n
    return us, 4.0 * b * lq * lk * H * DH / us / 1e6


stage = int(sys.argv[1]) if len(sys.argv) > 1 else 99
ok = True
cases = [(1, 128, 128, 1), (1, 128, 256, 1), (1, 128, 384, 1), (1, 128, 1024, 1), (1, 100, 200, 1), (1, 300, 1000, 2),
         (1, 1568, 1568, 8), (1, 1568, 6272, 8), (2, 1568, 6272, 8), (1, 1568, 15680, 8), (1, 300, 4804, 8)]
for i, (b, lq, lk, h) in enumerate(cases):
    if i >= stage:
        break
    ok &= run(b, lq, lk, h=h)
    if not ok:
        break
if ok and stage > len(cases):
    ok &= run(1, 1568, 6272, qscale=8.0)
    ok &= run(1, 300, 4804, qscale=8.0)
    ok &= run(1, 1568, 6272, groups=3)
    ok &= run(8, 1568, 3136)
print("ALL OK" if ok else "FAILED", flush=True)
if ok and stage > len(cases):
    for (b, lq, lk) in ((1, 1568, 6272), (1, 1568, 1568), (1, 1568, 15680), (8, 1568, 6272), (8, 1568, 3136)):
        t_old = timeit(b, lq, lk, 0)
        t_new = timeit(b, lq, lk, 64)
        print(f"B{b} Lq{lq} Lk{lk}: single {t_old[0]:.1f} us {t_old[1]:.0f} TF | pair {t_new[0]:.1f} us {t_new[1]:.0f} TF", flush=True)
sys.exit(0 if ok else 1)
