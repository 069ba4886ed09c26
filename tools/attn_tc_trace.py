#!/usr/bin/env python
"""Where the single-CTA attention kernel's time goes (development): per-CTA event stamps of the MMA issuer and of one
softmax warp, summarised over all CTAs.  python tools/attn_tc_trace.py [B] [Lk] [Lq]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mavlm_b200 import _lib, ops  # noqa: E402
lib = _lib.load()
b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lk = int(sys.argv[2]) if len(sys.argv) > 2 else 6272
lq = int(sys.argv[3]) if len(sys.argv) > 3 else 1568
H, DH = 8, 448
q = torch.randn(b, lq, H * DH, device="cuda").bfloat16()
k = torch.randn(b, lk, H * DH, device="cuda").bfloat16()
v = torch.randn(b, lk, H * DH, device="cuda").bfloat16()
for _ in range(3):
    ops.xattn(q, k, v, H)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.xattn(q, k, v, H)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
print(f"B{b} Lq{lq} Lk{lk}: {us:.1f} us per call, {4 * b * H * lq * lk * DH / us / 1e6:.0f} TFLOP/s, "
      f"SM clock now {torch.cuda.clock_rate()} MHz")
GRID = 148
buf = torch.zeros(GRID * 2 * 64, dtype=torch.int64, device="cuda")
lib.mavlm_debug_attn_tc_trace(buf.data_ptr())
ops.xattn(q, k, v, H)
torch.cuda.synchronize()
lib.mavlm_debug_attn_tc_trace(None)
t = buf.cpu().view(GRID, 2, 64).tolist()
rows = []
for c in range(GRID):
    sm = [(x >> 8, x & 255) for x in t[c][1] if x]
    mm = [(x >> 8, x & 255) for x in t[c][0] if x]
    if not sm:
        continue
    t0 = sm[0][0]
    segs, cur = [], None
    for ts, code in sm[1:]:
        if code == 10:
            cur = {"start": ts - t0}
            segs.append(cur)
        elif code == 2:
            end = ts - t0
        elif cur is not None:
            cur[code] = ts - t0
    rows.append((c, segs, end, [(ts - mm[0][0], code) for ts, code in mm]))
print("cycles since the CTA's own start (softmax warp 2 lane 0).  per segment: start | first S | keys done | PV done | "
      "partial written(14) / flags seen(15) | epilogue done")
for c, segs, end, mm in rows[:: max(1, len(rows) // 24)]:
    line = f"cta {c:3d} end {end:7d}: "
    for s in segs:
        line += (f"[{s['start']:6d} {s.get(11, 0):6d} {s.get(12, 0):6d} {s.get(13, 0):6d} "
                 f"{'w' + str(s[14]) if 14 in s else ''}{'f' + str(s[15]) if 15 in s else ''} {s.get(16, 0):6d}] ")
    print(line)
import statistics as st  # noqa: E402
ends = [r[2] for r in rows]
first = [r[1][0].get(11, 0) for r in rows]
print(f"CTAs {len(rows)}; end cycles median {st.median(ends):.0f} max {max(ends)} min {min(ends)}; first S after "
      f"{st.median(first):.0f} cycles (median)")
gaps, epi_w, epi_m, tails = [], [], [], []
for c, segs, end, mm in rows:
    for i, s in enumerate(segs):
        if 14 in s:
            epi_w.append(s[16] - s[13])
        else:
            epi_m.append(s[16] - s[13])
        tails.append(s[13] - s[12])
        if i + 1 < len(segs):
            gaps.append(segs[i + 1].get(11, 0) - s[12])
print(f"keys-done -> PV-done median {st.median(tails):.0f}; epilogue (partial write) median "
      f"{st.median(epi_w) if epi_w else 0:.0f}; epilogue (final, with merge) median {st.median(epi_m) if epi_m else 0:.0f}; "
      f"keys-done -> next segment's first S median {st.median(gaps) if gaps else 0:.0f}")
