#!/usr/bin/env python
"""Small-shape pass over the round-2 kernels for compute-sanitizer (memcheck):
   compute-sanitizer --tool memcheck python tools/sanitize_target.py
Covers: the GEMM with a second (tail-fill) problem and tile ranges, the column-sum epilogue (ragged M / N), the CTA-pair
attention kernel (ragged tails, split-KV merge), the single-CTA attention kernel, and a whole bf16 pipeline step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mavlm_b200 import _lib, ops, synthetic  # noqa: E402

dev = "cuda"
lib = _lib.load()
torch.manual_seed(0)
# tail fill + tile ranges (ragged M and N on both problems)
x = torch.randn(300, 136, device=dev).bfloat16()
w = torch.randn(520, 136, device=dev).bfloat16()
b = torch.randn(520, device=dev).bfloat16()
zf = torch.randn(700, 136, device=dev).bfloat16()
wk = torch.randn(264, 136, device=dev).bfloat16()
out = torch.zeros(700, 264, device=dev).bfloat16()
work = ops.GemmWork(zf, wk, None, out)
y = ops.linear_fill(x, w, b, act=1, fillers=[work])
work.run()
ref = ops.linear(zf, wk, None)
print("fill equal:", bool(torch.equal(out, ref)), "primary finite:", bool(torch.isfinite(y.float()).all()))
# column sums (ragged keys / queries)
h = 2
for dh in (128, 448):
    q = torch.randn(1, 300, h * dh, device=dev).bfloat16()
    k = torch.randn(1, 333 * 8 // 8, h * dh, device=dev).bfloat16()
    v = torch.randn_like(k)
    o, lse, _ = ops.xattn(q, k, v, h, want_lse=True)
    cs = ops.xattn_colsum(q, k, lse, h)
    print("colsum total", float(cs.sum()), "expected", h * 300)
# pair attention kernel, incl. a split-KV merge (force 3 groups) and batch 2
h = 8
q = torch.randn(2, 200, h * 448, device=dev).bfloat16()
k = torch.randn(2, 456, h * 448, device=dev).bfloat16()
v = torch.randn_like(k)
o_ref, _, _ = ops.xattn(q, k, v, h)
lib.mavlm_debug_set_flags(64)
o_pair, _, _ = ops.xattn(q, k, v, h)
lib.mavlm_debug_force_attn_groups(3)
o_pair3, _, _ = ops.xattn(q, k, v, h)
lib.mavlm_debug_force_attn_groups(0)
lib.mavlm_debug_set_flags(0)
print("pair vs single max diff", float((o_pair.float() - o_ref.float()).abs().max()), float((o_pair3.float() - o_ref.float()).abs().max()))
# a whole small bf16 step (0.5B dims, 2 chunks, frame scores on)
pipe, _ = synthetic.build_pipeline(896, 1152, dtype=torch.bfloat16, chunk_size=2, device=dev, vocab=50000)
pipe.tail_fill = True
xs = synthetic.synthetic_tower_tokens(1, 4, 1152).to(dev)
res = pipe(xs, torch.arange(4)[None])
torch.cuda.synchronize()
print("pipeline ok", tuple(res["sequence"].shape), bool(torch.isfinite(res["sequence"].float()).all()))
