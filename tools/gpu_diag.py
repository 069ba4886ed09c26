#!/usr/bin/env python
"""GPU bring-up diagnostics: every kernel family against torch references, each family in its own
subprocess with a timeout so that one faulting kernel does not hide the others.

    python tools/gpu_diag.py [family ...]        # families: elem simt gemm attn pipe perf
Writes gpurun_out/diag_<family>.log.  Not a test (tests/ holds the parity tests) and not a benchmark.
"""
from __future__ import annotations

import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def nerr(a, b):
    import torch
    a = a.double()
    b = b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def fam_elem():
    import torch
    import torch.nn.functional as F
    from mavlm_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(5, 729, 64, device=dev).to(dt)
        ref = F.interpolate(x.float().view(5, 27, 27, 64).permute(0, 3, 1, 2), size=(14, 14), mode="bilinear")
        ref = ref.permute(0, 2, 3, 1).reshape(5, 196, 64)
        y = ops.pool_pe(x, side=27)
        print(f"pool bilinear {dt}: err {nerr(y.float(), ref):.3e}")
        table = torch.randn(600, 64, device=dev)
        idx = torch.tensor([0, 5, 17, 300, 599], device=dev)
        y2 = ops.pool_pe(x, side=27, pe_table=table, frame_idx=idx)
        print(f"pool+pe {dt}: err {nerr(y2.float(), ref + table[idx][:, None, :]):.3e}")
        for mode, fn in (("average", F.avg_pool2d), ("max", F.max_pool2d)):
            r = fn(x.float().view(5, 27, 27, 64).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).reshape(5, 169, 64)
            print(f"pool {mode} {dt}: err {nerr(ops.pool_pe(x, side=27, mode=mode).float(), r):.3e}")
        xa = torch.randn(5, 7, 64, device=dev).to(dt)
        print(f"add_pe {dt}: err {nerr(ops.add_pe(xa, table, idx).float(), xa.float() + table[idx][:, None, :]):.3e}")
        for d in (32, 896, 3584):
            xl = (torch.randn(77, d, device=dev) * 3 + 1)
            g = torch.randn(d, device=dev).to(dt)
            b = torch.randn(d, device=dev).to(dt)
            ref = F.layer_norm(xl.double(), (d,), g.double(), b.double(), 1e-12)
            print(f"layernorm f32->{dt} d={d}: err {nerr(ops.layernorm(xl, g, b, 1e-12, out_dtype=dt).float(), ref):.3e}")
            xin = xl.to(dt)
            ref = F.layer_norm(xin.double(), (d,), g.double(), b.double(), 1e-12)
            print(f"layernorm {dt}->{dt} d={d}: err {nerr(ops.layernorm(xin, g, b, 1e-12, out_dtype=dt).float(), ref):.3e}")
        # assembly
        d, p = 64, 196
        mem = torch.randn(2 * 8 * p, d, device=dev).to(dt)
        frames = torch.randn(40, p, d, device=dev).to(dt)
        fine = torch.tensor([0, 3, 9, 39], device=dev)
        emb = torch.randn(2, d, device=dev).to(dt)
        nl = torch.randn(d, device=dev).to(dt)
        tab = torch.randn(50000, d, device=dev).to(dt)
        pm = torch.tensor([1986, 374, 264, 1550, 11591, 12126, 315, 279, 2766, 25], device=dev)
        pf = torch.tensor([9485, 525, 48876, 9124, 14087, 504, 279, 2766, 25], device=dev)
        n = 10 + mem.shape[0] + 1 + 9 + 4 * p + 1
        seq = torch.zeros(n, d, device=dev, dtype=dt)
        ops.assemble(seq, mem, mem.shape[0], frames, fine, p, emb, nl, tab, pm, pf)
        ref = torch.cat([tab[pm].float(), mem.float() + emb[0].float(), nl[None].float(), tab[pf].float(),
                         (frames[fine].float() + emb[1].float()).reshape(-1, d), nl[None].float()])
        print(f"assemble {dt}: err {nerr(seq.float(), ref.to(dt).float()):.3e}")
    torch.cuda.synchronize()


def fam_simt():
    import torch
    from mavlm_b200 import ops
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda"
    for (m, n, k) in ((128, 64, 16), (1568, 896, 896), (300, 100, 52), (1, 8, 4), (1568, 32, 128)):
        a = torch.randn(m, k, device=dev)
        w = torch.randn(n, k, device=dev)
        b = torch.randn(n, device=dev)
        r = torch.randn(m, n, device=dev)
        av = torch.randn(n, device=dev)
        ref = (a.double() @ w.double().T + b.double())
        print(f"simt gemm {m}x{n}x{k} bias: err {nerr(ops.linear(a, w, b), ref):.3e}")
        ref2 = torch.nn.functional.gelu(ref) + r.double() + av.double()
        print(f"simt gemm {m}x{n}x{k} gelu+resid+addvec: err "
              f"{nerr(ops.linear(a, w, b, act=1, resid=r, addvec=av), ref2):.3e}")
        ref3 = torch.relu(ref)
        print(f"simt gemm {m}x{n}x{k} relu: err {nerr(ops.linear(a, w, b, act=2), ref3):.3e}")
    for (bsz, h, lq, lk, dh) in ((1, 8, 1568, 392, 4), (2, 8, 200, 333 * 4, 16), (1, 8, 1568, 6272, 112), (1, 2, 64, 64, 2)):
        q = torch.randn(bsz, lq, h * dh, device=dev)
        kk = torch.randn(bsz, lk, h * dh, device=dev)
        v = torch.randn(bsz, lk, h * dh, device=dev)
        o, lse, cs = ops.xattn(q, kk, v, h, want_lse=True, want_col_scores=True)
        qd, kd, vd = (t.double().view(bsz, -1, h, dh).transpose(1, 2) for t in (q, kk, v))
        s = qd @ kd.transpose(-1, -2) / math.sqrt(dh)
        pr = s.softmax(-1)
        ref = (pr @ vd).transpose(1, 2).reshape(bsz, lq, h * dh)
        print(f"simt xattn B{bsz} H{h} {lq}x{lk} dh{dh}: o err {nerr(o, ref):.3e} lse err "
              f"{nerr(lse, torch.logsumexp(s, -1)):.3e} colscore err {nerr(cs, pr.sum(1).sum(1)):.3e}")
    torch.cuda.synchronize()


def _gemm_case(m, n, k, bn, act=0, resid=False, addvec=False, f32out=False, identity=False, verbose=False):
    import torch
    from mavlm_b200 import ops, _lib
    dev = "cuda"
    _lib.load().mavlm_debug_force_gemm_bn(bn)
    a = torch.randn(m, k, device=dev).bfloat16()
    if identity:
        w = torch.eye(n, k, device=dev).bfloat16()
    else:
        w = (torch.randn(n, k, device=dev) / math.sqrt(k)).bfloat16()
    b = torch.randn(n, device=dev).bfloat16()
    r = torch.randn(m, n, device=dev).bfloat16() if resid else None
    av = torch.randn(n, device=dev).bfloat16() if addvec else None
    out = ops.linear(a, w, b, act=act, resid=r, addvec=av, out_dtype=torch.float32 if f32out else None)
    torch.cuda.synchronize()
    ref = a.double() @ w.double().T + b.double()
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    elif act == 2:
        ref = torch.relu(ref)
    if resid:
        ref = ref + r.double()
    if addvec:
        ref = ref + av.double()
    e = nerr(out.float(), ref)
    tag = f"tc gemm M{m} N{n} K{k} BN{bn} act{act} resid{int(resid)} addv{int(addvec)} f32{int(f32out)} id{int(identity)}"
    print(f"{tag}: err {e:.3e} {'OK' if e < 2e-2 else 'BAD'}")
    if e >= 2e-2 and verbose:
        d = (out.double() - ref).abs()
        bad = d > 0.05 * ref.abs().max()
        print("   bad fraction", float(bad.float().mean()), "bad rows", int(bad.any(1).sum()), "bad cols",
              int(bad.any(0).sum()))
        print("   out[0,:16]", [round(float(x), 3) for x in out[0, :16]])
        print("   ref[0,:16]", [round(float(x), 3) for x in ref[0, :16]])
        print("   out[1,:8] ", [round(float(x), 3) for x in out[1, :8]], " ref[1,:8]",
              [round(float(x), 3) for x in ref[1, :8]])
        cols = torch.nonzero(bad.any(0)).flatten()[:20].tolist()
        rows = torch.nonzero(bad.any(1)).flatten()[:20].tolist()
        print("   first bad cols", cols, "first bad rows", rows)
    return e


def fam_gemm():
    import torch
    torch.manual_seed(0)
    _gemm_case(128, 64, 64, 64, identity=True, verbose=True)
    _gemm_case(128, 64, 64, 64, verbose=True)
    for bn in (1128, 1192, 1256, 64, 128, 192, 256):
        w = bn % 1000
        _gemm_case(128, w, 64, bn, verbose=True)
        _gemm_case(256, w, 64, bn, identity=True, verbose=True)
        _gemm_case(128, w, 256, bn, verbose=True)
        _gemm_case(384, 2 * w, 512, bn, verbose=True)
        _gemm_case(2000, 5 * w + 8, 1160, bn, resid=True, verbose=True)
        _gemm_case(1568, 896, 896, bn, act=1, verbose=True)
        _gemm_case(1568, 3584, 3584, bn, act=2, verbose=True)
        _gemm_case(200, 3584, 1152, bn, resid=True, f32out=True, verbose=True)
        _gemm_case(1568, 1024, 896, bn, resid=True, addvec=True, verbose=True)
    _gemm_case(1568, 14336, 3584, 0, act=2, verbose=True)
    _gemm_case(1568, 3584, 14336, 0, resid=True, f32out=True, verbose=True)
    _gemm_case(46656, 3584, 1152, 0, act=1, verbose=True)
    _gemm_case(77, 40, 24, 0, verbose=True)     # N and K tails (N%8==0, K%8==0)
    from mavlm_b200 import _lib
    _lib.load().mavlm_debug_force_gemm_bn(0)


def _attn_case(bsz, h, lq, lk, dh, qscale=1.0, verbose=True, strided=False):
    import torch
    from mavlm_b200 import ops
    dev = "cuda"
    q = (torch.randn(bsz, lq, h * dh, device=dev) * qscale).bfloat16()
    if strided:   # k|v interleaved in one buffer like the fused projection output
        kv = torch.randn(bsz, lk, 2 * h * dh, device=dev).bfloat16()
        k, v = kv[..., :h * dh], kv[..., h * dh:]
    else:
        k = torch.randn(bsz, lk, h * dh, device=dev).bfloat16()
        v = torch.randn(bsz, lk, h * dh, device=dev).bfloat16()
    o, lse, _ = ops.xattn(q, k, v, h, want_lse=True)
    torch.cuda.synchronize()
    qd, kd, vd = (t.double().reshape(bsz, -1, h, dh).transpose(1, 2) for t in (q, k, v))
    s = qd @ kd.transpose(-1, -2) / math.sqrt(dh)
    pr = s.softmax(-1)
    ref = (pr @ vd).transpose(1, 2).reshape(bsz, lq, h * dh)
    e = nerr(o.float(), ref)
    el = nerr(lse, torch.logsumexp(s, -1))
    print(f"tc xattn B{bsz} H{h} {lq}x{lk} dh{dh} qscale{qscale} strided{int(strided)}: o err {e:.3e} lse err {el:.3e} "
          f"{'OK' if e < 2e-2 and el < 1e-2 else 'BAD'}")
    if e >= 2e-2 and verbose:
        d = (o.double() - ref).abs()
        bad = d > 0.05 * ref.abs().max()
        print("   bad fraction", float(bad.float().mean()))
        print("   o[0,0,:8]  ", [round(float(x), 4) for x in o[0, 0, :8]])
        print("   ref[0,0,:8]", [round(float(x), 4) for x in ref[0, 0, :8]])
        print("   o[0,0,64:72]  ", [round(float(x), 4) for x in o[0, 0, 64:72]])
        print("   ref[0,0,64:72]", [round(float(x), 4) for x in ref[0, 0, 64:72]])
        print("   bad per head", [float(bad[..., i * dh:(i + 1) * dh].float().mean()) for i in range(h)])
    return e


def fam_attn():
    import torch
    torch.manual_seed(0)
    for dh in (128, 448):
        _attn_case(1, 1, 128, 64, dh)
        _attn_case(1, 1, 128, 128, dh)
        _attn_case(1, 2, 256, 640, dh)
        _attn_case(1, 1, 128, 100, dh)             # key tail
        _attn_case(2, 8, 1568, 1568, dh, strided=True)
        _attn_case(1, 8, 1568, 6272, dh)
        _attn_case(1, 8, 1568, 6272, dh, qscale=8.0)   # sharp softmax: exercises the lazy O rescale
        _attn_case(1, 2, 200, 3 * 1568, dh, qscale=4.0)
        _attn_case(1, 8, 1568, 15680, dh, qscale=2.0)      # evolution-sized key set, split items
        _attn_case(3, 8, 1568, 3136, dh, strided=True)     # 312 items over 148 CTAs: ~2 items + partials each
        _attn_case(1, 8, 130, 64, dh)                       # 16 units < 148 CTAs, one key block each


def _build_models(d, dv, dtype, seed=0, chunk=32, frames=64, depth=2):
    """Random-init weights shared between the oracle (numpy) and the CUDA modules."""
    import numpy as np
    import torch
    import mavlm_b200 as M
    cfg = M.Config()
    cfg.mm_hidden_size = d
    cfg.mm_intermediate_size = 4 * d
    cfg.depth = depth
    cfg.mm_dtype = torch.float32
    torch.manual_seed(seed)
    rmt = M.TransformerProjector(cfg)
    fuser = M.build_memory_fuser(d)
    import types
    proj = M.build_vision_projector(types.SimpleNamespace(mm_projector_type="mlp2x_gelu", mm_hidden_size=dv, hidden_size=d))
    pe = M.TemporalPositionalEncoding(600, d, learnable=False)
    tte = torch.nn.Embedding(2, d)
    newline = torch.randn(d) * d ** -0.5
    emb = torch.nn.Embedding(50000, d)
    pipe = M.VisualMemoryPipeline(mm_projector=proj, recurrent_memory_transformer=rmt, memory_fuser=fuser,
                                  positional_encoding=pe, token_type_embedding=tte, image_newline=newline,
                                  embed_tokens=emb, chunk_size=chunk)
    w = {}
    for pref, mod in (("recurrent_memory_transformer.", rmt), ("memory_fuser.", fuser), ("mm_projector.", proj),
                      ("token_type_embedding.", tte)):
        for k_, v_ in mod.state_dict().items():
            w[pref + k_] = v_.detach().double().numpy()
    w["image_newline"] = newline.double().numpy()
    w["positional_encoding.frame_embed"] = pe.frame_embed.numpy()
    w["embed"] = emb.weight.detach().double().numpy()
    pipe = pipe.to("cuda")
    pipe.image_newline = newline.to("cuda")
    if dtype != torch.float32:
        pipe = pipe.to(dtype)
        pipe.image_newline = pipe.image_newline.to(dtype)
        pe.frame_embed = pe.frame_embed.float()
    return pipe, w


def fam_pipe():
    """Whole-path parity lives in tests/test_gpu_parity.py (the oracle is test infrastructure and is only imported
    from tests/, smoke() and bench.py); this family just runs the path once at each size and reports finiteness."""
    import torch
    torch.manual_seed(0)
    for (d, dv, dt, frames, chunk) in ((64, 16, torch.float32, 6, 2), (896, 1152, torch.float32, 8, 4),
                                       (896, 1152, torch.bfloat16, 64, 32), (3584, 1152, torch.bfloat16, 64, 32)):
        t0 = time.time()
        pipe, _ = _build_models(d, dv, dt, chunk=chunk)
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(1, frames, 729, dv, generator=g)
        idx = torch.arange(frames)[None]
        res = pipe(x.to("cuda").to(dt), idx)
        torch.cuda.synchronize()
        ok = bool(torch.isfinite(res["sequence"].float()).all())
        print(f"pipe D{d} {dt} F{frames} C{chunk}: sequence {tuple(res['sequence'].shape)} finite={ok} "
              f"[{time.time() - t0:.1f}s]  (parity: pytest tests/test_gpu_parity.py)", flush=True)


def fam_perf():
    import torch
    from mavlm_b200 import ops, _lib
    dev = "cuda"
    lib = _lib.load()

    def timeit(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    shapes = [(1568, 3584, 3584), (1568, 14336, 3584), (1568, 3584, 14336), (12544, 14336, 3584), (46656, 3584, 1152),
              (46656, 3584, 3584), (12544, 3584, 3584), (12544, 3584, 14336), (1568, 7168, 3584), (8192, 8192, 8192)]
    for (m, n, k) in shapes:
        a = torch.randn(m, k, device=dev).bfloat16()
        w = torch.randn(n, k, device=dev).bfloat16()
        b = torch.randn(n, device=dev).bfloat16()
        out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        line = f"gemm {m}x{n}x{k}:"
        for bn in (0, 128, 192, 256, 1128, 1192, 1256):
            lib.mavlm_debug_force_gemm_bn(bn)
            ms = timeit(lambda: ops.linear(a, w, b, out=out))
            line += f"  BN{bn}: {ms * 1e3:.0f}us {2 * m * n * k / ms / 1e9:.0f}TF"
        lib.mavlm_debug_force_gemm_bn(0)
        ms = timeit(lambda: torch.nn.functional.linear(a, w, b))
        line += f"  | cublas {ms * 1e3:.0f}us {2 * m * n * k / ms / 1e9:.0f}TF"
        print(line, flush=True)
    for (bsz, lq, lk, dh) in ((1, 1568, 6272, 448), (8, 1568, 6272, 448), (1, 1568, 15680, 448), (1, 1568, 6272, 128),
                              (8, 1568, 3136, 448)):
        h = 8
        q = torch.randn(bsz, lq, h * dh, device=dev).bfloat16()
        k = torch.randn(bsz, lk, h * dh, device=dev).bfloat16()
        v = torch.randn(bsz, lk, h * dh, device=dev).bfloat16()
        ms = timeit(lambda: ops.xattn(q, k, v, h))
        fl = 4.0 * bsz * h * lq * lk * dh
        print(f"xattn B{bsz} {lq}x{lk} dh{dh}: {ms * 1e3:.0f}us {fl / ms / 1e9:.0f}TF", flush=True)
    h = 8
    for (bsz, lq, lk, dh) in ((1, 1568, 6272, 448), (8, 1568, 6272, 448)):
        q = torch.randn(bsz, lq, h * dh, device=dev).bfloat16()
        k = torch.randn(bsz, lk, h * dh, device=dev).bfloat16()
        v = torch.randn(bsz, lk, h * dh, device=dev).bfloat16()
        line = f"xattn B{bsz} groups sweep:"
        for g in (4, 6, 8, 9, 10, 11):
            lib.mavlm_debug_force_attn_groups(g)
            ms = timeit(lambda: ops.xattn(q, k, v, h))
            line += f"  g{g}: {ms * 1e3:.0f}us"
        lib.mavlm_debug_force_attn_groups(0)
        print(line, flush=True)
    for frames in (64, 256):
        x = torch.randn(frames, 729, 3584, device=dev).bfloat16()
        table = torch.randn(600, 3584, device=dev)
        idx = torch.arange(frames, device=dev)
        ms = timeit(lambda: ops.pool_pe(x, side=27, pe_table=table, frame_idx=idx))
        by = frames * (729 + 196) * 3584 * 2
        print(f"pool_pe F{frames}: {ms * 1e3:.0f}us {by / ms / 1e6:.0f}GB/s", flush=True)
    xl = torch.randn(1568, 3584, device=dev)
    g = torch.randn(3584, device=dev).bfloat16()
    ms = timeit(lambda: ops.layernorm(xl, g, g, 1e-12, out_dtype=torch.bfloat16))
    print(f"layernorm 1568x3584 f32->bf16: {ms * 1e3:.1f}us {1568 * 3584 * 6 / ms / 1e6:.0f}GB/s")


def fam_micro():
    """Per-kernel device time without host launch overhead: N back-to-back calls captured in one CUDA graph."""
    import torch
    from mavlm_b200 import ops
    dev = "cuda"

    def graph_time(fn, n=20, reps=5):
        fn()
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (n * reps) * 1e3   # us per call

    xl = torch.randn(1568, 3584, device=dev)
    gm = torch.randn(3584, device=dev).bfloat16()
    yl = torch.empty(1568, 3584, device=dev, dtype=torch.bfloat16)
    print(f"layernorm 1568x3584 f32->bf16 (L2-resident input): {graph_time(lambda: ops.layernorm(xl, gm, gm, 1e-12, out=yl)):.1f} us")
    xs = [torch.randn(1568, 3584, device=dev) for _ in range(12)]    # 270 MB ring: input comes from HBM
    it = [0]

    def ln_ring():
        it[0] = (it[0] + 1) % len(xs)
        ops.layernorm(xs[it[0]], gm, gm, 1e-12, out=yl)
    print(f"layernorm 1568x3584 f32->bf16 (HBM input): {graph_time(ln_ring, n=24):.1f} us  (ideal 33.7 MB / 6.55 TB/s = 5.1 us)")
    tiny = torch.randn(8, 196, 64, device=dev).bfloat16()
    tab = torch.randn(8, 64, device=dev)
    idx = torch.arange(8, device=dev)
    print(f"tiny add_pe (launch-to-launch floor): {graph_time(lambda: ops.add_pe(tiny, tab, idx, out=tiny)):.1f} us")
    from mavlm_b200 import _lib as _l
    for (m, n, k) in ((1568, 3584, 3584), (1568, 14336, 3584), (1568, 3584, 14336), (1568, 7168, 3584), (3136, 3584, 14336)):
        a = torch.randn(m, k, device=dev).bfloat16()
        w = torch.randn(n, k, device=dev).bfloat16()
        b = torch.randn(n, device=dev).bfloat16()
        out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        us = graph_time(lambda: ops.linear(a, w, b, out=out))
        print(f"gemm {m}x{n}x{k}: {us:.1f} us {2 * m * n * k / us / 1e6:.0f} TF")
    from mavlm_b200 import _lib
    lib = _lib.load()
    h, dh = 8, 448
    for (bsz, lq, lk) in ((1, 1568, 6272), (1, 1568, 1568), (8, 1568, 6272), (1, 1568, 15680)):
        q = torch.randn(bsz, lq, h * dh, device=dev).bfloat16()
        kk = torch.randn(bsz, lk, h * dh, device=dev).bfloat16()
        v = torch.randn(bsz, lk, h * dh, device=dev).bfloat16()
        us = graph_time(lambda: ops.xattn(q, kk, v, h), n=10)
        print(f"xattn B{bsz} {lq}x{lk} dh{dh}: {us:.1f} us {4.0 * bsz * h * lq * lk * dh / us / 1e6:.0f} TF")
    import mavlm_b200 as M
    for (f, hh, ww) in ((64, 720, 1280), (64, 360, 640)):
        fr = torch.randint(0, 256, (f, hh, ww, 3), dtype=torch.uint8, device=dev)
        us = graph_time(lambda: M.frames_preprocess(fr, dtype=torch.bfloat16), n=3)
        by = f * 3 * (hh * ww + 2 * hh * 384) + f * 3 * 384 * 384 * 2
        print(f"frames_preprocess {f}x{hh}x{ww} u8 -> bf16 pixel_values: {us:.1f} us ({us / f:.2f} us/frame) {by / us / 1e3:.0f} GB/s")
    x = torch.randn(64, 729, 3584, device=dev).bfloat16()
    us = graph_time(lambda: ops.pool_pe(x, side=27), n=5)
    print(f"pool F64: {us:.1f} us {64 * (729 + 196) * 3584 * 2 / us / 1e3:.0f} GB/s")


def fam_ab():
    """A/B of scheduling knobs on the full bench step, alternating within one process (box-to-box and
    thermal variation between separate runs is +-5 %)."""
    import torch
    from mavlm_b200 import _lib, synthetic
    lib = _lib.load()
    pipe, _ = synthetic.build_pipeline(3584, 1152, dtype=torch.bfloat16, chunk_size=32, device="cuda:0")
    x = synthetic.synthetic_tower_tokens(1, 64).to("cuda:0")
    idx = torch.arange(64, device="cuda:0")[None]

    def run(n=10):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            pipe(x, idx, validate=False, return_states=False)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    run(5)
    from mavlm_b200 import GraphedPipeline

    def graph_ms(gp, n=10):
        gp(x, idx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            gp(None, None)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    variants = {}
    for name, flags in (("pdl", 0), ("no-pdl", 16)):
        lib.mavlm_debug_set_flags(flags)
        variants[name] = GraphedPipeline(pipe, 1, 64)       # the launch attributes are baked in at capture time
    lib.mavlm_debug_set_flags(0)
    names = list(variants)
    for rnd in range(6):                                    # rotate the order and idle before every measurement: the
        order = names[rnd % len(names):] + names[:rnd % len(names)]   # power cap bites ~0.1 s into a burst
        res = {}
        for name in order:
            time.sleep(0.5)
            res[name] = graph_ms(variants[name], 20)
        print(f"round {rnd} (order {order}):  " + "   ".join(f"{name}: {res[name]:.3f} ms" for name in names), flush=True)
    ref = variants["no-pdl"](x, idx)["sequence"].float().clone()
    out = variants["pdl"](x, idx)["sequence"].float()
    torch.cuda.synchronize()
    print("pdl vs no-pdl max abs diff:", float((out - ref).abs().max()))


def fam_lnperf():
    """LayerNorm / assembly throughput, timed as CUDA-graph replays (no Python between launches), with a parity check
    against torch at rows that make a warp loop (rows > 148 * 12)."""
    import torch
    import torch.nn.functional as F
    from mavlm_b200 import ops
    dev = "cuda"
    torch.manual_seed(0)

    def graph_us(fn, reps=20, n=5):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
        return best

    for d, idt, odt in ((3584, torch.float32, torch.bfloat16), (3584, torch.bfloat16, torch.bfloat16),
                        (896, torch.float32, torch.bfloat16), (4096, torch.float32, torch.float32), (64, torch.float32, torch.float32)):
        for rows in (1, 77, 1568, 1777, 5000, 12544, 50176):
            x = (torch.randn(rows, d, device=dev) * 3 + 1).to(idt)
            g = torch.randn(d, device=dev).to(odt)
            b = torch.randn(d, device=dev).to(odt)
            out = torch.empty(rows, d, device=dev, dtype=odt)
            ops.layernorm(x, g, b, 1e-12, out_dtype=odt, out=out)
            ref = F.layer_norm(x.double(), (d,), g.double(), b.double(), 1e-12)
            err = nerr(out.float(), ref)
            us = graph_us(lambda: ops.layernorm(x, g, b, 1e-12, out_dtype=odt, out=out))
            by = rows * d * (x.element_size() + out.element_size())
            print(f"layernorm {idt}->{odt} d={d} rows={rows}: err {err:.2e}  {us:.1f} us  {by / us / 1e3:.0f} GB/s", flush=True)
    # assembly at the bench shape: 2 states in place, 32 fine frames
    d, p = 3584, 196
    dt = torch.bfloat16
    frames = torch.randn(64, p, d, device=dev).to(dt)
    fine = torch.arange(0, 64, 2, device=dev)
    emb = torch.randn(2, d, device=dev).to(dt)
    nl = torch.randn(d, device=dev).to(dt)
    tab = torch.randn(50000, d, device=dev).to(dt)
    pm = torch.tensor([1986, 374, 264, 1550, 11591, 12126, 315, 279, 2766, 25], device=dev)
    pf = torch.tensor([9485, 525, 48876, 9124, 14087, 504, 279, 2766, 25], device=dev)
    n_mem = 2 * 1568
    n = 10 + n_mem + 1 + 9 + 32 * p + 1
    seq = torch.zeros(n, d, device=dev, dtype=dt)
    us = graph_us(lambda: ops.assemble(seq, None, n_mem, frames, fine, p, emb, nl, tab, pm, pf))
    rows = n - n_mem
    print(f"assemble (mem in place) {rows} rows: {us:.1f} us  {2 * rows * d * 2 / us / 1e3:.0f} GB/s")
    ref = torch.cat([tab[pm].float(), torch.zeros(n_mem, d, device=dev), nl[None].float(), tab[pf].float(),
                     (frames[fine].float() + emb[1].float()).reshape(-1, d), nl[None].float()]).to(dt)
    print("assemble equal:", bool((seq == ref).all()))


FAMS = {"micro": fam_micro, "ab": fam_ab, "elem": fam_elem, "simt": fam_simt, "gemm": fam_gemm, "attn": fam_attn, "pipe": fam_pipe, "perf": fam_perf, "lnperf": fam_lnperf}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        FAMS[sys.argv[2]]()
        sys.exit(0)
    os.makedirs(OUT, exist_ok=True)
    fams = sys.argv[1:] or list(FAMS)
    for f in fams:
        log = os.path.join(OUT, f"diag_{f}.log")
        t0 = time.time()
        with open(log, "w") as fh:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", f], stdout=fh,
                                   stderr=subprocess.STDOUT, timeout=int(os.environ.get("DIAG_TIMEOUT", "420")))
                rc = r.returncode
            except subprocess.TimeoutExpired:
                rc = "TIMEOUT"
        print(f"=== {f}: rc={rc} ({time.time() - t0:.0f}s)")
        with open(log) as fh:
            txt = fh.read()
        print(txt[-6000:])
