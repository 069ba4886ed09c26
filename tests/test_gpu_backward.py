"""Backward / training parity (BASELINE config[3] shape class): gradients of the CUDA path (autograd
Functions whose forward AND backward run in libmavlm.so) against golden gradients produced by the
unmodified reference under PyTorch autograd in float64 (tools/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

import mavlm_b200 as M
from mavlm_b200 import ops, synthetic
from oracle import vismem_oracle as O

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GRAD_TOL = 2e-4        # fp32 tier vs float64 reference autograd, err = max|a-b| / max|b| per parameter


def err(a, b, floor=0.0):
    """max|a-b| / max(max|b|, floor).  `floor` guards parameters whose true gradient is exactly zero (k_proj.bias:
    softmax is invariant to a per-query constant), where the reference holds 1e-20 noise and fp32 holds 1e-10."""
    a = a.detach().double().cpu().numpy()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), floor, 1e-30))


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    w = {k[3:]: z[k] for k in z.files if k.startswith("w::")}
    g = {k[3:]: z[k] for k in z.files if k.startswith("g::")}
    return z, w, g


def _rmt(w, d, prefix="recurrent_memory_transformer."):
    cfg = M.Config()
    cfg.mm_hidden_size, cfg.mm_intermediate_size, cfg.depth, cfg.mm_dtype = d, 4 * d, 2, torch.float32
    rmt = M.TransformerProjector(cfg)
    rmt.load_state_dict({k[len(prefix):]: torch.from_numpy(v).float() for k, v in w.items() if k.startswith(prefix)})
    return rmt.to(DEV)


def test_op_level_gradients_match_torch_autograd():
    """Each autograd.Function against torch's own autograd of the same math (fp32)."""
    torch.manual_seed(0)
    x = torch.randn(70, 48, device=DEV, requires_grad=True)
    w = torch.randn(36, 48, device=DEV, requires_grad=True)
    b = torch.randn(36, device=DEV, requires_grad=True)
    r = torch.randn(70, 36, device=DEV, requires_grad=True)
    av = torch.randn(36, device=DEV, requires_grad=True)
    for act, fn in ((0, lambda t: t), (2, torch.relu), (1, torch.nn.functional.gelu)):
        use_r = act == 0
        y = ops.linear(x, w, b, act=act, resid=r if use_r else None, addvec=av if use_r else None)
        ref = fn(x @ w.T + b) + (r + av if use_r else 0)
        go = torch.randn_like(ref)
        got = torch.autograd.grad(y, [x, w, b] + ([r, av] if use_r else []), go)
        exp = torch.autograd.grad(ref, [x, w, b] + ([r, av] if use_r else []), go)
        for a_, e_ in zip(got, exp):
            assert err(a_, e_.double().cpu().numpy()) < 1e-5, act
    pre = torch.randn(33, 64, device=DEV, requires_grad=True)
    g = torch.randn(64, device=DEV, requires_grad=True)
    be = torch.randn(64, device=DEV, requires_grad=True)
    y = ops.layernorm(pre, g, be, 1e-12)
    ref = torch.nn.functional.layer_norm(pre, (64,), g, be, 1e-12)
    go = torch.randn_like(ref)
    for a_, e_ in zip(torch.autograd.grad(y, [pre, g, be], go), torch.autograd.grad(ref, [pre, g, be], go)):
        assert err(a_, e_.double().cpu().numpy()) < 1e-5
    h, dh = 4, 8
    q = torch.randn(2, 50, h * dh, device=DEV, requires_grad=True)
    kv = torch.randn(2, 77, 2 * h * dh, device=DEV, requires_grad=True)
    o, _, _ = ops.xattn(q, kv[..., :h * dh], kv[..., h * dh:], h)
    qh = q.view(2, 50, h, dh).transpose(1, 2)
    kh = kv[..., :h * dh].reshape(2, 77, h, dh).transpose(1, 2)
    vh = kv[..., h * dh:].reshape(2, 77, h, dh).transpose(1, 2)
    ref = ((qh @ kh.transpose(-1, -2) / dh ** 0.5).softmax(-1) @ vh).transpose(1, 2).reshape(2, 50, h * dh)
    go = torch.randn_like(ref)
    assert err(o, ref.detach().double().cpu().numpy()) < 1e-5
    for a_, e_ in zip(torch.autograd.grad(o, [q, kv], go), torch.autograd.grad(ref, [q, kv], go)):
        assert err(a_, e_.double().cpu().numpy()) < 1e-5


@pytest.mark.parametrize("d,rows", [(896, 700), (3584, 1568), (3584, 37), (4096, 300), (2056, 333)])
def test_layernorm_backward_wide_rows_match_torch_autograd(d, rows):
    """LayerNorm backward at the model's widths (the persistent vectorised kernel: rows > 2 x SMs make a CTA walk
    several rows with the next row prefetched; d = 2056 leaves a ragged last vector group): d(pre), d(gamma), d(beta)
    against torch's autograd of F.layer_norm, fp32 exactly and the bf16 tier against an fp64 evaluation."""
    torch.manual_seed(d + rows)
    pre = (torch.randn(rows, d, device=DEV) * 2 + 0.5).requires_grad_(True)
    g = torch.randn(d, device=DEV, requires_grad=True)
    be = torch.randn(d, device=DEV, requires_grad=True)
    go = torch.randn(rows, d, device=DEV)
    y = ops.layernorm(pre, g, be, 1e-12)
    ref = torch.nn.functional.layer_norm(pre.double(), (d,), g.double(), be.double(), 1e-12)
    exp = torch.autograd.grad(ref, [pre, g, be], go.double())
    for a_, e_ in zip(torch.autograd.grad(y, [pre, g, be], go), exp):
        assert err(a_, e_.cpu().numpy()) < 1e-5
    # bf16 tier: fp32 pre-LN sum in, bf16 gamma / beta / dy (what the training step runs)
    g16 = g.detach().bfloat16().requires_grad_(True)
    b16 = be.detach().bfloat16().requires_grad_(True)
    y16 = ops.layernorm(pre, g16, b16, 1e-12, out_dtype=torch.bfloat16)
    go16 = go.bfloat16()
    g64 = g16.detach().double().requires_grad_(True)
    b64 = b16.detach().double().requires_grad_(True)
    ref16 = torch.nn.functional.layer_norm(pre.double(), (d,), g64, b64, 1e-12)
    exp16 = torch.autograd.grad(ref16, [pre, g64, b64], go16.double())
    for a_, e_ in zip(torch.autograd.grad(y16, [pre, g16, b16], go16), exp16):
        assert err(a_, e_.cpu().numpy()) < 8e-3


def test_rmt_bptt_gradients_against_reference_golden():
    """2 chunks (formation x2 + evolution): every RMT parameter's gradient vs the reference's autograd."""
    z, w, g = _load("rmt_grads.npz")
    rmt = _rmt(w, 16)
    frames = torch.from_numpy(z["frames"]).float().to(DEV)
    rmt.memory_cache = []
    for i in range(2):
        cache, _ = rmt(frames[2 * i:2 * i + 2])
    loss = sum((s * s).mean() for s in cache)
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    loss.backward()
    n = 0
    gmax = max(np.abs(v).max() for v in g.values())
    for name, p in rmt.named_parameters():
        ref = g["recurrent_memory_transformer." + name]
        assert p.grad is not None, name
        floor = gmax * (1e-2 if name.endswith("k_proj.bias") else 1e-6)     # k bias: true gradient is exactly 0
        assert err(p.grad, ref, floor) < GRAD_TOL, (name, err(p.grad, ref, floor))
        n += 1
    assert n == len(g)


def test_whole_path_gradients_against_reference_golden():
    """loss = mean(sequence^2) through memory (3 chunks) + fuser + type embeddings + newline + prompt embeddings."""
    z, w, g = _load("path_grads.npz")
    d = 16
    pipe, _ = synthetic.build_pipeline(d, 4, dtype=torch.float32, chunk_size=2, device=DEV, vocab=50000)
    pipe.recurrent_memory_transformer.load_state_dict(
        {k[len("recurrent_memory_transformer."):]: torch.from_numpy(v).float() for k, v in w.items()
         if k.startswith("recurrent_memory_transformer.")})
    pipe.memory_fuser.load_state_dict({k[len("memory_fuser."):]: torch.from_numpy(v).float() for k, v in w.items()
                                       if k.startswith("memory_fuser.")})
    pipe.token_type_embedding.load_state_dict({"weight": torch.from_numpy(w["token_type_embedding.weight"]).float()})
    pipe.image_newline = torch.nn.Parameter(torch.from_numpy(w["image_newline"]).float().to(DEV))
    rows = torch.from_numpy(z["embed_rows"])
    with torch.no_grad():
        pipe.embed_tokens.weight[rows] = torch.from_numpy(w["embed_rows"]).float().to(DEV)
    zz = torch.from_numpy(z["z"]).float().to(DEV)[None]
    out = pipe.memory_forward_train(zz)
    assert err(out["sequence"][0], z["sequence"]) < 2e-5
    loss = (out["sequence"] ** 2).mean()
    loss.backward()
    gmax = max(np.abs(v).max() for v in g.values())
    floor = 1e-6 * gmax
    for pref, mod in (("recurrent_memory_transformer.", pipe.recurrent_memory_transformer),
                      ("memory_fuser.", pipe.memory_fuser), ("token_type_embedding.", pipe.token_type_embedding)):
        for name, p in mod.named_parameters():
            fl = gmax * 1e-2 if name.endswith("k_proj.bias") else floor      # k bias: true gradient is exactly 0
            assert err(p.grad, g[pref + name], fl) < GRAD_TOL, (pref + name, err(p.grad, g[pref + name], fl))
    assert err(pipe.image_newline.grad, g["image_newline"], floor) < GRAD_TOL
    assert err(pipe.embed_tokens.weight.grad[rows.to(DEV)], g["embed_rows"], floor) < GRAD_TOL
    # inference path == training path forward
    with torch.no_grad():
        inf = pipe.memory_forward(zz)
    assert err(inf["sequence"][0], out["sequence"][0].detach().double().cpu().numpy()) < 1e-5


def test_single_chunk_leaves_evolution_attention_without_gradient():
    """SURVEY.md §3.2 (probed on the reference): with one chunk memory_update_attention gets no gradient."""
    z, w, _ = _load("rmt_grads.npz")
    rmt = _rmt(w, 16)
    rmt.memory_cache = []
    cache, _ = rmt(torch.from_numpy(z["frames"]).float().to(DEV))
    (cache[-1] ** 2).mean().backward()
    assert all(p.grad is None for p in rmt.memory_update_attention.parameters())
    assert all(p.grad is not None for p in rmt.layers.parameters())
    assert rmt.initial_memory.grad is not None and rmt.memory_pos_embed.grad is not None


# ------------------------------------------------------------------------------------------------
# bf16 tensor-core tier
# ------------------------------------------------------------------------------------------------
def _nerr(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_gemm_ex_bf16_all_operand_layouts():
    """tcgen05 GEMM with K-major / MN-major operands (dgrad / wgrad / attention-backward layouts), batched."""
    from mavlm_b200 import autograd as ag
    torch.manual_seed(0)
    m, n, k = 328, 200, 264
    a = torch.randn(m, k, device=DEV).bfloat16()
    b = (torch.randn(n, k, device=DEV) / 16).bfloat16()
    ref = a.double() @ b.double().T
    for ta in (False, True):
        for tb in (True, False):
            aa = a.t().contiguous() if ta else a            # stored [K, M] when trans_a
            bb = b if tb else b.t().contiguous()            # stored [N, K] when trans_b, else [K, N]
            out = ag.gemm_ex(aa, ta, bb, tb, m, n, k)
            assert _nerr(out, ref) < 1e-2, (ta, tb)
            o32 = ag.gemm_ex(aa, ta, bb, tb, m, n, k, out_dtype=torch.float32)
            assert _nerr(o32, ref) < 1e-5, (ta, tb)
            acc = o32.clone()
            ag.gemm_ex(aa, ta, bb, tb, m, n, k, out=acc, accumulate=True)
            assert _nerr(acc, 2 * ref) < 1e-5, (ta, tb)
    # fp16: the inference layouts (A K-major; B [N, K] or [K, N]); a transposed A is refused
    ah, bh = a.half(), b.half()
    refh = ah.double() @ bh.double().T
    for tb in (True, False):
        bb = bh if tb else bh.t().contiguous()
        out = ag.gemm_ex(ah, False, bb, tb, m, n, k)
        assert out.dtype == torch.float16 and _nerr(out, refh) < 2e-3, tb
        o32 = ag.gemm_ex(ah, False, bb, tb, m, n, k, out_dtype=torch.float32)
        assert _nerr(o32, refh) < 1e-5, tb
        acc = o32.clone()
        ag.gemm_ex(ah, False, bb, tb, m, n, k, out=acc, accumulate=True)
        assert _nerr(acc, 2 * refh) < 1e-5, tb
    with pytest.raises(Exception, match="fp16"):
        ag.gemm_ex(ah.t().contiguous(), True, bh, True, m, n, k)


def test_op_level_gradients_bf16():
    torch.manual_seed(0)
    x = torch.randn(1568, 896, device=DEV).bfloat16().requires_grad_()
    w = (torch.randn(3584, 896, device=DEV) / 30).bfloat16().requires_grad_()
    b = torch.randn(3584, device=DEV).bfloat16().requires_grad_()
    y = ops.linear(x, w, b, act=2)
    ref = torch.relu(x.float() @ w.float().T + b.float())
    go = torch.randn_like(ref).bfloat16()
    got = torch.autograd.grad(y, [x, w, b], go)
    exp = torch.autograd.grad(ref, [x, w, b], go.float())
    for a_, e_ in zip(got, exp):
        assert a_.dtype == torch.bfloat16 and _nerr(a_, e_) < 2e-2
    for dh in (128, 448):
        h = 8
        q = torch.randn(1, 1568, h * dh, device=DEV).bfloat16().requires_grad_()
        kv = torch.randn(1, 784, 2 * h * dh, device=DEV).bfloat16().requires_grad_()
        o, _, _ = ops.xattn(q, kv[..., :h * dh], kv[..., h * dh:], h)
        qh = q.float().view(1, -1, h, dh).transpose(1, 2)
        kh = kv.float()[..., :h * dh].reshape(1, -1, h, dh).transpose(1, 2)
        vh = kv.float()[..., h * dh:].reshape(1, -1, h, dh).transpose(1, 2)
        ref = ((qh @ kh.transpose(-1, -2) / dh ** 0.5).softmax(-1) @ vh).transpose(1, 2).reshape(1, -1, h * dh)
        go = torch.randn_like(ref).bfloat16()
        got = torch.autograd.grad(o, [q, kv], go)
        exp = torch.autograd.grad(ref, [q, kv], go.float())
        for a_, e_ in zip(got, exp):
            assert _nerr(a_, e_) < 3e-2, dh


def test_bf16_training_step_matches_fp32_tier():
    """OV-0.5B dims, 2 chunks (BPTT + evolution), loss = mean(sequence^2): bf16 tensor-core gradients against the
    fp32 tier's (which is pinned to the reference's autograd above)."""
    grads = {}
    for dt in (torch.float32, torch.bfloat16):
        pipe, _ = synthetic.build_pipeline(896, 1152, dtype=dt, chunk_size=4, device=DEV)
        g = torch.Generator().manual_seed(5)
        z = torch.randn(1, 8, 196, 896, generator=g).bfloat16().to(DEV).to(dt)
        out = pipe.memory_forward_train(z)
        (out["sequence"].float() ** 2).mean().backward()
        grads[dt] = {n: p.grad.detach().float().clone() for n, p in pipe.named_parameters() if p.grad is not None}
    assert set(grads[torch.float32]) == set(grads[torch.bfloat16])
    gmax = max(float(v.abs().max()) for v in grads[torch.float32].values())
    worst = 0.0
    for n, g32 in grads[torch.float32].items():
        g16 = grads[torch.bfloat16][n]
        e = float((g16 - g32).abs().max() / max(float(g32.abs().max()), 1e-3 * gmax))
        worst = max(worst, e)
        assert e < 0.1, (n, e)
    print("worst bf16-vs-fp32 gradient error", worst)


def test_training_step_as_one_cuda_graph_matches_eager_autograd():
    """GraphedTrainStep (forward + loss + BPTT backward captured once): same loss and gradients as the eager autograd
    step (bias / LayerNorm sums use atomics: not bitwise), new inputs take effect on replay, gradients are
    overwritten, not accumulated."""
    pipe, _ = synthetic.build_pipeline(896, 1152, dtype=torch.bfloat16, chunk_size=4, device=DEV)
    g = torch.Generator().manual_seed(5)
    z1 = torch.randn(2, 8, 196, 896, generator=g).bfloat16().to(DEV)
    z2 = (z1.float() * 0.5).bfloat16()

    def eager(z):
        for p_ in pipe.parameters():
            p_.grad = None
        out = pipe.memory_forward_train(z)
        loss = (out["sequence"].float() ** 2).mean()
        loss.backward()
        return float(loss.detach()), {n: p_.grad.detach().float().clone() for n, p_ in pipe.named_parameters() if p_.grad is not None}

    l1, g1 = eager(z1)
    l2, g2 = eager(z2)
    step = pipe.graphed_train(2, 8)
    for z, l_ref, g_ref in ((z1, l1, g1), (z2, l2, g2), (z1, l1, g1)):
        loss, seq = step(z)
        torch.cuda.synchronize()
        assert abs(float(loss.detach()) - l_ref) < 1e-3 * abs(l_ref)
        got = {n: p_.grad for n, p_ in pipe.named_parameters() if p_.grad is not None}
        assert set(got) == set(g_ref)
        gmax = max(float(v.abs().max()) for v in g_ref.values())
        for n, r in g_ref.items():
            e = float((got[n].float() - r).abs().max() / max(float(r.abs().max()), 1e-3 * gmax))
            assert e < 2e-2, (n, e)
    assert abs(l1 - l2) > 1e-3 * abs(l1)
