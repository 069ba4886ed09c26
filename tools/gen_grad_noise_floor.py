#!/usr/bin/env python
"""The REFERENCE's own bf16-vs-fp32 GRADIENT deviation at OV-7B dims (tests/golden/grad_noise_floor.json).

BASELINE.json states a bf16 tolerance (2e-2 of the tensor's max) for the fused output tokens and the final memory
state; for config 4 it asks for "gradients vs reference within tolerance".  This script measures what the UNMODIFIED
reference modules themselves deliver: `TransformerProjector` (MemoryController.py) + the fuser MLP + type embeddings +
assembly glue (llava_arch.py:132-136, 545-557, 620-629, 708-731) under PyTorch autograd, once in fp32 and once the way
the reference trains -- everything `.to(torch.bfloat16)` (finetune_short.sh:71) -- on the SAME bf16-rounded weights and
inputs the GPU tests use (tests/test_gpu_configs.py::_grad_parity_7b: synthetic.build_pipeline seed 0, z seeds 5 / 6).

Finding (8 host threads, torch 2.11 CPU, AMX bf16): the reference's bf16 gradients deviate from its fp32 gradients by
up to 20-40 % of a tensor's max (mlp.0.weight), ~10 % for the LayerNorm affine parameters, while all gradients
together agree to 0.6 % in relative L2 (cosine 0.99999).  The 196 tokens of a memory slot are near-identical at t = 0
(initial_memory std 3e-3 against memory_pos_embed std 1), so weight gradients are sums of strongly cancelling
per-token terms: bf16 rounding of the activations / activation gradients does not cancel, the signal does.  No bf16
implementation that rounds activations meets 2e-2 per tensor; the GPU tests therefore hold every tensor to
max(2e-2, the reference's own deviation) and the whole gradient to the relative-L2 / cosine bars recorded here.

    python tools/gen_grad_noise_floor.py          # needs /root/reference (or baseline/_ref); ~3 min on 8 cores
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "grad_noise_floor.json")

MEMORY_PROMPT_IDS = (1986, 374, 264, 1550, 11591, 12126, 315, 279, 2766, 25)      # llava_arch.py:708
FRAME_PROMPT_IDS = (9485, 525, 48876, 9124, 14087, 504, 279, 2766, 25)            # llava_arch.py:714


def build_reference(arch, weights, hidden, dtype):
    c = arch.Config()
    c.mm_hidden_size, c.mm_hidden_act, c.mm_num_attention_heads, c.patch_size = hidden, "relu", 8, 196
    c.mm_layer_norm_eps, c.mm_intermediate_size, c.num_memory_tokens, c.depth, c.mm_dtype = 1e-12, 4 * hidden, 8, 2, torch.float32
    rmt = arch.TransformerProjector(c)
    fuser = torch.nn.Sequential(torch.nn.Linear(hidden, 4 * hidden), torch.nn.GELU(), torch.nn.Linear(4 * hidden, hidden))
    tte = torch.nn.Embedding(2, hidden)
    mods = {"recurrent_memory_transformer.": rmt, "memory_fuser.": fuser, "token_type_embedding.": tte}
    for pref, mod in mods.items():
        mod.load_state_dict({k[len(pref):]: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in weights.items()
                             if k.startswith(pref)}, strict=True)
        mod.to(dtype)
    newline = torch.nn.Parameter(torch.from_numpy(weights["image_newline"]).to(dtype))
    tab = weights["embed_tokens.weight"]
    pm = torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(tab[list(MEMORY_PROMPT_IDS)])).to(dtype))
    pf = torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(tab[list(FRAME_PROMPT_IDS)])).to(dtype))
    return rmt, fuser, tte, newline, pm, pf


def reference_gradients(arch, weights, z, chunk, dtype):
    """loss = mean(sequence^2) over the videos of z [B, F, P, D], reference modules + the reference's glue."""
    hidden = z.shape[-1]
    rmt, fuser, tte, newline, pm, pf = build_reference(arch, weights, hidden, dtype)
    z = z.to(dtype)
    total, n_el = None, 0
    for b in range(z.shape[0]):
        image = z[b]
        f = image.shape[0]
        fine_idx = torch.clamp(torch.round(torch.linspace(0, f - 1, steps=min(32, f))).long(), 0, f - 1)   # :513-522
        fine = image[fine_idx]
        rmt.memory_cache = []                                                                              # :532
        bounds = list(range(0, f, chunk)) + [f]
        for b0, b1 in zip(bounds, bounds[1:]):
            cache, _ = rmt(image[b0:b1])                                                                   # :534-537
        mem = fuser(torch.cat(cache, dim=0))                                                               # :545-546
        mem = mem + tte(torch.zeros((mem.shape[0], 196), dtype=torch.long))                                # :548-551
        fine = fine + tte(torch.ones((fine.shape[0], 196), dtype=torch.long))                              # :552-554
        seq = torch.cat([pm, mem.flatten(0, 1), newline[None], pf, fine.flatten(0, 1), newline[None]], dim=0)
        s = (seq.float() ** 2).sum()
        total = s if total is None else total + s
        n_el += seq.numel()
    loss = total / n_el
    loss.backward()
    out = {}
    for pref, mod in (("recurrent_memory_transformer.", rmt), ("memory_fuser.", fuser), ("token_type_embedding.", tte)):
        for n, p in mod.named_parameters():
            if p.grad is not None:
                out[pref + n] = p.grad.detach().double().numpy()
    out["image_newline"] = newline.grad.detach().double().numpy()
    out["embed.prompt_mem"] = pm.grad.detach().double().numpy()
    out["embed.prompt_frm"] = pf.grad.detach().double().numpy()
    return float(loss.detach()), out


def main():
    from baseline import ref_arm
    from mavlm_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    root = ref_arm.find_reference_root()
    if root is None:
        raise SystemExit("no reference install (/root/reference or baseline/_ref)")
    arch, _ = ref_arm.load_reference(root)
    hidden = 3584
    _, w = synthetic.build_pipeline(hidden, 1152, dtype=torch.float32, chunk_size=32, device="cpu", vocab=50000)
    wq = synthetic.round_weights_like(w, torch.bfloat16)
    del w
    result = {"how": "tools/gen_grad_noise_floor.py: UNMODIFIED reference modules, torch CPU autograd, bf16 (.to(bfloat16), the "
                     "reference's training dtype) vs fp32, same bf16-rounded weights and inputs",
              "metric": "max|g_bf16 - g_fp32| / max|g_fp32| per tensor; rel_l2 = ||g_bf16 - g_fp32|| / ||g_fp32||",
              "torch": torch.__version__, "cases": {}}
    for name, batch, frames, seed in (("B2_F32", 2, 32, 5), ("B1_F64", 1, 64, 6)):
        g = torch.Generator().manual_seed(seed)
        z = torch.randn(batch, frames, 196, hidden, generator=g).bfloat16().float()
        l32, g32 = reference_gradients(arch, wq, z, 32, torch.float32)
        l16, g16 = reference_gradients(arch, wq, z, 32, torch.bfloat16)
        per = {}
        for k, r in g32.items():
            a = g16[k]
            per[k] = {"max_norm": float(np.abs(a - r).max() / max(np.abs(r).max(), 1e-30)),
                      "rel_l2": float(np.linalg.norm(a - r) / max(np.linalg.norm(r), 1e-30))}
        keys = [k for k in g32 if not k.endswith("k_proj.bias")]          # true gradient exactly 0: pure noise on both sides
        a = np.concatenate([g16[k].ravel() for k in keys])
        r = np.concatenate([g32[k].ravel() for k in keys])
        result["cases"][name] = {"loss_fp32": l32, "loss_bf16": l16, "per_tensor": per,
                                 "global_rel_l2": float(np.linalg.norm(a - r) / np.linalg.norm(r)),
                                 "global_cosine": float(a @ r / np.linalg.norm(a) / np.linalg.norm(r))}
        worst = sorted(((v["max_norm"], k) for k, v in per.items() if not k.endswith("k_proj.bias")), reverse=True)[:5]
        print(name, "loss", l32, l16, "global rel L2", result["cases"][name]["global_rel_l2"], "worst", worst, flush=True)
    with open(OUT, "w") as fh:
        json.dump(result, fh, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
