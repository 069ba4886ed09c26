#!/usr/bin/env python
"""Generate tests/golden/*.npz|json by EXECUTING THE UNMODIFIED REFERENCE in this container.

The reference (read-only under /root/reference) is imported by file path -- nothing is copied.
The fixtures pin ``oracle/vismem_oracle.py`` (tests/test_oracle_golden.py) and, through it, the
CUDA path.  /root/reference does not exist on the GPU box, so only the committed fixtures travel.

    python tools/gen_golden.py            # rewrites tests/golden/

Everything is seeded; torch CPU fp32 (fp64 where noted).
"""
from __future__ import annotations

import importlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("MAVLM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _install_llava_stub():
    """`import llava` fails under transformers 5.x, so register a stub package
    whose __path__ points at the reference and stub the one broken import (Q-Former)."""
    if "llava" in sys.modules:
        return
    pkg = types.ModuleType("llava")
    pkg.__path__ = [os.path.join(REF, "llava")]
    sys.modules["llava"] = pkg
    q = types.ModuleType("llava.model.multimodal_resampler.qformer")

    class Qformer:  # never constructed on this path
        pass

    q.Qformer = Qformer
    mp = types.ModuleType("llava.model")
    mp.__path__ = [os.path.join(REF, "llava", "model")]
    sys.modules["llava.model"] = mp
    sys.modules["llava.model.multimodal_resampler.qformer"] = q


def sd_np(module, prefix=""):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def build_rmt(MC, d, depth=2, seed=0):
    cfg = MC.Config()
    cfg.mm_hidden_size = d
    cfg.mm_hidden_act = "relu"
    cfg.mm_num_attention_heads = 8
    cfg.patch_size = 196
    cfg.mm_layer_norm_eps = 1e-12
    cfg.mm_intermediate_size = 4 * d
    cfg.num_memory_tokens = 8
    cfg.depth = depth
    cfg.mm_dtype = torch.float32
    torch.manual_seed(seed)
    m = MC.TransformerProjector(cfg)
    # non-trivial LN affine so that gamma/beta handling is actually pinned
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("layernorm.weight"):
                p.add_(0.1 * torch.randn_like(p))
            if n.endswith("layernorm.bias"):
                p.add_(0.1 * torch.randn_like(p))
    return m.eval()


def gen_rmt(MC):
    d = 32
    m = build_rmt(MC, d)
    w = sd_np(m, "recurrent_memory_transformer.")
    torch.manual_seed(1234)
    frames = torch.randn(6, 196, d)
    out = {"frames": frames.numpy()}
    # (1) 3 chunks of 2 frames: formation + evolution
    m.memory_cache = []
    m.frame_attn_scores = []
    with torch.no_grad():
        for i in range(3):
            cache, scores = m(frames[2 * i:2 * i + 2])
    for i, s in enumerate(cache):
        out[f"state{i}"] = s.numpy().copy()
    for i, s in enumerate(scores):
        out[f"score{i}"] = s.numpy().copy()
    # (2) stress: q_proj x8, inputs x4 -- sharp softmax
    ms = build_rmt(MC, d)
    with torch.no_grad():
        for n, p in ms.named_parameters():
            if "q_proj" in n:
                p.mul_(8.0)
    ms.memory_cache = []
    with torch.no_grad():
        for i in range(3):
            cache_s, _ = ms(4.0 * frames[2 * i:2 * i + 2])
    out["stress_state_last"] = cache_s[-1].numpy().copy()
    out["stress_state_first"] = cache_s[0].numpy().copy()
    # the same two runs in float64 (reference module .double()): the sharp-softmax case amplifies
    # fp32 rounding to ~2e-4, so the oracle is pinned tightly in fp64 and loosely in fp32
    for tag, mod, scale in (("f64_state_last", m, 1.0), ("f64_stress_state_last", ms, 4.0)):
        md = build_rmt(MC, d).double()
        md.load_state_dict({k: v.double() for k, v in mod.state_dict().items()})
        md.memory_cache = []
        with torch.no_grad():
            for i in range(3):
                cache_d, _ = md(scale * frames[2 * i:2 * i + 2].double())
        out[tag] = cache_d[-1].numpy().copy()
    # (3) cache cap: 12 chunks of 1 frame -> 10 retained; ragged tail: 5 frames in chunks of 2
    torch.manual_seed(4321)
    frames12 = torch.randn(12, 196, d)
    out["frames12"] = frames12.numpy()
    m.memory_cache = []
    with torch.no_grad():
        for i in range(12):
            cache12, _ = m(frames12[i:i + 1])
    out["cap_len"] = np.array(len(cache12))
    out["cap_state_first"] = cache12[0].numpy().copy()
    out["cap_state_last"] = cache12[-1].numpy().copy()
    out["cap_state_sums"] = np.array([float(s.double().sum()) for s in cache12])
    np.savez(os.path.join(OUT, "rmt_small.npz"), **out, **{"w::" + k: v for k, v in w.items()})


def gen_rmt_grads(MC):
    """Gradients of loss = mean(fuser-free final states^2) through 2 chunks (BPTT) -- pins the
    backward restatement.  fp64 for a clean reference."""
    d = 16
    m = build_rmt(MC, d, seed=7).double()
    w = sd_np(m, "recurrent_memory_transformer.")
    torch.manual_seed(99)
    frames = torch.randn(4, 196, d, dtype=torch.float64)
    m.memory_cache = []
    for i in range(2):
        cache, _ = m(frames[2 * i:2 * i + 2])
    loss = sum((s * s).mean() for s in cache)
    loss.backward()
    out = {"frames": frames.numpy(), "loss": np.array(float(loss))}
    for n, p in m.named_parameters():
        out["g::recurrent_memory_transformer." + n] = p.grad.numpy().copy()
    np.savez(os.path.join(OUT, "rmt_grads.npz"), **out, **{"w::" + k: v for k, v in w.items()})


def gen_path_grads(MC):
    """Gradients of loss = mean(sequence^2) for memory + fuser + type embeddings + newline + prompt embeddings
    (reference TransformerProjector; the glue below follows llava_arch.py:545-557, 620-629, 708-731), fp64.
    6 pooled frames in chunks of 2 -> 3 chunks (formation x3, evolution x2), 6 fine frames."""
    d = 16
    rmt = build_rmt(MC, d, seed=11).double()
    torch.manual_seed(12)
    fuser = torch.nn.Sequential(torch.nn.Linear(d, 4 * d), torch.nn.GELU(), torch.nn.Linear(4 * d, d)).double()
    tte = torch.nn.Embedding(2, d).double()
    newline = torch.nn.Parameter(torch.randn(d, dtype=torch.float64) * d ** -0.5)
    emb = torch.nn.Embedding(50000, d).double()
    z = torch.randn(6, 196, d, dtype=torch.float64)
    rmt.memory_cache = []
    for i in range(3):
        cache, _ = rmt(z[2 * i:2 * i + 2])
    mem = fuser(torch.cat(cache, dim=0))
    mem = mem + tte(torch.zeros((mem.shape[0], 196), dtype=torch.long))
    fine_idx = torch.clamp(torch.round(torch.linspace(0, 5, steps=6)).long(), 0, 5)
    fine = z[fine_idx] + tte(torch.ones((6, 196), dtype=torch.long))
    pm = emb(torch.tensor([1986, 374, 264, 1550, 11591, 12126, 315, 279, 2766, 25]))
    pf = emb(torch.tensor([9485, 525, 48876, 9124, 14087, 504, 279, 2766, 25]))
    seq = torch.cat([pm, mem.flatten(0, 1), newline[None], pf, fine.flatten(0, 1), newline[None]], dim=0)
    loss = (seq * seq).mean()
    loss.backward()
    out = {"z": z.numpy(), "loss": np.array(float(loss.detach())), "sequence": seq.detach().numpy()}
    for pref, mod in (("recurrent_memory_transformer.", rmt), ("memory_fuser.", fuser), ("token_type_embedding.", tte)):
        for n, p_ in mod.named_parameters():
            out["w::" + pref + n] = p_.detach().numpy().copy()
            out["g::" + pref + n] = p_.grad.numpy().copy()
    out["w::image_newline"] = newline.detach().numpy().copy()
    out["g::image_newline"] = newline.grad.numpy().copy()
    rows = sorted(set([1986, 374, 264, 1550, 11591, 12126, 315, 279, 2766, 25, 9485, 525, 48876, 9124, 14087, 504]))
    out["embed_rows"] = np.array(rows)
    out["w::embed_rows"] = emb.weight.detach().numpy()[rows].copy()
    out["g::embed_rows"] = emb.weight.grad.numpy()[rows].copy()
    np.savez(os.path.join(OUT, "path_grads.npz"), **out)


def gen_pe(PE):
    out = {}
    for d in (32, 896, 3584):
        pe = PE.TemporalPositionalEncoding(max_frames=600, embed_dim=d, learnable=False)
        t = pe.frame_embed.numpy()
        if d == 32:
            out["table32"] = t.copy()
        out[f"rows{d}"] = t[[0, 1, 7, 131, 599]].copy()
        out[f"sum{d}"] = np.array(float(pe.frame_embed.double().sum()))
    pe = PE.TemporalPositionalEncoding(max_frames=600, embed_dim=32, learnable=False)
    torch.manual_seed(5)
    x = torch.randn(5, 7, 32)
    idx = torch.tensor([0, 3, 17, 256, 599])
    out["x"] = x.numpy()
    out["idx"] = idx.numpy()
    out["y"] = pe(x, idx).numpy()
    out["y_default_idx"] = pe(x).numpy()
    out["y_bf16"] = pe(x.bfloat16(), idx).float().numpy()
    np.savez(os.path.join(OUT, "pe.npz"), **out)


def gen_projector_fuser(PB, MF):
    cfg = types.SimpleNamespace(mm_projector_type="mlp2x_gelu", mm_hidden_size=48, hidden_size=32)
    torch.manual_seed(0)
    proj = PB.build_vision_projector(cfg).eval()
    torch.manual_seed(11)
    x = torch.randn(3, 50, 48)
    out = {"proj_x": x.numpy(), "proj_y": proj(x).detach().numpy()}
    out.update({"w::" + k: v for k, v in sd_np(proj, "mm_projector.").items()})
    d = 32
    torch.manual_seed(0)
    fuser = torch.nn.Sequential(torch.nn.Linear(d, 4 * d), torch.nn.GELU(), torch.nn.Linear(4 * d, d)).eval()  # llava_arch.py:132-136
    torch.manual_seed(12)
    m = torch.randn(16, 196, d)
    out["fuser_x"] = m.numpy()
    out["fuser_y"] = fuser(m).detach().numpy()
    out.update({"w::" + k: v for k, v in sd_np(fuser, "memory_fuser.").items()})
    for nl in (1, 2):
        torch.manual_seed(3)
        enc = MF.MemoryFuser(hidden_dim=d, num_layers=nl, num_heads=4, dropout=0.1, device="cpu").eval()
        torch.manual_seed(13)
        xe = torch.randn(4, 196, d)
        with torch.no_grad():
            # fast-path off so that the eval result is the plain post-LN math
            torch.backends.mha.set_fastpath_enabled(False)
            ye = enc(xe)
        out[f"enc{nl}_x"] = xe.numpy()
        out[f"enc{nl}_y"] = ye.numpy()
        out.update({f"w::enc{nl}." + k: v for k, v in sd_np(enc).items()})
    np.savez(os.path.join(OUT, "projector_fuser.npz"), **out)


class _Tower(torch.nn.Module):
    """Fake SigLIP: [F,Dv,27,27] 'pixels' -> [F,729,Dv] tokens (the tower is out of scope)."""
    num_patches_per_side = 27

    def forward(self, images):
        return images.flatten(2).transpose(1, 2).contiguous()


def build_harness(ARCH, PB, d=16, dv=4, vocab=64, pool_mode="bilinear"):
    class Inner(torch.nn.Module):
        def __init__(self):
            super().__init__()
            cfg = types.SimpleNamespace(mm_projector_type="mlp2x_gelu", mm_hidden_size=dv, hidden_size=d)
            self.vision_tower = _Tower()
            self.mm_projector = PB.build_vision_projector(cfg)
            c = ARCH.Config()
            c.mm_hidden_size = d
            c.mm_hidden_act = "relu"
            c.mm_num_attention_heads = 8
            c.patch_size = 196
            c.mm_layer_norm_eps = 1e-12
            c.mm_intermediate_size = 4 * d
            c.num_memory_tokens = 8
            c.depth = 2
            c.mm_dtype = torch.float32
            self.recurrent_memory_transformer = ARCH.TransformerProjector(c)
            self.memory_fuser = torch.nn.Sequential(torch.nn.Linear(d, 4 * d), torch.nn.GELU(), torch.nn.Linear(4 * d, d))
            self.positional_encoding = ARCH.TemporalPositionalEncoding(max_frames=600, embed_dim=d, learnable=False)
            self.token_type_embedding = torch.nn.Embedding(2, d)
            self.image_newline = torch.nn.Parameter(torch.randn(d) * d ** -0.5)
            # vocab must cover the hard-coded prompt ids (max 48876): embed via modulo table
            self._emb = torch.nn.Embedding(vocab, d)

        def embed_tokens(self, ids):
            return self._emb(ids % self._emb.num_embeddings)

        def get_vision_tower(self):
            return self.vision_tower

    class Harness(torch.nn.Module, ARCH.LlavaMetaForCausalLM):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.model = Inner()
            self.config = types.SimpleNamespace(mm_spatial_pool_mode=pool_mode, mm_patch_merge_type="spatial_unpad",
                                                image_aspect_ratio="anyres_max_9", mm_newline_position="one_token",
                                                dropout_frames=False, tokenizer_model_max_length=32768,
                                                tokenizer_padding_side="right")
            self.device = torch.device("cpu")

        def get_model(self):
            return self.model

    torch.manual_seed(0)
    return Harness().eval()


def gen_pool_and_full(ARCH, PB):
    h = build_harness(ARCH, PB)
    torch.manual_seed(21)
    x = torch.randn(2, 729, 8)
    out = {"x": x.numpy(), "bilinear": h.get_2dPool(x).numpy()}
    xd = x.double()
    out["bilinear_f64"] = h.get_2dPool(xd).numpy()
    for mode in ("average", "max"):
        h.config.mm_spatial_pool_mode = mode
        out[mode] = h.get_2dPool(x).numpy()
    h.config.mm_spatial_pool_mode = "bilinear"
    # other geometries of the same op (stride 3; 24x24 tower)
    out["bilinear_s3"] = h.get_2dPool(x, stride=3).numpy()
    np.savez(os.path.join(OUT, "pool.npz"), **out)

    # full prepare_inputs_labels_for_multimodal: 70 raw frames -> 64 sampled -> 2 chunks
    torch.manual_seed(33)
    video = torch.randn(70, 4, 27, 27)
    ids = torch.tensor([[5, 7, ARCH.IMAGE_TOKEN_INDEX, 9, 11]])
    with torch.no_grad():
        res = h.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, [video], modalities=["video"])
    emb = res[4]
    w = sd_np(h.model)
    full = {"video": video.numpy(), "input_ids": ids.numpy(), "inputs_embeds": emb.numpy()}
    full.update({"w::" + k: v for k, v in w.items()})
    np.savez(os.path.join(OUT, "full_path.npz"), **full)

    # ---- text / vision splice variants (llava_arch.py:745-878) on the same video: labels / mask / position ids,
    # truncation, left and right padding, a batch of two with the batch-1 IndexError fallback (:799-802)
    sp = {}
    full_seq = emb[0][2:-2]                                   # the video token sequence itself
    sp["video_sequence"] = full_seq.numpy()
    # (the fork is batch-1 for videos: llava_arch.py:436, :561, :723-731 -- so two single-sample calls)
    ids2 = torch.tensor([[5, ARCH.IMAGE_TOKEN_INDEX, 9, 11, 0], [7, ARCH.IMAGE_TOKEN_INDEX, 8, 0, 0]])
    mask2 = torch.tensor([[1, 1, 1, 1, 0], [1, 1, 1, 0, 0]])
    lab2 = torch.tensor([[1, -100, 2, 3, -100], [5, -100, 6, -100, -100]])
    pos2 = torch.arange(5)[None].expand(2, 5).contiguous()
    for tag, side, maxlen, row in (("right", "right", 32768, 0), ("left_trunc", "left", 9000, 1)):
        h.config.tokenizer_padding_side = side
        h.config.tokenizer_model_max_length = maxlen
        with torch.no_grad():
            r = h.prepare_inputs_labels_for_multimodal(ids2[row:row + 1], pos2[row:row + 1], mask2[row:row + 1], None,
                                                       lab2[row:row + 1], [video], modalities=["video"])
        _, p_out, m_out, _, e_out, l_out = r
        sp[f"{tag}.position_ids"] = p_out.numpy()
        sp[f"{tag}.attention_mask"] = m_out.numpy()
        sp[f"{tag}.labels"] = l_out.numpy()
        sp[f"{tag}.shape"] = np.array(e_out.shape)
        sp[f"{tag}.rows_head"] = e_out[:, :12].numpy()
        sp[f"{tag}.rows_tail"] = e_out[:, -12:].numpy()
        sp[f"{tag}.rows_stride"] = e_out[:, ::97].numpy()
        sp[f"{tag}.colsum"] = e_out.double().sum(dim=1).numpy()
    h.config.tokenizer_padding_side = "right"
    h.config.tokenizer_model_max_length = 32768
    sp["input_ids"], sp["mask"], sp["labels"], sp["pos"] = ids2.numpy(), mask2.numpy(), lab2.numpy(), pos2.numpy()
    sp["embed_table"] = w["_emb.weight"]
    np.savez(os.path.join(OUT, "splice.npz"), **sp)

    # short video (< 32 frames): 5 frames, single ragged chunk, 5 fine frames
    torch.manual_seed(34)
    video5 = torch.randn(5, 4, 27, 27)
    with torch.no_grad():
        res5 = h.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, [video5], modalities=["video"])
    np.savez(os.path.join(OUT, "full_path_short.npz"), video=video5.numpy(), inputs_embeds=res5[4].numpy())


def gen_indices(SEG):
    out = {"sample": {}, "fine": {}, "bounds": {}}
    for f in (1, 5, 31, 32, 33, 63, 64, 65, 70, 96, 100, 127, 128, 129, 255, 256, 300, 599, 600, 1000, 1024, 4097):
        if f < 32:
            n = f
        else:
            n = max((f // 32) * 32, 64)
        out["sample"][str(f)] = torch.linspace(0, f - 1, steps=n).long().tolist()          # llava_arch.py:451
    for n in (1, 5, 31, 32, 33, 64, 96, 128, 256, 576, 1024):
        k = min(32, n)
        idx = torch.clamp(torch.round(torch.linspace(0, n - 1, steps=k)).long(), 0, n - 1)   # llava_arch.py:520-522
        out["fine"][str(n)] = idx.tolist()
    for t, d in ((64, 32), (70, 32), (5, 32), (32, 32), (33, 32), (256, 16), (250, 16), (1024, 8), (7, 8), (0, 32)):
        out["bounds"][f"{t},{d}"] = SEG.uniform_segment_variant(torch.zeros(t, 1), d=d)
    with open(os.path.join(OUT, "indices.json"), "w") as fh:
        json.dump(out, fh)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    _install_llava_stub()
    MC = _load("ref_memory_controller", "llava/model/memory_module/MemoryController.py")
    PE = _load("ref_position_encoding", "llava/model/memory_module/position_encoding.py")
    MF = _load("ref_memory_fuser", "llava/model/memory_module/MemoryFuser.py")
    SEG = _load("ref_segment", "llava/model/memory_module/segment.py")
    PB = importlib.import_module("llava.model.multimodal_projector.builder")
    ARCH = importlib.import_module("llava.model.llava_arch")
    gen_rmt(MC)
    gen_rmt_grads(MC)
    gen_path_grads(MC)
    gen_pe(PE)
    gen_projector_fuser(PB, MF)
    gen_pool_and_full(ARCH, PB)
    gen_indices(SEG)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
