"""Reference arm: the UNMODIFIED reference's own CPU implementation of the visual-memory path, driven through its
own entry point (`LlavaMetaForCausalLM.prepare_inputs_labels_for_multimodal`, llava_arch.py:388-878).

Where the reference comes from: `baseline/_ref/` (git-ignored; written by
`python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`,
which `__graft_entry__.build()` runs when /root/reference is present) -- it travels to the GPU box with the snapshot --
or /root/reference itself in the build container.  Nothing of the reference is tracked in this repository.

`import llava` does not work under transformers 5.x (llava/__init__.py pulls in every language model; the Q-Former
resampler imports helpers that transformers removed), so the package is entered through a stub parent package and the
one broken, unused import is stubbed -- the same path-import `tools/gen_golden.py` uses to produce the golden fixtures
(SURVEY.md §8c).  The modules on the path (MemoryController.py, position_encoding.py, multimodal_projector/builder.py,
llava_arch.py) run unmodified.  The SigLIP tower is outside the path (BASELINE.json north_star: untouched): a stand-in
"tower" hands the synthetic tower tokens through, exactly as the golden harness does.

Test / benchmark infrastructure only; the product package never imports this file.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from typing import Dict, Optional

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference_root() -> Optional[str]:
    """Directory that contains the reference's `llava/` package, or None."""
    for cand in (os.environ.get("MAVLM_REFERENCE"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "llava", "model", "llava_arch.py")):
            return cand
    return None


def load_reference(root: str):
    """Returns (llava.model.llava_arch, llava.model.multimodal_projector.builder) of the reference at `root`."""
    if "llava" not in sys.modules:
        pkg = types.ModuleType("llava")
        pkg.__path__ = [os.path.join(root, "llava")]
        sys.modules["llava"] = pkg
        mp = types.ModuleType("llava.model")
        mp.__path__ = [os.path.join(root, "llava", "model")]
        sys.modules["llava.model"] = mp
        q = types.ModuleType("llava.model.multimodal_resampler.qformer")

        class Qformer:  # never constructed on this path (resampler call is commented out, llava_arch.py:301)
            pass

        q.Qformer = Qformer
        sys.modules["llava.model.multimodal_resampler.qformer"] = q
    arch = importlib.import_module("llava.model.llava_arch")
    pb = importlib.import_module("llava.model.multimodal_projector.builder")
    return arch, pb


class _PassThroughTower(torch.nn.Module):
    """[F, Dv, 27, 27] -> [F, 729, Dv]: the tower is not on the path; its OUTPUT is the synthetic input."""
    num_patches_per_side = 27

    def forward(self, images):
        return images.flatten(2).transpose(1, 2).contiguous()


def build_reference_model(arch, pb, weights: Dict[str, np.ndarray], hidden: int, vision_dim: int,
                          dtype: torch.dtype = torch.float32):
    """The reference's modules at (hidden, vision_dim) with `weights` (reference state_dict keys, as
    mavlm_b200.synthetic.export_weights returns them) loaded; hyper-parameters as llava_arch.py:117-150."""
    vocab = weights["embed_tokens.weight"].shape[0]
    max_frames = weights["positional_encoding.frame_embed"].shape[0]

    class Inner(torch.nn.Module):
        def __init__(self):
            super().__init__()
            cfg = types.SimpleNamespace(mm_projector_type="mlp2x_gelu", mm_hidden_size=vision_dim, hidden_size=hidden)
            self.vision_tower = _PassThroughTower()
            self.mm_projector = pb.build_vision_projector(cfg)
            c = arch.Config()
            c.mm_hidden_size = hidden
            c.mm_hidden_act = "relu"
            c.mm_num_attention_heads = 8
            c.patch_size = 196
            c.mm_layer_norm_eps = 1e-12
            c.mm_intermediate_size = 4 * hidden
            c.num_memory_tokens = 8
            c.depth = 2
            c.mm_dtype = torch.float32
            self.recurrent_memory_transformer = arch.TransformerProjector(c)
            self.memory_fuser = torch.nn.Sequential(torch.nn.Linear(hidden, 4 * hidden), torch.nn.GELU(),
                                                    torch.nn.Linear(4 * hidden, hidden))
            self.positional_encoding = arch.TemporalPositionalEncoding(max_frames=max_frames, embed_dim=hidden,
                                                                       learnable=False)
            self.token_type_embedding = torch.nn.Embedding(2, hidden)
            self.image_newline = torch.nn.Parameter(torch.zeros(hidden))
            self.embed_tokens = torch.nn.Embedding(vocab, hidden)

        def get_vision_tower(self):
            return self.vision_tower

    class Harness(torch.nn.Module, arch.LlavaMetaForCausalLM):
        def __init__(self):
            torch.nn.Module.__init__(self)
            self.model = Inner()
            self.config = types.SimpleNamespace(mm_spatial_pool_mode="bilinear", mm_patch_merge_type="spatial_unpad",
                                                image_aspect_ratio="anyres_max_9", mm_newline_position="one_token",
                                                dropout_frames=False, tokenizer_model_max_length=32768,
                                                tokenizer_padding_side="right")
            self.device = torch.device("cpu")

        def get_model(self):
            return self.model

    h = Harness().eval()
    inner = h.model
    with torch.no_grad():
        for pref, mod in (("recurrent_memory_transformer.", inner.recurrent_memory_transformer),
                          ("memory_fuser.", inner.memory_fuser), ("mm_projector.", inner.mm_projector),
                          ("token_type_embedding.", inner.token_type_embedding)):
            sd = {k[len(pref):]: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in weights.items()
                  if k.startswith(pref)}
            mod.load_state_dict(sd, strict=True)
        inner.image_newline.copy_(torch.from_numpy(weights["image_newline"]).float())
        inner.embed_tokens.weight.copy_(torch.from_numpy(weights["embed_tokens.weight"]).float())
    if dtype != torch.float32:
        h = h.to(dtype)
    return h


def tokens_as_video(tower_tokens: torch.Tensor) -> torch.Tensor:
    """[F, 729, Dv] tower tokens -> the [F, Dv, 27, 27] 'video' the pass-through tower turns back into them."""
    f, n, dv = tower_tokens.shape
    return tower_tokens.transpose(1, 2).reshape(f, dv, 27, 27).contiguous()


@torch.no_grad()
def reference_pass(h, arch, video: torch.Tensor) -> torch.Tensor:
    """One call of the reference's prepare_inputs_labels_for_multimodal on one video; returns the visual token sequence
    [L, D] (the embeddings between the two text tokens on either side of the image placeholder)."""
    ids = torch.tensor([[5, 7, arch.IMAGE_TOKEN_INDEX, 9, 11]])
    res = h.prepare_inputs_labels_for_multimodal(ids, None, None, None, None, [video], modalities=["video"])
    return res[4][0][2:-2]
