#!/usr/bin/env python
"""Measure the REFERENCE's own bf16-vs-fp32 drift of the memory state (unmodified reference modules,
path-imported, CPU) -> tests/golden/noise_floor.json.

The bf16 tolerance of 2e-2 (BASELINE.json) sits at the reference's own noise floor for default
weights; in the mandatory stress variant (q_proj x8, inputs x4: sharp softmax) the
recurrence amplifies bf16 rounding chaotically and the reference itself drifts by 16 % / 81 %, so the
stress test bounds the CUDA path by this measured floor instead of 2e-2.

    python tools/gen_noise_floor.py        # ~2 min on 8 cores
"""
import importlib.util
import json
import os

import torch

REF = os.environ.get("MAVLM_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "noise_floor.json")


def load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    MC = load("ref_mc", "llava/model/memory_module/MemoryController.py")
    torch.set_num_threads(os.cpu_count())

    def build(d, qs):
        cfg = MC.Config()
        cfg.mm_hidden_size, cfg.mm_intermediate_size, cfg.depth, cfg.mm_dtype = d, 4 * d, 2, torch.float32
        torch.manual_seed(0)
        m = MC.TransformerProjector(cfg)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "q_proj" in n:
                    p.mul_(qs)
        return m

    def nerr(a, b):
        return float((a.double() - b.double()).abs().max() / b.double().abs().max())

    out = {"how": "reference TransformerProjector, D=896, 32 pooled frames N(0,1)*x_scale seed 1234, chunks of 16, "
                  "bf16 module vs fp32 module on the same bf16-rounded inputs; err = max|a-b|/max|b| per state"}
    for tag, qs, xs in (("default", 1.0, 1.0), ("stress", 8.0, 4.0)):
        m32, m16 = build(896, qs), build(896, qs).bfloat16()
        g = torch.Generator().manual_seed(1234)
        xb = (torch.randn(32, 196, 896, generator=g) * xs).bfloat16()
        m32.memory_cache, m16.memory_cache = [], []
        with torch.no_grad():
            for i in range(2):
                c32, _ = m32(xb[16 * i:16 * i + 16].float())
                c16, _ = m16(xb[16 * i:16 * i + 16])
        out[tag] = {"q_scale": qs, "x_scale": xs, "state0": nerr(c16[0].float(), c32[0]),
                    "state1": nerr(c16[1].float(), c32[1])}
        print(tag, out[tag])
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
