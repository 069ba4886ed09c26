#!/usr/bin/env python
"""Measure the legacy memories / scene segmentation on one B200: device time per op with CUDA
events, algorithmic bytes -> achieved GB/s against the measured HBM peak.  The CPU arm (numpy oracle on a bounded
sample of the same workload) is added by `python bench.py --workload legacy`, the only place besides tests/ and
smoke() that may execute oracle/.

    python tools/legacy_bench.py [--frames 512] [--out gpurun_out/legacy_bench.json]

Algorithmic bytes (bf16, row = 729 x 1152 tokens = 1.68 MB), distinct rows per streamed frame: drop 2 (last kept + new),
merge 6 at keep >= 3 (the two sources of the pending average, its two neighbours, the new frame, the written average),
k_drop keep + 1, k_merge keep + 3; k-means per iteration (keep + 1) rows read per frame for
the distances (centroids from L2 counted once) + 1 row per frame for the update; segmentation 1 row per frame.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        for k in ("hbm_gbps", "hbm_gb_s", "hbm_GBps"):
            if k in j:
                return float(j[k]), "MEASURED_PEAKS.json"
        for k, v in j.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v), "MEASURED_PEAKS.json"
    except Exception:
        pass
    return 6550.0, "fallback"


def dev_time(fn, warm=1, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def run(frames: int = 512, keep: int = 3, cpu_frames: int = 24, oracle=None, out=None, quiet: bool = False) -> dict:
    """`oracle`: the oracle module (passed in by bench.py) or None to skip the CPU arm."""
    import mavlm_b200  # noqa: F401
    from mavlm_b200 import legacy as L
    from mavlm_b200 import _lib
    from gen_golden_legacy import ntm_params
    O = oracle

    torch.cuda.set_device(0)
    T, T0, P, D = frames, keep, 729, 1152
    row_b = P * D * 2
    peak, peak_src = peak_hbm()
    g = torch.Generator(device="cuda").manual_seed(0)
    base = torch.randn(8, P, D, device="cuda", generator=g)
    scene = torch.randint(0, 8, (T,), device="cuda", generator=g).sort()[0]
    x = (base[scene] + 0.5 * torch.randn(T, P, D, device="cuda", generator=g)).to(torch.bfloat16)
    del base
    res = {"frames": T, "keep": T0, "row_bytes": row_b, "hbm_peak_gbps": peak, "peak_source": peak_src, "ops": {}}
    lib = _lib.load()

    def record(name, ms, bytes_, launches, extra=None):
        r = {"ms": ms, "frames_per_s": T / ms * 1e3, "algorithmic_gb": bytes_ / 1e9, "achieved_gbps": bytes_ / ms / 1e6,
             "frac_of_hbm_peak": bytes_ / ms / 1e6 / peak, "launches": launches}
        if extra:
            r.update(extra)
        res["ops"][name] = r
        if not quiet:
            print(name, json.dumps(r), flush=True)

    n = T - T0
    coins = [random.randint(0, 1) for _ in range(n)]
    # distinct rows a streamed frame touches (repeats across a launch's row pairs are L2 hits: ncu shows 3.43 MB DRAM
    # per drop launch = 2 rows, 8.47 MB per merge launch = 5 rows read, the written row stays in L2)
    per_frame_rows = {L.DROP: 2, L.MERGE: 2 + min(2, T0 - 1) + 1 + 1, L.K_DROP: T0 + 1, L.K_MERGE: 2 + (T0 - 1) + 1 + 1}
    for name, mode in (("drop", L.DROP), ("merge", L.MERGE), ("k_drop", L.K_DROP), ("k_merge", L.K_MERGE)):
        c0 = lib.mavlm_launch_count()
        L.stream_compress(x, T0, mode, coins, return_steps=False)
        launches = lib.mavlm_launch_count() - c0
        ms = dev_time(lambda: L.stream_compress(x, T0, mode, coins, return_steps=False))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.stream_compress(x, T0, mode, coins, return_steps=False)
        host_ms = (time.perf_counter() - t0) * 1e3          # time the host needs to enqueue the launches
        torch.cuda.synchronize()
        record(f"stream_{name}", ms, n * per_frame_rows[mode] * row_b, launches,
               {"us_per_frame": ms * 1e3 / n, "host_enqueue_ms": host_ms})

    # the same chains, 8 videos at a time (grid.z = video): the per-frame latency is shared by the batch
    Bv, Tv = 8, min(T, 128)
    vids = [x[i * (T // Bv):i * (T // Bv) + Tv] if T // Bv >= Tv else x[:Tv] for i in range(Bv)]
    cb = [coins[:Tv - T0] for _ in range(Bv)]
    for name, mode in (("drop", L.DROP), ("merge", L.MERGE), ("k_drop", L.K_DROP), ("k_merge", L.K_MERGE)):
        ms = dev_time(lambda: L.stream_compress_batched(vids, T0, mode, cb, return_steps=False))
        nb = Bv * (Tv - T0)
        by = nb * per_frame_rows[mode] * row_b
        r = {"videos": Bv, "frames_per_video": Tv, "ms": ms, "frames_per_s": Bv * Tv / ms * 1e3, "algorithmic_gb": by / 1e9,
             "achieved_gbps": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak, "us_per_frame_index": ms * 1e3 / (Tv - T0)}
        res["ops"][f"stream_{name}_batch8"] = r
        if not quiet:
            print(f"stream_{name}_batch8", json.dumps(r), flush=True)

    w = torch.ones(T, device="cuda")
    init = torch.arange(T0)
    X = x.reshape(T, -1)
    c0 = lib.mavlm_launch_count()
    _, _, _, iters = L._kmeans(X, T0, w, init)
    launches = lib.mavlm_launch_count() - c0
    ms = dev_time(lambda: L._kmeans(X, T0, w, init))
    record("weighted_kmeans", ms, (iters + 1) * T * 2 * row_b, launches, {"iterations": iters + 1})

    ms = dev_time(lambda: L.adjacent_cosine(x))
    record("adjacent_cosine", ms, T * row_b, 1)
    ms = dev_time(lambda: L.frame_means(x))
    record("frame_means", ms, T * row_b, 1)
    pooled = torch.randn(T, 196, 3584, device="cuda", generator=g).to(torch.bfloat16)
    ms = dev_time(lambda: L.sample_scenes_priority(pooled, 32))
    record("sample_scenes_priority_196x3584", ms, T * 196 * 3584 * 2, 4)
    del pooled

    ntm = L.NeuralTuringMachine().eval().cuda().to(torch.bfloat16)
    ntm.load_state_dict({k: torch.from_numpy(v) for k, v in ntm_params(700, D).items()})
    ntm = ntm.to(torch.bfloat16)
    Tn = min(T, 64)
    fn = lambda m, nf, update_ratio: ntm.gated_update(m, nf, update_ratio)   # noqa: E731
    with torch.no_grad():
        c0 = lib.mavlm_launch_count()
        L.attention_feature(x[:Tn], T0, fn, 0.2)
        launches = lib.mavlm_launch_count() - c0
        ms = dev_time(lambda: L.attention_feature(x[:Tn], T0, fn, 0.2))
    steps = (Tn - T0 + T0 - 1) // T0
    m_rows = T0 * P
    flops = steps * (2 * 2 * m_rows * D * D + 2 * 2 * m_rows * m_rows * D)
    res["ops"]["turing_memory"] = {"frames": Tn, "ms": ms, "frames_per_s": Tn / ms * 1e3, "tflops": flops / ms / 1e9,
                                   "launches": launches}
    if not quiet:
        print("turing_memory", json.dumps(res["ops"]["turing_memory"]), flush=True)
    if O is None:
        return _save(res, out)

    # numpy oracle on a bounded sample of the same frames (host cores: numpy / BLAS threads as configured)
    Tc = cpu_frames
    xc = x[:Tc].float().cpu().numpy()
    cpu = {}
    for name, fn_ in (("drop", lambda: O.drop_feature(xc, T0, coins[:Tc - T0])), ("merge", lambda: O.merge_feature(xc, T0)),
                      ("k_drop", lambda: O.k_drop_feature(xc, T0, coins[:Tc - T0])),
                      ("k_merge", lambda: O.k_merge_feature(xc, T0))):
        t0 = time.perf_counter()
        fn_()
        cpu[f"stream_{name}"] = {"frames": Tc, "frames_per_s": Tc / (time.perf_counter() - t0)}
    t0 = time.perf_counter()
    O.kmeans_feature(xc, T0, list(range(T0)), random.randint, weights=np.ones(Tc, np.float32), weighted=True)
    cpu["weighted_kmeans"] = {"frames": Tc, "frames_per_s": Tc / (time.perf_counter() - t0)}
    res["cpu_oracle"] = {"kind": "port", "cores": os.cpu_count(), "sample": f"first {Tc} frames", "ops": cpu}
    if not quiet:
        print("cpu_oracle", json.dumps(res["cpu_oracle"]), flush=True)
    return _save(res, out)


def _save(res: dict, out) -> dict:
    if out:
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "w") as f:
            json.dump(res, f, indent=1)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--keep", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "legacy_bench.json"))
    args = ap.parse_args()
    run(args.frames, args.keep, out=args.out)


if __name__ == "__main__":
    main()
