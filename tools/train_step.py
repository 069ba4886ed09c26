#!/usr/bin/env python
"""BASELINE config[3]: OV-7B finetune_short-style forward+backward of the memory module and fuser, bf16,
batch of videos x 32 frames (one chunk: memory_update_attention gets no gradient) and a
64-frame case (2 chunks: evolution attention trains too).  Pooled-token input (frame features are detached),
loss = mean-square of the assembled sequence.  Reports ms per training step (fwd+bwd), frames/s and the
algorithmic TFLOP/s (backward counted as 2x forward minus the frame-side dgrad that is never needed).

    python tools/train_step.py [--dims 7b|0.5b] [--batch 8] [--frames 32 64] -> gpurun_out/train_step.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mavlm_b200 import synthetic  # noqa: E402


def fwd_gflop(frames, chunk, d, lq=1568, p=196, depth=2, cap=10):
    n_chunks = -(-frames // chunk)
    fl = 0.0
    for t in range(n_chunks):
        c = min(chunk, frames - t * chunk)
        fl += depth * (4.0 * lq * d * d + 4.0 * c * p * d * d + 4.0 * lq * c * p * d + 16.0 * lq * d * d)
        if t > 0:
            fl += 8.0 * lq * d * d + 4.0 * lq * (min(t, cap) * lq) * d
    fl += 16.0 * min(n_chunks, cap) * lq * d * d
    return fl / 1e9


class UpstreamGrad(torch.autograd.Function):
    """A scalar whose gradient w.r.t. `seq` is the given `dseq` (stands in for the LLM's backward); no extra kernels."""

    @staticmethod
    def forward(ctx, seq, dseq):
        ctx.save_for_backward(dseq)
        return seq.new_zeros(())

    @staticmethod
    def backward(ctx, g):
        (dseq,) = ctx.saved_tensors
        return dseq, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", default="7b")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, nargs="+", default=[32, 64])
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--no-graph", action="store_true", help="eager autograd only (the run that is put under ncu)")
    ap.add_argument("--upstream-grad", action="store_true",
                    help="backward from a fixed d(sequence) as the LLM would hand it over (no loss kernels in the step) "
                         "instead of the parity tests' mean-square loss")
    args = ap.parse_args()
    d = synthetic.OV_DIMS[args.dims]
    pipe, _ = synthetic.build_pipeline(d, 1152, dtype=torch.bfloat16, chunk_size=32, device="cuda:0")
    out = []
    for frames in args.frames:
        g = torch.Generator(device="cuda:0").manual_seed(1234)
        z = torch.randn(args.batch, frames, 196, d, device="cuda:0", generator=g).bfloat16()

        dseq = None

        def loss_fn(seq):
            # --upstream-grad: sum(seq * dseq) has d(loss)/d(seq) = dseq, i.e. the backward starts from a given upstream
            # gradient (one fused multiply-reduce instead of the three fp32 passes of the mean-square loss)
            nonlocal dseq
            if not args.upstream_grad:
                return (seq.float() ** 2).mean()
            if dseq is None:
                gd = torch.Generator(device="cuda:0").manual_seed(77)
                dseq = (torch.randn(seq.shape, device="cuda:0", generator=gd) * 1e-4).to(seq.dtype)
            return UpstreamGrad.apply(seq, dseq)

        def step():
            for p_ in pipe.parameters():
                p_.grad = None
            res = pipe.memory_forward_train(z)
            loss = loss_fn(res["sequence"])
            loss.backward()
            return float(loss.detach())

        step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            loss = step()
        torch.cuda.synchronize()
        ms_eager = (time.perf_counter() - t0) / args.iters * 1e3
        eager_grads = {n_: p_.grad.detach().clone() for n_, p_ in pipe.named_parameters() if p_.grad is not None}
        ms, same = ms_eager, None
        if not args.no_graph:
            # the same step as ONE CUDA graph (forward + loss + backward), device-timed
            gstep = pipe.graphed_train(args.batch, frames, loss_fn=loss_fn)
            gstep(z)
            torch.cuda.synchronize()
            # (not bitwise: bias / LayerNorm gradient sums use atomics) worst normalised difference over the tensors
            same = max(float((p_.grad.float() - eager_grads[n_].float()).abs().max() / eager_grads[n_].float().abs().max().clamp_min(1e-30))
                       for n_, p_ in pipe.named_parameters() if p_.grad is not None and not n_.endswith("k_proj.bias"))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(max(args.iters, 5)):
                gstep(None)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / max(args.iters, 5)
            loss = float(gstep.loss)
            del gstep
        gf = 3.0 * args.batch * fwd_gflop(frames, 32, d)
        n_grads = sum(p_.grad is not None for p_ in pipe.parameters())
        rec = {"dims": args.dims, "batch": args.batch, "frames": frames, "ms_per_step": ms,
               "ms_per_step_eager_autograd": ms_eager, "graph_vs_eager_worst_grad_diff": same,
               "frames_per_s": args.batch * frames / ms * 1e3, "algorithmic_tflops_fwd_bwd": gf / ms,
               "params_with_grad": n_grads, "loss": loss, "upstream_grad": bool(args.upstream_grad),
               "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        out.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "train_step.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
