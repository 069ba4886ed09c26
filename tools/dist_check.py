#!/usr/bin/env python
"""Multi-GPU check + BASELINE config 3 measurement, run under torchrun on N GPUs of one node (NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py

1. parity: videos sharded over ranks (no data-path collective, results all-gathered) and one long video with the
   frame-sharded pre-pass (pooled tokens all-gathered over NVLink) against the single-GPU result.  The frame-sharded
   path must match bit for bit; the video-sharded one within bf16 noise (the attention kernel's key-range split
   depends on the local batch size, so the online-softmax merge order differs).
2. config 3: OV-7B, 8 videos x 256 frames, 16-frame chunks, videos sharded over the ranks; device time (max over
   ranks) of the local recurrences + the NCCL all-gather of the assembled sequences.
Writes gpurun_out/dist_check.json on rank 0.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mavlm_b200 import dist as mdist, synthetic  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}

    # ---- 1. parity at 0.5B dims (same seeds on every rank)
    pipe, _ = synthetic.build_pipeline(896, 1152, dtype=torch.bfloat16, chunk_size=4, device=dev)
    x = synthetic.synthetic_tower_tokens(5, 12, seed=7).to(dev)            # 5 videos x 12 frames (ragged over 2 ranks)
    idx = torch.arange(12)[None].expand(5, 12).contiguous()
    ref = pipe(x, idx, return_states=False)["sequence"]
    got = mdist.encode_videos_sharded(pipe, x, idx)
    out["videos_sharded_err"] = float((got.float() - ref.float()).abs().max() / ref.float().abs().max())
    out["videos_sharded_equal"] = out["videos_sharded_err"] < 1e-2
    long_ref = pipe(x[:1], idx[:1])["sequence"]
    long_got = mdist.encode_long_video_frame_sharded(pipe, x[0], idx[0])["sequence"]
    out["frame_sharded_equal"] = bool(torch.equal(long_got, long_ref))
    del pipe

    # ---- 2. config 3 timing (pooled-token input, like tools/sweep.py)
    videos, frames, chunk = 8, 256, 16
    pipe, _ = synthetic.build_pipeline(3584, 1152, dtype=torch.bfloat16, chunk_size=chunk, device=dev)
    mine = mdist.shard_range(videos, rank, world)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    z = torch.randn(len(mine), frames, 196, 3584, device=dev, generator=g).bfloat16()
    counts = [len(mdist.shard_range(videos, r, world)) for r in range(world)]

    def step():
        seq = pipe.memory_forward(z, return_states=False)["sequence"]
        return mdist.all_gather_rows(seq, counts)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        res = step()
    e1.record()
    torch.cuda.synchronize()
    ms = mdist.max_over_ranks(e0.elapsed_time(e1) / n, device=dev)
    out["config3"] = {"videos": videos, "frames": frames, "chunk": chunk, "ms": ms,
                      "frames_per_s": videos * frames / ms * 1e3, "gathered_shape": list(res.shape)}
    if rank == 0:
        print(json.dumps(out), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"dist_check_n{world}.json"), "w") as fh:
            json.dump(out, fh, indent=1)
    dist.destroy_process_group()
    assert out["videos_sharded_equal"] and out["frame_sharded_equal"], out


if __name__ == "__main__":
    main()
