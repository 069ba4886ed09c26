"""bench.py's output contract, checked on the CPU through the reference arm (the only arm that runs without a GPU):
exactly one line on stdout, valid JSON, the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                       cwd=ROOT, env=e)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line():
    out = _run("--workload", "legacy", "--impl", "reference", "--steps", "1", "--warmup", "0")
    lines = [l for l in out.split("\n") if l.strip()]
    assert len(lines) == 1, out
    j = json.loads(lines[0])
    assert REQUIRED <= set(j), REQUIRED - set(j)
    assert j["impl"] == "reference" and j["value"] > 0 and j["higher_is_better"] is True
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and "workload" in j["config"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    out = _run("--workload", "legacy", "--impl", "reference", "--steps", "1", "--warmup", "0",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""


def test_library_noise_on_fd1_does_not_reach_stdout():
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.isolate_stdout(); "
            "os.write(1, b'NCCL version banner\\n'); print('python noise'); bench.emit_line('{\"ok\": 1}')" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert p.stdout == '{"ok": 1}\n'
    assert "NCCL version banner" in p.stderr and "python noise" in p.stderr


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_line_has_every_contract_key():
    out = _run("--steps", "5", "--warmup", "3")
    lines = [l for l in out.split("\n") if l.strip()]
    assert len(lines) == 1, out[:2000]
    j = json.loads(lines[0])
    need = (REQUIRED - {"impl"}) | {"clocks", "gpu_launches", "roofline", "roofline_kernels", "sustained", "config3"}
    assert need <= set(j), need - set(j)
    assert j["n_gpus"] == 1 and j["steps"] == 5 and j["warmup"] == 3 and j["dtype"] == "bf16" and j["scaling"] == "weak"
    assert j["value"] > 1000 and abs(j["value"] - 64 * 1e3 / j["ms_per_step"]) < 1e-6 * j["value"]
    r = j["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    assert min(abs(r["peak"] - r["achieved"] / r["frac_of_burst"]),
               abs(r["peak"] - r["achieved"] / r["frac_of_sustained"])) < 1e-6 * r["peak"]   # burst or sustained, stated
    fams = {k["kernel"].split(" ")[0].split("<")[0]: k for k in j["roofline_kernels"]}
    assert {"gemm_tc_kernel", "attn_tc_kernel", "layernorm_kernel", "pool_pe_kernel", "assemble_kernel"} <= set(fams)
    for k in j["roofline_kernels"]:
        assert k["bound"] in ("tensor", "hbm") and abs(k["frac"] - k["achieved"] / k["peak"]) < 1e-9 and k["ms_per_step"] > 0
    assert 0.9 < sum(k["share_of_step"] for k in j["roofline_kernels"]) < 1.05
    s_ = j["sustained"]
    assert s_["seconds"] >= 5.0 and s_["value"] > 1000 and set(s_["clocks"]) >= {"sm_mhz", "reasons"}
    c3 = j["config3"]
    assert c3["scaling"] == "strong" and c3["videos_per_rank"] == 8 and c3["value"] > 1000
    assert j["e2e"]["copy_only"]["ms_per_step"] > 0 and isinstance(j["e2e"]["bound"], str)
    e = j["e2e"]
    assert e["value"] > 1000 and e["h2d_bytes_per_step"] == 64 * 729 * 1152 * 2 and e["d2h_bytes_per_step"] > 0
    c = j["cpu_baseline"]
    have_ref = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "llava", "model", "llava_arch.py"))
    assert c["kind"] == ("reference" if have_ref else "port") and c["value"] > 0 and c["cores"] >= 1 and c["sample"]
    assert c["gpu_vs_cpu_sequence_err"] < 3e-2
    assert j["gpu_launches"] > 0 and set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}


def test_reference_cpu_runs_the_unmodified_reference_modules_and_agrees_with_the_oracle():
    """The CPU arm of bench.py (ReferenceCpu) at small dims: when the reference install (baseline/_ref, written by
    __graft_entry__.build()) or /root/reference is present it drives the UNMODIFIED modules (kind "reference") and its
    visual token sequence equals the numpy oracle's on the same weights; otherwise it says "port"."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from baseline import ref_arm
    from mavlm_b200 import synthetic
    from oracle import vismem_oracle as O
    old = bench.HIDDEN, bench.VISION, bench.CHUNK
    bench.HIDDEN, bench.VISION, bench.CHUNK = 64, 16, 32
    try:
        _, w = synthetic.build_pipeline(64, 16, dtype=torch.float32, device="cpu", vocab=50000)
        w32 = {k: v.astype(np.float32) for k, v in w.items()}
        x = synthetic.synthetic_tower_tokens(1, 40, 16, dtype=torch.float32)[0]
        ref = bench.ReferenceCpu(w32, x)
        have = ref_arm.find_reference_root() is not None
        assert ref.kind == ("reference" if have else "port") and ref.cores >= 1
        dt, seq = ref.run(40)                                  # 40 raw frames -> 64 sampled (llava_arch.py:437-451)
        assert dt > 0
        idx = O.sample_frame_indices(40)
        want = O.visual_memory_path(x.numpy()[idx] if have else x.numpy(), idx if have else np.arange(40), w32,
                                    pe_table=w32["positional_encoding.frame_embed"],
                                    prompt_mem=w32["embed_tokens.weight"][list(O.MEMORY_PROMPT_IDS)],
                                    prompt_frm=w32["embed_tokens.weight"][list(O.FRAME_PROMPT_IDS)], chunk=32)["sequence"]
        assert seq.shape == want.shape and O.normalized_max_error(seq, want) < 2e-5
        assert ("UNMODIFIED reference" in ref.describe(40)) == have
    finally:
        bench.HIDDEN, bench.VISION, bench.CHUNK = old
