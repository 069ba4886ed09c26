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
    need = (REQUIRED - {"impl"}) | {"clocks", "gpu_launches", "roofline"}
    assert need <= set(j), need - set(j)
    assert j["n_gpus"] == 1 and j["steps"] == 5 and j["warmup"] == 3 and j["dtype"] == "bf16" and j["scaling"] == "weak"
    assert j["value"] > 1000 and abs(j["value"] - 64 * 1e3 / j["ms_per_step"]) < 1e-6 * j["value"]
    r = j["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    e = j["e2e"]
    assert e["value"] > 1000 and e["h2d_bytes_per_step"] == 64 * 729 * 1152 * 2 and e["d2h_bytes_per_step"] > 0
    c = j["cpu_baseline"]
    assert c["kind"] == "port" and c["value"] > 0 and c["cores"] >= 1 and c["sample"]
    assert j["gpu_launches"] > 0 and set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
