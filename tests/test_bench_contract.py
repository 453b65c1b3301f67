"""bench.py's reference arm runs on host cores only and keeps the JSON contract of the driver
(one line; impl, metric, unit, value, cpu_baseline, e2e with zero copy bytes).  CPU-only test."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "evals/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["steps"] == 1 and j["warmup"] == 1 and j["n_gpus"] == 1
    assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0 and j["dtype"] == "f64" and "config5" in j["config"]["workload"]
    assert "ln-posterior evals/sec" in j["metric"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert res.returncode == 0 and res.stdout.strip() == ""
