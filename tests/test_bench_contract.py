"""bench.py contract pieces that can be checked without a GPU: the reference arm (CPU oracle port) prints the
required JSON line on rank 0 and stays silent on the other ranks; the GPU arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=env)


def test_reference_arm_prints_the_contract_line_on_rank0_only():
    args = ("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-rays", "2", "--levels", "2")
    r0 = _run({"RANK": "0", "WORLD_SIZE": "2"}, *args)
    assert r0.returncode == 0, r0.stderr[-400:]
    line = json.loads([ln for ln in r0.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "rays/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 2 and line["steps"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["metric"].startswith("training rays/sec") and "workload" in line["config"]
    r1 = _run({"RANK": "1", "WORLD_SIZE": "2"}, *args)
    assert r1.returncode == 0 and not [ln for ln in r1.stdout.splitlines() if ln.startswith("{")]


def test_gpu_arm_fails_loudly_without_cuda():
    r = _run({"RANK": "0", "WORLD_SIZE": "1"}, "--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "needs a GPU" in (r.stderr + r.stdout)
