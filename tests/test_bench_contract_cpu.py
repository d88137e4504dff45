"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm prints exactly one
JSON line on stdout (whatever libraries print goes to stderr) with the keys the driver reads, non-zero ranks of a
multi-rank reference run print nothing, and the CUDA arm refuses to run on a CPU-only box instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=e, timeout=900)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "3"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train spectra/s" and d["unit"] == "spectra/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
