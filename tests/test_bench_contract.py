"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm prints
exactly one JSON line on stdout with the keys the driver reads; the GPU arm refuses to run
without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import helpers


def _run(args):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    return subprocess.run([sys.executable, os.path.join(helpers.ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=env)


def test_reference_arm_prints_one_json_line(built):
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--config", "1", "--cpu-sample", "48"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GCUPS" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["scaling"] == "weak" and d["vs_baseline"] is None


def test_gpu_arm_has_no_cpu_fallback(built):
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the arm runs for real (tests -m gpu / the driver)
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
