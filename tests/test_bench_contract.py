"""CPU: the two things about bench.py that can be checked without a GPU -- the reference arm (`--impl reference`, the
oracle port on the host cores) prints exactly ONE JSON line carrying the contract's keys, and the native arm refuses
to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

import helpers

BENCH = os.path.join(helpers.ROOT, "bench.py")


def _run(*flags):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    return subprocess.run([sys.executable, BENCH, *flags], capture_output=True, text=True, env=env, cwd=helpers.ROOT,
                          timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--size", "32")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("render+warp+photometric fwd+bwd frames/sec")
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert abs(d["value"] - 2 * d["config"]["pairs_per_step"] / (d["ms_per_step"] / 1e3)) <= 1e-6 * d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, cwd=helpers.ROOT, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a box without a GPU")
def test_native_arm_refuses_to_run_without_a_gpu():
    r = _run("--steps", "1", "--size", "32")
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr
