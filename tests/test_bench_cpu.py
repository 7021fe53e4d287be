"""CPU: bench.py's reference arm on BASELINE.json configs[0] (the reference's own CPU-runnable case) prints the
contract's JSON line, a rank > 0 of a multi-rank launch exits without work, and our own arm refuses to run
without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args, env=None):
    e = dict(os.environ)
    e.pop("RANK", None), e.pop("WORLD_SIZE", None), e.pop("LOCAL_RANK", None)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_config1_json_line():
    r = _bench("--impl", "reference", "--workload", "cube64_4tot_3lat", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout, got %d" % len(lines)
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gvoxel/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "cube64_4tot_3lat" and d["config"]["same_config"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert abs(d["value"] - 64 ** 3 / (d["ms_per_step"] * 1e-3) / 1e9) <= 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and abs(cb["value"] - d["value"]) <= 1e-9 and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    r = _bench("--impl", "reference", "--gpus", "2", "--workload", "cube64_4tot_3lat", "--steps", "1", "--warmup", "0",
               env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_our_arm_has_no_cpu_fallback():
    r = _bench("--workload", "cube64_4tot_3lat", "--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
