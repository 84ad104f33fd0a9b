"""bench.py contract (CPU part): `--impl reference` prints one JSON line with the keys the driver reads, and exits 0
on ranks other than 0 without output."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra=None, *args):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--seconds", "20",
                           "--steps", "1", "--warmup", "0", *args], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_json_line():
    res = _run()
    assert res.returncode == 0, res.stderr
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["metric"] == "PCM inter-channel samples/sec encoded"
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "frames" in cb["sample"]
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_are_silent():
    res = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert res.returncode == 0 and res.stdout.strip() == ""
