"""bench.py's reference arm (the reference's own CPU implementation on the host cores) prints the JSON line the driver
parses, and stays silent on ranks other than 0. CPU only; a small workload and a sub-second budget keep it short."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, *flags):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ml100k", "--steps", "2",
                           "--warmup", "1", "--cpu-budget", "0.4", *flags], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)


def test_reference_arm_prints_the_contract_line():
    p = _run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sgd_rating_updates_per_sec" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and "users" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cfg = d["config"]
    assert "workload" in cfg and cfg["updates_per_step"] == cfg["iters_per_step"] * cfg["users_in_sample"]
    assert "model" not in cfg
    if cb["kind"] == "reference":  # the fixed-RNG restatement is reported beside the reference, labelled
        port = cb["port_fixed_rng"]
        assert port["kind"] == "port" and port["value"] > cb["value"] and "not the reference" in port["what"]


def test_reference_arm_is_silent_on_other_ranks():
    p = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert p.returncode == 0 and not [l for l in p.stdout.splitlines() if l.startswith("{")]
