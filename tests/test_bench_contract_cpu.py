"""CPU: the reference arm of bench.py (`--impl reference`: the reference's own CPU path, oracle/_ref
when it was built here, else the oracle port) prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    run = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert run.returncode == 0, run.stderr
    lines = [ln for ln in run.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("option PDE solves/sec")
    assert d["unit"] == "solves/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["nodes"] == 2049 and d["config"]["time_steps"] == 1000
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_of_the_hjbqvi_configs():
    """--impl reference --config er / oc: the reference's own HJBQVI<..>::solve on a bounded sample."""
    for cfg, extra in (("er", ["--cpu-sample", "2", "--hjbqvi-refine", "2"]), ("oc", [])):
        run = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", cfg,
                              "--steps", "1", "--warmup", "0"] + extra, capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert run.returncode == 0, run.stderr
        lines = [ln for ln in run.stdout.splitlines() if ln.strip()]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference"
        assert ("exchange_rate" if cfg == "er" else "optimal_consumption") in d["config"]["workload"]
