"""bench.py's reference arm runs without a GPU: its JSON line carries what the driver reads (metric, unit, config of the
GPU arm, impl = reference, a cpu_baseline describing the run, an e2e block without copies)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line(tmp_path):
    run = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert run.returncode == 0, run.stderr[-2000:]
    lines = [ln for ln in run.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["steps"] == 1 and line["higher_is_better"] is True
    assert line["metric"].startswith("fragments/sec") and line["unit"] == "fragments/s" and line["value"] > 0
    assert line["config"]["workload"].startswith("BASELINE.json configs[1]") and line["config"]["fragments"] == 250000000
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] == line["value"] and "fragments" in cpu["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None and line["data"] == "synthetic"
