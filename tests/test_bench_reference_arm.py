"""``bench.py --impl reference`` on the host cores (no GPU needed): the JSON line the driver parses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rendered IRs/sec (fwd+bwd)" and d["unit"] == "IR/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] >= 1 and d["value"] > 0
    # nothing extrapolated: value = receivers * steps / measured seconds of THIS run, and config says one receiver per step
    assert d["config"]["receivers_per_gpu"] == 1 and d["config"]["rays"] == 2050
    assert abs(d["value"] - d["steps"] * 1000.0 / (d["ms_per_step"] * d["steps"])) < 1e-9 * max(1.0, d["value"]) + 1e-6
    assert d["e2e"] == {"value": d["value"], "unit": "IR/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
