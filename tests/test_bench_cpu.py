"""The reference arm of bench.py (`--impl reference`: the oracle port on the host cores, no GPU) prints ONE JSON line that
keeps the driver's contract: metric / unit / value of the CUDA arm's headline, `impl`, a `cpu_baseline` describing the
run, an `e2e` block with zero copied bytes, and a `config` that states the sample it actually ran."""
import json
import os
import subprocess
import sys


def test_reference_arm_line():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample", "1500"], capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["n_gpus"] == 1 and rec["steps"] == 1
    assert rec["unit"] == "particles/s" and rec["value"] > 0 and rec["higher_is_better"] is True
    assert rec["metric"].startswith("particles/sec")
    assert rec["e2e"] == {"value": rec["value"], "unit": rec["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = rec["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == rec["value"] and "1500-particle" in cb["sample"]
    assert rec["config"]["particles_per_gpu"] == 1500 and "bounded sample of 1500 particles" in rec["config"]["workload"]
    assert rec["gpu_launches"] == 0
