"""bench.py contract checks that need no GPU: the reference arm (the oracle port on host cores) prints exactly one
JSON line on stdout with the keys the driver reads, and the B200 arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "7", "--warmup", "4",
                        "--batch", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bc_train_steps_per_sec" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["config"]["workload"].startswith("cfg2") and d["config"]["global_batch"] == 2
    # the arm reports what it ran: --steps / --warmup honoured up to the stated caps, ms_per_step unscaled
    assert d["steps"] == min(7, d["config"]["steps_cap"]) and d["warmup"] == min(4, d["config"]["warmup_cap"])
    assert abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_b200_arm_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the arm runs for real (driver); nothing to check here
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")]
