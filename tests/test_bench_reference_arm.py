"""bench.py --impl reference runs the reference's own CPU code (oracle/_ref, else the oracle port)
and prints one JSON line with the contract's keys — checked here on a tiny sample (CPU only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    env = dict(os.environ, W2B_REF_CAL_TOKENS="30000", W2B_REF_MAX_TOKENS="60000", W2B_REF_MIN_TOKENS="40000",
               W2B_REF_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "words/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 4
    assert d["e2e"] == {"value": d["value"], "unit": "words/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "bitlevel=1" in d["config"]["workload"] and "size=800" in d["config"]["workload"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
