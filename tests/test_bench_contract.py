"""bench.py's reference arm (the CPU side of the driver's ratio) runs without a GPU: one JSON line
with the contract's keys, all host cores in use whatever OMP_NUM_THREADS the launcher exported."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun gives its workers
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--cpu-sample", "200000"], capture_output=True, text=True, env=env,
                         cwd=ROOT, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "NU points/s"
    assert d["config"]["workload"] == "c3_t1" and d["config"]["M"] == 100_000_000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] > 0
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
