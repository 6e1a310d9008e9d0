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
                          "--warmup", "1", "--cpu-sample", "200000", "--no-cpu-full"], capture_output=True, text=True, env=env,
                         cwd=ROOT, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "NU points/s"
    # the driver's run times the full M = 1e8 per step; this CPU-suite smoke run bounds it and says so
    assert d["config"]["workload"] == "c3_t1" and d["config"]["M_per_gpu"] == 200_000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] > 0 and cb["extrapolated"] is False
    assert abs(d["value"] - d["config"]["M_per_gpu"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_both_arms_describe_the_workload_with_the_same_config():
    sys.path.insert(0, ROOT)
    import bench

    cfg = bench.make_config("c3_t1", 1)
    assert cfg == {"workload": "c3_t1", "type": 1, "M_per_gpu": 100_000_000, "N": [256, 256, 256], "eps": 1e-6,
                   "points": "uniform", "seed": "1+rank", "n_gpus": 1}
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("make_config(a.workload,") == 2   # main_ours and main_reference
    assert len(bench.src_sha16()) == 16
