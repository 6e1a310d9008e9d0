#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -q --timeout 600 --tb=short 2>&1 | grep -v Warning | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02n_bench_n2.json 2> gpurun_out/r02n_bench_n2.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02n_bench_n2.json"))
print("N=2", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), {k: (round(v.get("ms_per_step", 0), 3) if "ms_per_step" in v else {m: round(x.get("ms_per_step", -1), 3) if isinstance(x, dict) else x for m, x in v.items()}) for k, v in d["also"].items() if k.startswith(("c4", "c5", "c3_t1_points", "t1_n512"))})
PY
tail -2 gpurun_out/r02n_bench_n2.err
