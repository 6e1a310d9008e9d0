#!/bin/bash
# Multi-GPU visit: H2D probe at 1..N ranks, NCCL tests, bench at the rank counts given.
# Usage: gpurun --gpus N -- 'bash tools/gpu_multi2.sh <tag> "<probe ranks>" "<bench ranks>"'
TAG=${1:-m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for n in $2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 tools/h2d_probe_multi.py 2>/dev/null | tail -1 | tee -a $OUT/h2d_probe.jsonl
done
python -m pytest tests/test_gpu_parallel.py -m gpu -q --timeout 900 --tb=short 2>&1 | grep -v Warning | tail -6 | tee $OUT/pytest_parallel.log
for n in $3; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 5 --warmup 3 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  python - <<PY
import json
d = json.load(open("$OUT/bench_n$n.json"))
print("N=$n", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2), {k: (round(v.get("ms_per_step", 0), 3) if "ms_per_step" in v else {m: round(x.get("ms_per_step", -1), 3) if isinstance(x, dict) else x for m, x in v.items()}) for k, v in d["also"].items() if k.startswith(("c4", "c5", "c3_t1_points", "t1_n512"))})
PY
  tail -2 $OUT/bench_n$n.err
done
