#!/bin/bash
# A/B of library builds on the headline workloads: gpurun -- 'bash tools/gpu_ab.sh <tag> "" _s2 ...'
TAG=$1; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
for L in "$@"; do
  LIBF=$PWD/jax_finufft_b200/libb200nufft$L.so
  for W in c3_t1 c3_t2; do
    B2N_LIB=$LIBF python bench.py --steps 5 --warmup 3 --no-extras --workload $W > $OUT/$W$L.json 2> $OUT/$W$L.err
    python - <<PY
import json
d = json.load(open("$OUT/$W$L.json"))
print("lib$L $W", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d.get("stages_ms", {}).items() if v})
PY
  done
done
