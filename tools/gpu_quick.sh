#!/bin/bash
# Quick GPU visit: tests (optional) + bench with stage timings.  Usage: gpurun -- 'bash tools/gpu_quick.sh <tag> [notest]'
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$2" != "notest" ]; then
  python -m pytest tests -m gpu -q -x --timeout 900 --tb=short 2>&1 | grep -v Warning | tail -25 > $OUT/pytest_gpu.log
  tail -8 $OUT/pytest_gpu.log
fi
python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json; tail -5 $OUT/bench.err
