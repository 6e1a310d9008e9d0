#!/bin/bash
OUT=gpurun_out/r02j; mkdir -p $OUT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sweep.py -q -x -k "stacked" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 --tb=short -s 2>&1 | grep -v Warning | grep "PARITY C4\|passed\|failed\|rror" | tail -8
timeout 600 python tools/bench_configs.py c4 2>&1 | grep "^{" | cut -c1-400
