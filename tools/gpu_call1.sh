#!/bin/bash
# round-2 first visit: memcheck of the new kernels on a small case, GPU tests, A/B of kernel generations
OUT=gpurun_out/r02a; mkdir -p $OUT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sweep.py -q -x -k "16x16x16 and 1e-06" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $OUT/memcheck.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 --tb=short -s 2>&1 | grep -v Warning > $OUT/pytest_gpu_full.log
grep -E "PARITY|passed|failed|rror" $OUT/pytest_gpu_full.log | tail -400 > $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu_full.log | cut -c1-250
timeout 900 bash tools/gpu_ab.sh r02a_ab "" _g1 _st0
