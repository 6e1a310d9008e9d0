#!/bin/bash
OUT=gpurun_out/r02h; mkdir -p $OUT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sweep.py -q -x -k "16x16x16 and 1e-06" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck.log
for L in "" _mb18 _mb20; do B2N_LIB=$PWD/jax_finufft_b200/libb200nufft$L.so timeout 300 python tools/stage_times.py 2>&1 | tail -1; done
B2N_LIB=$PWD/jax_finufft_b200/libb200nufft_mb20.so timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 --tb=short 2>&1 | grep -v Warning | tail -6
