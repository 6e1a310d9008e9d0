#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runtime.py -m gpu -q --timeout 1200 --tb=short 2>&1 | grep -v Warning | tail -5
for L in "" _nomatch; do B2N_LIB=$PWD/jax_finufft_b200/libb200nufft$L.so timeout 300 python tools/stage_times.py 2>&1 | tail -1; done
