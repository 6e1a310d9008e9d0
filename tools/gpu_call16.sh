#!/bin/bash
B2N_LIB=$PWD/jax_finufft_b200/libb200nufft_yci.so timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -q --timeout 1200 --tb=short 2>&1 | grep -v Warning | tail -2
B2N_LIB=$PWD/jax_finufft_b200/libb200nufft_yci.so timeout 300 python tools/stage_times.py 2>&1 | tail -1
