#!/bin/bash
for L in "" _s4 _nt4 _s4nt4; do echo "lib$L"; B2N_LIB=$PWD/jax_finufft_b200/libb200nufft$L.so timeout 600 python tools/bench_configs.py c4 2>&1 | grep "^{" | cut -c60-330; done
B2N_LIB=$PWD/jax_finufft_b200/libb200nufft_s4.so timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -q -k stacked 2>&1 | tail -2
