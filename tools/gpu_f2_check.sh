#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -q --timeout 900 --tb=short 2>&1 | grep -v Warning | tail -6
timeout 600 python tools/bench_configs.py c5grad 2>&1 | grep "^{" | cut -c1-200
