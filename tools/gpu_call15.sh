#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sweep.py tests/test_gpu_parity.py -m gpu -q --timeout 1200 --tb=short 2>&1 | grep -v Warning | tail -3
timeout 300 python tools/stage_times.py 2>&1 | tail -1
timeout 300 python tools/stage_times.py 1e8 clustered 2>&1 | tail -1
