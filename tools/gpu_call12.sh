#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 --tb=short 2>&1 | grep -v Warning | tail -5
timeout 300 python tools/stage_times.py 2>&1 | tail -1
timeout 300 python tools/stage_times.py 1e8 clustered 2>&1 | tail -1
timeout 600 python tools/bench_configs.py c1 c2 c4 c5 c5grad 2>&1 | cut -c1-400 | tee gpurun_out/r02i_configs.jsonl
