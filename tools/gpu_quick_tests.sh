#!/bin/bash
# usage: gpurun -- 'bash tools/gpu_quick_tests.sh "<pytest args>"'
timeout 1500 python -m pytest $1 -m gpu -q --timeout 1200 --tb=short -s 2>&1 | grep -v Warning | grep -E "PARITY sigma|passed|failed|rror|assert" | tail -20
