#!/bin/bash
# ncu full capture of selected kernels.  Usage: gpurun -- 'bash tools/gpu_ncu.sh <tag> <kernel-regex> [workload] [skip] [count]'
TAG=${1:-n}; RE=${2:-k_swr_spread}; WL=${3:-c3_t1}; SKIP=${4:-1}; CNT=${5:-1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o $OUT/prof python bench.py --steps 1 --warmup 1 --no-extras --workload $WL > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log; ls -la $OUT
