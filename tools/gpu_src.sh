#!/bin/bash
# Source-level ncu capture of one kernel.  Usage: gpurun -- 'bash tools/gpu_src.sh <tag> <kernel-regex> [workload] [env assignments...]'
TAG=${1:-s}; RE=${2:-k_swr_spread}; WL=${3:-c3_t1}; shift 3; OUT=gpurun_out/$TAG; mkdir -p $OUT
for kv in "$@"; do export "$kv"; done
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 1 -c 1 -f -o $OUT/prof python bench.py --steps 1 --warmup 1 --no-extras --workload $WL > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/source.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
rm -f $OUT/prof.ncu-rep   # the CSV pages are what gets read; the report itself is tens of MB
ls -la $OUT
