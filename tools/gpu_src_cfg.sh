#!/bin/bash
# Source-level ncu capture of one kernel while tools/bench_configs.py runs a config.
# Usage: gpurun -- 'bash tools/gpu_src_cfg.sh <tag> <kernel-regex> <config> [skip]'
TAG=${1:-s}; RE=${2:-k_rt2_spread}; CFG=${3:-c4}; SKIP=${4:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c 1 -f -o $OUT/prof python tools/bench_configs.py $CFG > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/source.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
rm -f $OUT/prof.ncu-rep
