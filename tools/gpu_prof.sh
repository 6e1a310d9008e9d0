#!/bin/bash
# ncu visit: launch lists + full captures of the dominant kernels.  Usage: gpurun -- 'bash tools/gpu_prof.sh <tag>'
TAG=${1:-p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_t1.csv python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_t1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_t2.csv python bench.py --steps 2 --warmup 1 --no-extras --workload c3_t2 > $OUT/ncu_t2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_swr_spread|k_key_hist|k_partition|k_place' -s 4 -c 4 -f -o $OUT/prof_t1 python bench.py --steps 1 --warmup 1 --no-extras > $OUT/ncu_full_t1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_swr_interp|k_amplify' -s 2 -c 2 -f -o $OUT/prof_t2 python bench.py --steps 1 --warmup 1 --no-extras --workload c3_t2 > $OUT/ncu_full_t2.log 2>&1
ls -la $OUT
