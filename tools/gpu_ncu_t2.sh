#!/bin/bash
# ncu full capture (source level) of the type-2 hot kernel of the library named by $2 (suffix)
TAG=$1; L=$2; OUT=gpurun_out/$TAG; mkdir -p $OUT
export B2N_LIB=$PWD/jax_finufft_b200/libb200nufft$L.so
ncu --set full --clock-control none --import-source on -k regex:'k_swr2?_interp' -s 1 -c 1 -f -o $OUT/prof_t2 python bench.py --steps 1 --warmup 1 --no-extras --workload c3_t2 > $OUT/ncu_full_t2.log 2>&1
python tools/ncu_summary.py $OUT/prof_t2.ncu-rep $OUT/ncu_t2_summary.csv > $OUT/ncu_t2_summary.txt 2>&1
ncu -i $OUT/prof_t2.ncu-rep --page source --csv -k regex:'k_swr2?_interp' > $OUT/source_interp.csv 2>/dev/null
python tools/sass_hist.py $OUT/source_interp.csv 1e8 > $OUT/sass_hist_interp.txt 2>&1
rm -f $OUT/*.ncu-rep
tail -12 $OUT/ncu_t2_summary.txt
