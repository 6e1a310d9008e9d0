#!/bin/bash
OUT=gpurun_out/${1:-sortprof}; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:'k_key_hist|k_partition|k_place' -s 3 -c 3 -f -o $OUT/prof python bench.py --steps 1 --warmup 1 --no-extras > $OUT/ncu.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
for k in k_key_hist k_partition k_place; do ncu -i $OUT/prof.ncu-rep --page source --csv -k regex:$k > $OUT/source_$k.csv 2>/dev/null; done
rm -f $OUT/prof.ncu-rep; ls -la $OUT
