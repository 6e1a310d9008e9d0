#!/bin/bash
# One GPU-box visit: tests, bench, ncu launch lists + full captures summarised ON the box (the
# .ncu-rep files stay there; CSV summaries come back).  Usage: gpurun -- 'bash tools/gpu_round.sh <tag> [notest] [noncu]'
TAG=${1:-r02}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv > $OUT/pcie.txt 2>&1
lscpu | head -20 >> $OUT/pcie.txt
if [ "$2" != "notest" ]; then
  python -m pytest tests -m gpu -q --timeout 1200 --tb=short -s 2>&1 | grep -v Warning > $OUT/pytest_gpu_full.log
  grep -E "PARITY|passed|failed|error" $OUT/pytest_gpu_full.log | tail -60 > $OUT/pytest_gpu.log
  tail -40 $OUT/pytest_gpu_full.log | cut -c1-300
fi
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json; tail -3 $OUT/bench.err
python bench.py --steps 5 --warmup 3 --workload c3_t2 > $OUT/bench_t2.json 2> $OUT/bench_t2.err
if [ "$3" != "noncu" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_t1.csv python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_t1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_t2.csv python bench.py --steps 2 --warmup 1 --no-extras --workload c3_t2 > $OUT/ncu_t2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_swr2?_spread|k_key_hist|k_partition|k_place<' -s 4 -c 4 -f -o $OUT/prof_t1 python bench.py --steps 1 --warmup 1 --no-extras > $OUT/ncu_full_t1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_swr2?_interp|k_amplify' -s 2 -c 2 -f -o $OUT/prof_t2 python bench.py --steps 1 --warmup 1 --no-extras --workload c3_t2 > $OUT/ncu_full_t2.log 2>&1
python tools/ncu_summary.py $OUT/prof_t1.ncu-rep $OUT/ncu_t1_summary.csv > $OUT/ncu_t1_summary.txt 2>&1
python tools/ncu_summary.py $OUT/prof_t2.ncu-rep $OUT/ncu_t2_summary.csv > $OUT/ncu_t2_summary.txt 2>&1
python tools/traffic_stamp.py $TAG $OUT/roofline_traffic.json "c3_t1:spread:k_swr2?_spread:$OUT/ncu_t1_summary.csv" "c3_t2:interp:k_swr2?_interp:$OUT/ncu_t2_summary.csv" > /dev/null 2>&1
ncu -i $OUT/prof_t1.ncu-rep --page source --csv -k regex:'k_swr2?_spread' > $OUT/source_spread.csv 2>/dev/null
ncu -i $OUT/prof_t2.ncu-rep --page source --csv -k regex:'k_swr2?_interp' > $OUT/source_interp.csv 2>/dev/null
python tools/sass_hist.py $OUT/source_spread.csv 1e8 > $OUT/sass_hist_spread.txt 2>&1
python tools/sass_hist.py $OUT/source_interp.csv 1e8 > $OUT/sass_hist_interp.txt 2>&1
rm -f $OUT/*.ncu-rep
fi
ls -la $OUT
