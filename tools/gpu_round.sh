#!/bin/bash
# One GPU-box visit: tests, bench, ncu launch lists + full captures summarised ON the box (the
# .ncu-rep files stay there; CSV summaries come back).  Usage: gpurun -- 'bash tools/gpu_round.sh <tag> [notest]'
TAG=${1:-r01}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv > $OUT/pcie.txt 2>&1
lscpu | head -20 >> $OUT/pcie.txt
if [ "$2" != "notest" ]; then
  python -m pytest tests -m gpu -q --timeout 900 --tb=short 2>&1 | grep -v Warning | tail -15 > $OUT/pytest_gpu.log
  tail -3 $OUT/pytest_gpu.log
fi
python tools/h2d_probe.py > $OUT/h2d.txt 2>&1; cat $OUT/h2d.txt
python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json; tail -3 $OUT/bench.err
python bench.py --steps 5 --warmup 3 --workload c3_t2 > $OUT/bench_t2.json 2> $OUT/bench_t2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_t1.csv python bench.py --steps 2 --warmup 1 --no-extras > $OUT/ncu_t1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_t2.csv python bench.py --steps 2 --warmup 1 --no-extras --workload c3_t2 > $OUT/ncu_t2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_swr_spread|k_key_hist|k_partition|k_place<' -s 4 -c 4 -f -o $OUT/prof_t1 python bench.py --steps 1 --warmup 1 --no-extras > $OUT/ncu_full_t1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_swr_interp|k_amplify' -s 2 -c 2 -f -o $OUT/prof_t2 python bench.py --steps 1 --warmup 1 --no-extras --workload c3_t2 > $OUT/ncu_full_t2.log 2>&1
python tools/ncu_summary.py $OUT/prof_t1.ncu-rep $OUT/ncu_t1_summary.csv > $OUT/ncu_t1_summary.txt 2>&1
python tools/ncu_summary.py $OUT/prof_t2.ncu-rep $OUT/ncu_t2_summary.csv > $OUT/ncu_t2_summary.txt 2>&1
ncu -i $OUT/prof_t1.ncu-rep --page source --csv -k regex:k_swr_spread > $OUT/source_spread.csv 2>/dev/null
ncu -i $OUT/prof_t2.ncu-rep --page source --csv -k regex:k_swr_interp > $OUT/source_interp.csv 2>/dev/null
rm -f $OUT/*.ncu-rep
ls -la $OUT
