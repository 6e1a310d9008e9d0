#!/bin/bash
OUT=gpurun_out/r02g; mkdir -p $OUT
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "binsort or sort" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $OUT/memcheck.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 --tb=short 2>&1 | grep -v Warning | tail -12
timeout 300 python tools/stage_times.py 2>&1 | tail -1
B2N_SORT_GLOBAL_HIST=1 timeout 300 python tools/stage_times.py 2>&1 | tail -1
timeout 300 python tools/stage_times.py 1e8 clustered 2>&1 | tail -1
B2N_SORT_GLOBAL_HIST=1 timeout 300 python tools/stage_times.py 1e8 clustered 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_t1.csv python bench.py --steps 1 --warmup 1 --no-extras > $OUT/ncu_t1.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("$OUT/launches_t1.csv")))
h=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
H=rows[h]; kn=H.index("Kernel Name"); mv=H.index("Metric Value")
seen=[]
for r in rows[h+2:]:
    if len(r)>mv: seen.append((r[kn][:60], r[mv]))
for k,v in seen[-22:]: print(k, v)
PY
