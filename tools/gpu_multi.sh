#!/bin/bash
# Multi-GPU visit: NCCL tests + bench at N ranks.  Usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh <tag> <N>'
TAG=${1:-m}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
python -m pytest tests/test_gpu_parallel.py -m gpu -q --timeout 900 --tb=short 2>&1 | grep -v Warning | tail -15 > $OUT/pytest_parallel.log
tail -5 $OUT/pytest_parallel.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
cat $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
