#!/bin/bash
# launch list of the bench's forward step (ncu --metrics gpu__time_duration.sum): tools/gpu_launch_list.sh <tag> [skip] [count]
# (one ncu pass per gpurun call; the plain run comes first)
TAG=${1:-r02w}; SKIP=${2:-400}; CNT=${3:-360}
VQ3D_BENCH_NO_GRAPH=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/bench_${TAG}_plain.log 2> gpurun_out/bench_${TAG}_plain.err || exit 1
tail -c 300 gpurun_out/bench_${TAG}_plain.log
VQ3D_BENCH_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $CNT --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log | cut -c1-300
wc -l gpurun_out/launches_${TAG}.csv
