#!/bin/bash
# round-end evidence at the bench's batch: launch list of one eager forward + full captures of the dominant kernels
set -x
TAG=${1:-r01z}
B=${2:-8}
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/bench_${TAG}_plain.log 2> gpurun_out/bench_${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/ncu_${TAG}.log 2>&1
python tools/prof_case.py stack18 --batch $B > gpurun_out/stack18_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:preact_tc_kernel -s 1 -c 1 -o gpurun_out/prof_${TAG}_stack18 \
    python tools/prof_case.py stack18 --reps 1 --batch $B > gpurun_out/ncu_${TAG}_stack18.log 2>&1
python tools/prof_case.py stack4_512 --batch $B > gpurun_out/stack4_512_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:preact_row_kernel -s 1 -c 1 -o gpurun_out/prof_${TAG}_stack4_512 \
    python tools/prof_case.py stack4_512 --reps 1 --batch $B > gpurun_out/ncu_${TAG}_stack4_512.log 2>&1
cat gpurun_out/stack18_plain.log gpurun_out/stack4_512_plain.log; tail -2 gpurun_out/ncu_${TAG}_stack18.log
