#!/bin/bash
# round-1 baseline evidence: launch list of one eager forward + full capture of the 18-ch stack kernel
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/ops_r01a.tsv > gpurun_out/bench_r01a.log 2> gpurun_out/bench_r01a.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 527 -c 527 --csv --log-file gpurun_out/launches_r01a.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r01a.log 2>&1
python tools/prof_case.py stack18 > gpurun_out/stack18_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:preact_fused -s 4 -c 2 -o gpurun_out/prof_r01a_stack18 \
    python tools/prof_case.py stack18 > gpurun_out/ncu_r01a_stack18.log 2>&1
cat gpurun_out/bench_r01a.log | cut -c1-600; cat gpurun_out/stack18_plain.log; tail -3 gpurun_out/ncu_r01a.log; tail -3 gpurun_out/ncu_r01a_stack18.log
