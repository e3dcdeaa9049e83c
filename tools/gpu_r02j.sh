#!/bin/bash
mkdir -p gpurun_out
python tools/prof_case.py stack18 --batch 8 --reps 1 > gpurun_out/r02j_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:preact_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02j_stack18_b8 python tools/prof_case.py stack18 --batch 8 --reps 1 > gpurun_out/r02j_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/r02j_ncu.log
