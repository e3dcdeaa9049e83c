#!/bin/bash
# full ncu capture of the up-block kernels at batch 8
mkdir -p gpurun_out
python tools/prof_case.py up18_128 --batch 8 --reps 1 > gpurun_out/r02g_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'up_expand|preact_tc_kernel|up_lo' -s 3 -c 3 -f -o gpurun_out/r02g_up18 python tools/prof_case.py up18_128 --batch 8 --reps 1 > gpurun_out/r02g_ncu.log 2>&1
echo rc=$?; tail -3 gpurun_out/r02g_ncu.log
