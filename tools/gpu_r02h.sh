#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "up_block" 2>&1 | tail -3
python tools/prof_case.py up18_128 --batch 8 --reps 5 | tail -1
python tools/prof_case.py up32_64 --batch 8 --reps 5 | tail -1
python tools/prof_vq.py 1048576 32 512 > gpurun_out/r02h_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02h_vq_1Mx32x512 python tools/prof_vq.py 1048576 32 512 > gpurun_out/r02h_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/r02h_ncu.log
