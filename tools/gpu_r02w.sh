#!/bin/bash
mkdir -p gpurun_out
VQ3D_FUSED_BLOCK_BWD=1 python tools/prof_bwd.py 18 18 same 64 64 32 2 > gpurun_out/r02w_plain.log 2>&1 &&
VQ3D_FUSED_BLOCK_BWD=1 ncu --set full --clock-control none --import-source on -k regex:preact_same_bwd -s 2 -c 2 -f -o gpurun_out/r02w_same_bwd python tools/prof_bwd.py 18 18 same 64 64 32 2 > gpurun_out/r02w_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/r02w_ncu.log
