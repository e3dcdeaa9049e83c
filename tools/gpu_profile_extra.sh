#!/bin/bash
# full ncu captures of block kernels: tools/gpu_profile_extra.sh <tag> <batch> "<case> <kernel regex>" ...
TAG=${1:-r01x}; B=${2:-2}; shift 2
[ $# -eq 0 ] && set -- "down4_512 preact_down_row_kernel" "down8_256 preact_down_row_kernel" "up18_128 preact_fused_kernel" "down16_128 conv3d_tc_kernel"
for spec in "$@"; do
  set -- $spec
  python tools/prof_case.py $1 --batch $B > gpurun_out/$1_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -o gpurun_out/prof_${TAG}_$1 \
      python tools/prof_case.py $1 --reps 1 --batch $B > gpurun_out/ncu_${TAG}_$1.log 2>&1
  cat gpurun_out/$1_plain.log; tail -1 gpurun_out/ncu_${TAG}_$1.log
done
