#!/bin/bash
# full ncu capture of one prof_case: tools/gpu_profile_case.sh <case> <kernel regex> <tag> [precision]
set -x
CASE=$1; KRE=$2; TAG=$3; PREC=${4:-bf16}
python tools/prof_case.py $CASE --precision $PREC > gpurun_out/${CASE}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -o gpurun_out/prof_${TAG}_${CASE} \
    python tools/prof_case.py $CASE --reps 1 --precision $PREC > gpurun_out/ncu_${TAG}_${CASE}.log 2>&1
cat gpurun_out/${CASE}_plain.log; tail -2 gpurun_out/ncu_${TAG}_${CASE}.log
