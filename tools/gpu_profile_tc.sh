#!/bin/bash
# full ncu capture of the tensor-core stack kernel on one case: tools/gpu_profile_tc.sh <case> <tag>
set -x
CASE=${1:-stack18}; TAG=${2:-r01b}
python tools/prof_case.py $CASE > gpurun_out/${CASE}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:preact_tc -s 1 -c 1 -o gpurun_out/prof_${TAG}_${CASE} \
    python tools/prof_case.py $CASE --reps 1 > gpurun_out/ncu_${TAG}_${CASE}.log 2>&1
cat gpurun_out/${CASE}_plain.log; tail -3 gpurun_out/ncu_${TAG}_${CASE}.log
