#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/microbench/tc_microbench > gpurun_out/r02_tc_microbench.txt 2>&1; echo "microbench rc=$?"
grep "umma" gpurun_out/r02_tc_microbench.txt
