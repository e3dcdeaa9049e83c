#!/bin/bash
# DRAM traffic of the dominant op (the 50-block 18-channel stack = 3 kernel launches) measured on bench.py itself
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02r_plain.json 2> gpurun_out/r02r_plain.err &&
ncu --set full --clock-control none --kernel-name-base mangled -k regex:preact_tc_kernelILi18ELi9 -s 3 -c 3 -f -o gpurun_out/r02r_bench_stack50 python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02r_ncu.log 2>&1
echo rc=$?; tail -3 gpurun_out/r02r_ncu.log
