#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/microbench/tc_microbench > gpurun_out/r02_tc_microbench.txt 2>&1; echo "microbench rc=$?"
cat gpurun_out/r02_tc_microbench.txt
timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_full_config.py -m gpu -q -p no:cacheprovider -s -k "bf16_mode or batch8" > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r02b_pytest.log | grep -n "got\|rel \|worst\|passed\|failed" | head -150
