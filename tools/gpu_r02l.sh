#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_callers.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6
python tools/prof_train.py --workload downscaled_256x256x128 2>&1 | grep -E "wall|conv3d|conv1x1" | head -8
python tools/prof_train.py --workload full_512x512x128 2>&1 | grep -E "wall|conv3d|conv1x1" | head -8
python tools/bench_train.py --workload downscaled_256x256x128 --graph --steps 5 | tail -1
python tools/bench_train.py --workload full_512x512x128 --graph --steps 3 | tail -1
