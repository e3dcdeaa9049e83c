#!/bin/bash
# round 2, first GPU pass: the whole -m gpu suite (new oracle-pinned Full-config tests included), the bench line, the reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
nproc >> gpurun_out/r02a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=30 -p no:cacheprovider -s > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 60 gpurun_out/r02a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/r02a_ops.tsv > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r02a_bench.err
head -c 6000 gpurun_out/r02a_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err; echo "ref rc=$?"
cat gpurun_out/r02a_ref.json
