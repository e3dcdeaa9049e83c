#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_config.py -m gpu -q -p no:cacheprovider -k "up_block or tensor_core_stack or full_512 or bf16" 2>&1 | tail -8
python tools/prof_case.py up18_128 --batch 8 --reps 5 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline --profile-out gpurun_out/r02i_ops.tsv > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02i_bench.err
head -n 12 gpurun_out/r02i_ops.tsv
python -c "
import json; l=json.load(open('gpurun_out/r02i_bench.json')); print({k:l[k] for k in ('value','ms_per_step')}, l['e2e']['value'], l['batch1']['value'], l['extract']['value'])"
