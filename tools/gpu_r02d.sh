#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "up_block or tensor_core_stack" > gpurun_out/r02d_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 40 gpurun_out/r02d_pytest.log
VQ3D_TC_DEBUG=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-train --no-cpu-baseline --profile-out gpurun_out/r02d_ops.tsv > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"
grep "preact_up_tc" gpurun_out/r02d_bench.err | sort | uniq | head
tail -c 1500 gpurun_out/r02d_bench.err
head -n 30 gpurun_out/r02d_ops.tsv
python -c "
import json; l=json.load(open('gpurun_out/r02d_bench.json')); print({k:l[k] for k in ('value','ms_per_step')}, l['e2e']['value'], l['batch1'], l['extract']['value'])"
