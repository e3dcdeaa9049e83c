#!/bin/bash
# 2 GPUs: the bench line under torchrun (C3 training step with the NCCL gradient + EMA all-reduces inside the captured step)
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02k_bench_2gpu.json 2> gpurun_out/r02k_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 2500 gpurun_out/r02k_bench_2gpu.err
python - <<'P'
import json
try:
    l=json.loads(open('gpurun_out/r02k_bench_2gpu.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', l['e2e']); print('hu', l['e2e_hu_int16']['value']); print('train', json.dumps(l.get('train_step'))[:1500])
except Exception as e: print('parse error', e)
P
