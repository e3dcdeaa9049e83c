#!/bin/bash
# N GPUs (torchrun): the bench line incl. the data-parallel C3 training step and the copy-only host<->device ceiling
N=${1:-8}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02x_bench_${N}gpu.json 2> gpurun_out/r02x_bench_${N}gpu.err; echo "bench$N rc=$?"
tail -c 1500 gpurun_out/r02x_bench_${N}gpu.err
python - <<P
import json
l=json.loads(open('gpurun_out/r02x_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', l['e2e']['value'], 'copy_only', l['e2e']['copy_only'], 'hu', l['e2e_hu_int16']['value'], 'extract e2e', l['extract']['e2e']['value'])
print('train', json.dumps(l.get('train_step'))[:1200])
P
