#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 16 gpurun_out/r02x_pytest.log
( time timeout 900 python bench.py --profile-out gpurun_out/r02x_ops.tsv > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -c 800 gpurun_out/r02x_bench.err
python - <<'P'
import json
l=json.loads(open('gpurun_out/r02x_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','steps')}); print('e2e', l['e2e']['value'], 'copy_only', l['e2e']['copy_only']['value'], 'hu', l['e2e_hu_int16']['value'], 'batch1', l['batch1']['value'])
print('roofline', {k:l['roofline'][k] for k in ('achieved','frac','traffic','avg_launch_us','share_of_step')})
print('extract', l['extract']['value'], l['extract']['e2e']['value'], l['extract'].get('index_mismatch_vs_oracle'))
print('quantizer', l['quantizer']['value'], l['quantizer']['second_point']['value'])
print('train', {k:(v.get('ms_per_step'), v.get('cuda_graph')) for k,v in l['train_step'].items()})
print('cpu', l['cpu_baseline'])
P
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02x_ref.json 2> gpurun_out/r02x_ref.err ) 2>&1 | grep real
cat gpurun_out/r02x_ref.json | head -c 1200
