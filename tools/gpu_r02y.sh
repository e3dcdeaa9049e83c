#!/bin/bash
mkdir -p gpurun_out
VQ3D_FUSED_BLOCK_BWD=1 python tools/prof_train.py --workload downscaled_256x256x128 2>&1 | grep -E "preact_same_backward|preact_block" | head -12
for shp in "18 18 same 64 64 32" "8 8 same 16 16 8" "2 2 same 64 64 32"; do
VQ3D_FUSED_BLOCK_BWD=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02y_l.csv python tools/prof_bwd.py $shp 2 > /dev/null 2>&1
python - <<P
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02y_l.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
print("$shp", [(r[ki].split('(')[0][-30:], int(float(r[vi]))) for r in rows[-8:] if 'vq3d' in r[ki]])
P
done
