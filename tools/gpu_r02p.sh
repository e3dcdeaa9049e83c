#!/bin/bash
mkdir -p gpurun_out
python tools/prof_bwd.py 4 8 down 512 512 128 2 > gpurun_out/r02p_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02p_down4_bwd_launches.csv python tools/prof_bwd.py 4 8 down 512 512 128 2 > gpurun_out/r02p_ncu.log 2>&1
echo rc=$?
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02p_down4_bwd_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
half=rows[1+len(rows[1:])//2:]
for r in half: print(r[ki][:90], r[vi])
P
