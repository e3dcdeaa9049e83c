#!/bin/bash
# launch list (per-kernel device time) of the tensor-core 'up' block at the bench's batch 8
mkdir -p gpurun_out
python tools/prof_case.py up18_128 --batch 8 --reps 2 > gpurun_out/r02e_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02e_up18_launches.csv python tools/prof_case.py up18_128 --batch 8 --reps 2 > gpurun_out/r02e_ncu.log 2>&1
echo rc=$?
cat gpurun_out/r02e_plain.log | tail -3
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02e_up18_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ki][:70], r[vi])
P
