#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` report of the bench's dominant op (run here; no GPU needed).

    python tools/ncu_traffic.py gpurun_out/x.ncu-rep "<kernel name regex>" "<op tag as bench.py prints it>" <launches per op> [out.json]

Sums dram__bytes_read.sum + dram__bytes_write.sum over the captured launches of the op (the 50-block stack is 3 launches:
24 + 24 + 2 blocks) and writes {"<op tag>": {"dram_bytes_per_launch": ..., "source": ...}} -- `launch` in bench.py's roofline is
one call of the op (all of its kernel launches), like `achieved`."""
import csv, io, json, os, re, subprocess, sys

rep, kre, tag, per_op = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
out = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
sel = [r for r in data if re.search(kre, r[ki])]
assert sel and len(sel) % per_op == 0, (len(sel), per_op)
rd = sum(float(r[ri]) * scale[units[ri]] for r in sel) / (len(sel) // per_op)
wr = sum(float(r[wi]) * scale[units[wi]] for r in sel) / (len(sel) // per_op)
tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[ti]]
ms = sum(float(r[ti]) * tscale for r in sel) / (len(sel) // per_op)
table = {}
if os.path.exists(out):
    try:
        table = json.load(open(out))
    except Exception:
        table = {}
table[tag] = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_written": wr, "kernel_launches_per_op": per_op, "ncu_ms_per_op": ms,
              "source": f"{os.path.basename(rep)} (ncu --set full --clock-control none on bench.py itself at its default batch): sum over the {per_op} kernel launches of one call of the op"}
json.dump(table, open(out, "w"), indent=1)
print(json.dumps(table[tag], indent=1))
