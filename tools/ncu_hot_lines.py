#!/usr/bin/env python
"""Maps the SASS page of an ncu report (--import-source on) back to source lines with nvdisasm -g line markers.

    python tools/ncu_hot_lines.py report.ncu-rep build/file.o mangled_kernel_name [top]
"""
import csv, io, os, re, subprocess, sys, tempfile
rep, obj, fun = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
# HOT_OUTER=1: attribute inlined code (mbarrier waits, ELU helpers ...) to the line of the kernel body that called it
dis = subprocess.run(["nvdisasm", "-gi" if os.environ.get("HOT_OUTER") else "-g", cubin], capture_output=True, text=True).stdout
dis = dis[dis.index(".text." + fun + ":"):]
nxt = dis.find("\n.text.", 10)
nxt2 = dis.find("\n\t.section", 10)
end = min(x for x in (nxt, nxt2, len(dis)) if x > 0)
dis = dis[:end]
lines = []          # source line of the i-th instruction (outermost inlining site = the line in the kernel body)
cur = 0
for l in dis.splitlines():
    ms = re.findall(r'"[^"]*?([^/"]+)", line (\d+)', l) if "//## File" in l else None
    if ms:
        cur = (ms[-1][0], int(ms[-1][1]))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h = rows[1]
si, ii, so = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
stall_cols = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
data = rows[2:]
if len(data) > len(lines):          # several kernels in the report: KERNEL_INDEX-th block of this function's length (default the first)
    k0 = int(os.environ.get("KERNEL_OFFSET", "0"))
    data = data[k0:k0 + len(lines)]
assert len(data) == len(lines), (len(data), len(lines))
agg = {}
for r, ln in zip(data, lines):
    a = agg.setdefault(ln, [0, 0, {}])
    a[0] += int(r[si]); a[1] += int(r[ii])
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            a[2][h[c]] = a[2].get(h[c], 0) + v
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print(f"total samples {ts}, warp instructions {ti}")
src_cache = {}
def src(ln):
    f, n = ln
    for root in (os.path.dirname(os.path.abspath(obj)) + "/../csrc", "."):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:90]
    return ""
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ", ".join(f"{k[6:]}:{v}" for k, v in sorted(a[2].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100 * a[0] / ts:5.1f}% smp {100 * a[1] / ti:5.1f}% inst  {ln[0]}:{ln[1]:<4d} [{st}]  {src(ln)}")
