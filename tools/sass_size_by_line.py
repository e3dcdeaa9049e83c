#!/usr/bin/env python
"""SASS instruction count per source line (innermost and outermost inlining site) of one kernel: code-size / icache budget.
    python tools/sass_size_by_line.py build/file.o mangled_name"""
import os, re, subprocess, sys, tempfile, collections
obj, fun = sys.argv[1:3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
dis = dis[dis.index(".text." + fun + ":"):]
ends = [x for x in (dis.find("\n.text.", 10), dis.find("\n\t.section", 10)) if x > 0]
dis = dis[:min(ends)] if ends else dis
cnt = collections.Counter(); cur = None; n = 0
for l in dis.splitlines():
    if "//## File" in l:
        ms = re.findall(r'"[^"]*?([^/"]+)", line (\d+)', l)
        cur = (ms[-1][0], int(ms[-1][1]))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        cnt[cur] += 1; n += 1
print("total", n, "instructions =", n * 16, "bytes")
for k, v in sorted(cnt.items(), key=lambda kv: kv[0][1] if kv[0] else 0):
    if v >= 12: print(k, v)
