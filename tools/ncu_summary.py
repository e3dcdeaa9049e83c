#!/usr/bin/env python
"""Summarises ncu outputs into small text files for profiles/ (run here; no GPU needed).

    python tools/ncu_summary.py rep  gpurun_out/x.ncu-rep  profiles/x.summary.txt
    python tools/ncu_summary.py list gpurun_out/launches.csv profiles/x.launches.tsv
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def rep(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, raw page of {path}\n")
        for r in data:
            f.write(f"\n== {r[hdr.index('Kernel Name')]}  (ID {r[0]})\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"{k:95s} {r[i]:>18s} {units[i]}\n")
            for i, h in enumerate(hdr):
                if ("pipe_tensor" in h or "l1tex__t_sector_hit_rate" in h or "lts__t_bytes.sum" == h or "l1tex__t_bytes.sum" == h
                        or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")) and h not in KEYS \
                        and r[i] not in ("", "0", "0.000000"):
                    f.write(f"{h:95s} {r[i]:>18s} {units[i]}\n")
    print(open(out).read())


def lst(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        name, ns = r[4], float(r[14])
        if r[13] == "us":
            ns *= 1e3
        elif r[13] == "ms":
            ns *= 1e6
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; {len(rows)} launches, total {total / 1e6:.3f} ms "
                f"(cold-cache, serialised: compare SHARES)\nkernel\tlaunches\ttotal_us\tshare\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name}\t{n}\t{ns / 1e3:.1f}\t{ns / total:.4f}\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    {"rep": rep, "list": lst}[sys.argv[1]](sys.argv[2], sys.argv[3])
