#!/usr/bin/env python
"""Where a training step's time goes: wall time vs the sum of per-C-call CUDA-event times, grouped by op."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import torch
import bench
from vqvae import _ops
from vqvae.parallel import training_step
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="downscaled_256x256x128")
a = ap.parse_args()
kind, shape = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
m = bench.build_model(kind).to(dev).train()
x = bench.synthetic_volume(shape, 42).to(dev)
opt = m.configure_optimizers()
batch = (x, [shape[4]])
o = _ops.default()
for _ in range(2):
    training_step(m, opt, batch)
torch.cuda.synchronize()
o.profile = []
l0 = o.launches
t0 = time.perf_counter()
training_step(m, opt, batch)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
prof, o.profile = o.profile, None
groups = {}
for name, tag, nbytes, flops, e0, e1 in prof:
    g = groups.setdefault(name, [0, 0.0])
    g[0] += 1; g[1] += e0.elapsed_time(e1)
tot = sum(g[1] for g in groups.values())
print(f"{a.workload}: wall {wall * 1e3:.1f} ms (profiling on), C calls {len(prof)}, kernels {o.launches - l0}, sum of event times {tot:.1f} ms")
for k, g in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} calls {g[0]:6d}  {g[1]:9.2f} ms  {g[1] / g[0] * 1e3:8.1f} us/call")
tags = {}
for name, tag, nbytes, flops, e0, e1 in prof:
    if name in ("conv3d_backward", "conv3d", "conv3d_tc", "conv1x1_backward", "preact_same_backward", "preact_block"):
        g = tags.setdefault((name, tag), [0, 0.0])
        g[0] += 1; g[1] += e0.elapsed_time(e1)
for k, g in sorted(tags.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"  {k[0]:16s} {k[1]:32s} calls {g[0]:5d}  {g[1]:8.2f} ms  {g[1] / g[0] * 1e3:8.1f} us/call")
