#!/usr/bin/env python
"""Run-to-run reproducibility of the bf16 forward at batch B (eager and CUDA-graph replay)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = bench.build_model("full").to(dev)
x = torch.cat([bench.synthetic_volume((1, 1, 512, 512, 128), 42 + i) for i in range(B)]).to(dev)
def snap(r):
    return r[0].clone(), [i.clone() for i in r[1][2]]
def cmp(tag, a, b):
    bad = float(((a[0] - b[0]).abs() > 1e-3 + 1e-3 * b[0].abs()).float().mean())
    print(tag, f"voxels off {bad:.3e}  max abs {float((a[0] - b[0]).abs().max()):.3e}  idx mismatch",
          [f"{float((i != j).float().mean()):.2e}" for i, j in zip(a[1], b[1])], flush=True)
with torch.no_grad():
    r = [snap(model(x)) for _ in range(3)]
    cmp("eager run 0 vs 1", r[0], r[1]); cmp("eager run 0 vs 2", r[0], r[2])
    model.enable_cuda_graphs()
    g = [snap(model(x)) for _ in range(4)]
    cmp("graph run 1 vs 2", g[1], g[2]); cmp("graph run 1 vs 3", g[1], g[3]); cmp("eager vs graph", r[0], g[3])
