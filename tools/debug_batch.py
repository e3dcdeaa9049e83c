#!/usr/bin/env python
"""bf16 (product) mode: batch-2 forward vs the two single-volume forwards, and run-to-run reproducibility."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device("cuda", 0)
shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (512, 512, 128)
model = bench.build_model("full").to(dev)
xs = [bench.synthetic_volume((1, 1) + shape, 42 + i).to(dev) for i in range(2)]
def cmp(tag, a, b):
    dec_a, (_, _, idx_a) = a
    dec_b, (_, _, idx_b) = b
    bad = float(((dec_a - dec_b).abs() > 1e-3 + 1e-3 * dec_b.abs()).float().mean())
    print(tag, "decoded voxels off:", f"{bad:.3e}", "max abs", float((dec_a - dec_b).abs().max()),
          "idx mismatch per level:", [f"{float((i != j).float().mean()):.2e}" for i, j in zip(idx_a, idx_b)], flush=True)
with torch.no_grad():
    s0 = model(xs[0]); s0 = (s0[0].clone(), (None, None, [i.clone() for i in s0[1][2]]))
    s0b = model(xs[0]); s0b = (s0b[0].clone(), (None, None, [i.clone() for i in s0b[1][2]]))
    cmp("run-to-run (single volume)", s0, s0b)
    s1 = model(xs[1]); s1 = (s1[0].clone(), (None, None, [i.clone() for i in s1[1][2]]))
    b = model(torch.cat(xs))
    for i, s in enumerate((s0, s1)):
        cmp(f"batch-2 item {i} vs single", (b[0][i:i + 1], (None, None, [t[i:i + 1] for t in b[1][2]])), s)
