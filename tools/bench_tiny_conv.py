#!/usr/bin/env python
"""GPU time per call of the tensor-core convolution on the tiny top-level shapes (back-to-back launches, CUDA events)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import torch
from vqvae import _ops
o = _ops.default()
dev = "cuda"
for (cin, cout, k, shp, B) in [(128, 32, 1, (8, 8, 2), 1), (256, 128, 1, (8, 8, 2), 1), (128, 128, 3, (8, 8, 2), 1), (64, 64, 3, (16, 16, 4), 1),
                               (128, 32, 1, (8, 8, 2), 8), (128, 128, 3, (8, 8, 2), 8), (32, 16, 1, (8, 8, 2), 1)]:
    x = torch.randn(B, cin, *shp, device=dev)
    w = torch.randn(cout, cin, k, k, k, device=dev) * 0.05
    with torch.no_grad():
        for _ in range(5):
            y = o.conv3d(x, w, pad=(k - 1) // 2, circular=k > 1)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                y = o.conv3d(x, w, pad=(k - 1) // 2, circular=k > 1)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
    print(f"{cin}->{cout} k{k} @{shp} B{B}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per conv (graph replay)", flush=True)
