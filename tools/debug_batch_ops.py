#!/usr/bin/env python
"""Batch consistency of the bf16 tensor-core paths at the Full model's shapes: op(cat[x0, x1]) vs cat[op(x0), op(x1)]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from vqvae import _ops, layers as L
o = _ops.default()
DEV = "cuda"

def rand_stack(C, n, seed):
    torch.manual_seed(seed)
    blocks = [L.PreActFixupResBlock(C, C, "same") for _ in range(n)]
    with torch.no_grad():
        for b in blocks:
            b.initialize_weights(n)
            for p in b.parameters():
                p.add_(torch.randn(p.shape) * 0.05)
    return L.BlockSequence(*blocks).to(DEV).eval()

def check(tag, fn, xs):
    with torch.no_grad():
        o.profile = []
        yb = fn(torch.cat(xs))
        names = sorted(set(e[0] for e in o.profile)); o.profile = None
        ys = torch.cat([fn(x) for x in xs])
    d = (yb - ys).abs()
    per = [float(d[i].max()) for i in range(len(xs))]
    print(f"{tag:48s} {names} max|diff| per item {per}  scale {float(ys.abs().max()):.3f}", flush=True)

g = torch.Generator().manual_seed(0)
for C, n, shp in [(18, 3, (128, 128, 32)), (72, 3, (32, 32, 8)), (32, 3, (8, 8, 2)), (8, 3, (32, 32, 8)), (16, 3, (128, 128, 32)),
                  (64, 3, (32, 32, 8)), (32, 3, (64, 64, 16)), (16, 3, (16, 16, 4))]:
    seq = rand_stack(C, n, C + n)
    xs = [torch.randn(1, C, *shp, generator=g).to(DEV) for _ in range(2)]
    check(f"stack{n} {C}ch @{shp}", seq, xs)
for (cin, cout, mode, shp) in [(16, 32, "down", (128, 128, 32)), (64, 128, "down", (32, 32, 8)), (128, 256, "down", (16, 16, 4)), (256, 128, "up", (8, 8, 2)),
                               (72, 32, "up", (32, 32, 8)), (18, 8, "up", (128, 128, 32)), (32, 16, "up", (64, 64, 16)), (8, 16, "down", (256, 256, 64)),
                               (4, 8, "down", (512, 512, 128)), (8, 4, "up", (256, 256, 64)), (72, 8, "same", (32, 32, 8)), (18, 2, "same", (128, 128, 32))]:
    torch.manual_seed(cin + cout)
    blk = L.PreActFixupResBlock(cin, cout, mode)
    with torch.no_grad():
        blk.initialize_weights(4)
        for p in blk.parameters():
            p.add_(torch.randn(p.shape) * 0.05)
    blk = blk.to(DEV).eval()
    xs = [torch.randn(1, cin, *shp, generator=g).to(DEV) for _ in range(2)]
    check(f"{mode} {cin}->{cout} @{shp}", blk, xs)
