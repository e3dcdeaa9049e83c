#!/usr/bin/env python
"""One PreActFixupResBlock forward + backward (training path: composed differentiable ops) for ncu launch lists:
    python tools/prof_bwd.py CIN COUT MODE H W Z [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import torch
from vqvae.layers import PreActFixupResBlock
cin, cout, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
H, W, Z = (int(a) for a in sys.argv[4:7])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 2
torch.manual_seed(0)
blk = PreActFixupResBlock(cin, cout, mode).cuda()
x = torch.randn(1, cin, H, W, Z, device="cuda", requires_grad=True)
for _ in range(reps):
    y = blk(x)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok")
