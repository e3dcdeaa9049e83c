#!/usr/bin/env python
"""One quantizer forward (eval) for ncu: python tools/prof_vq.py N D K"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import torch
from vqvae.layers import Quantizer
N, D, K = (int(a) for a in sys.argv[1:4])
dev = torch.device("cuda", 0)
q = Quantizer(K, D, 0.1); q.first_pass.fill_(0); q = q.to(dev).eval()
x = torch.randn(1, D, N // 4096, 64, 64, device=dev)
with torch.no_grad():
    for _ in range(3):
        q(x)
    torch.cuda.synchronize()
print("ok")
