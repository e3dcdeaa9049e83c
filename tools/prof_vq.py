#!/usr/bin/env python
"""One quantizer configuration a few times (for an ncu capture of vq_tc_kernel): tools/prof_vq.py [D] [K] [N]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import torch
from vqvae.layers import Quantizer

D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1          # batch: the planes of one latent vector are N / B * 4 bytes apart
g = torch.Generator().manual_seed(1)
q = Quantizer(K, D, 0.1)
q.embed.copy_(torch.randn(K, D, generator=g)); q.first_pass.fill_(0)
q = q.cuda().eval()
x = torch.randn(B, D, N // B // 4096, 64, 64, generator=g).cuda()
with torch.no_grad():
    for _ in range(3):
        q(x)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        q(x)
    t1.record(); torch.cuda.synchronize()
print(f"D={D} K={K} N={N} B={B}: {t0.elapsed_time(t1) / 10:.3f} ms per call")
