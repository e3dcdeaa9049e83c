#!/usr/bin/env python
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from vqvae import _ops
from vqvae.layers import Quantizer
o = _ops.default()
N, D, K = 1 << 20, 128, 512
g = torch.Generator().manual_seed(1)
q = Quantizer(K, D, 0.1)
q.embed.copy_(torch.randn(K, D, generator=g)); q.embed_avg.copy_(q.embed); q.first_pass.fill_(0); q.cluster_size.fill_(1.0)
q = q.cuda().train(True)
x = torch.randn(1, D, N // 4096, 64, 64, device="cuda")
with torch.no_grad():
    for it in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); _, _, idx = q(x); e1.record(); torch.cuda.synchronize()
        ws = [w for w in o._ws.values()][0]
        dbg = ws[:64].cpu().numpy().view(np.uint32)[4:9].tolist()
        n = q.embed.norm(dim=1)
        used = torch.bincount(idx.flatten(), minlength=K)
        print(f"step {it}: {e0.elapsed_time(e1):.3f} ms  dbg[single,rerank,none,ovf,sumnc]={dbg}  |e| min/mean/max = {n.min():.3f}/{n.mean():.3f}/{n.max():.3f}"
              f"  cluster_size min/max {q.cluster_size.min():.1f}/{q.cluster_size.max():.1f} used codes {(used > 0).sum().item()} max count {used.max().item()}", flush=True)
