#!/usr/bin/env python
"""Debug aid for the tensor-core quantizer: index mismatches against the exact SIMT scan, and (library built with
-DVQ3D_VQT_DEBUG) how many vectors took the single-candidate / re-rank / full-scan paths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from vqvae import _ops
from vqvae.layers import Quantizer

o = _ops.default()
for (N, D, K) in [(65536, 32, 512), (100000, 32, 512), (1 << 20, 32, 512), (65536, 64, 1024), (40000, 128, 2048), (1 << 20, 32, 4096)]:
    g = torch.Generator().manual_seed(1)
    q = Quantizer(K, D, 0.1)
    q.embed.copy_(torch.randn(K, D, generator=g)); q.first_pass.fill_(0)
    q = q.cuda().eval()
    x = torch.randn(1, D, N, 1, 1, generator=g).cuda()
    with torch.no_grad():
        o.vq_tensor_cores = True
        _, _, i1 = q(x)
        torch.cuda.synchronize()
        ws = [w for w in o._ws.values()][0]
        dbg = ws[:64].cpu().numpy().view(np.uint32)[4:9].tolist()
        o.vq_tensor_cores = False
        _, _, i0 = q(x)
    bad = (i1 != i0).flatten().nonzero().flatten().cpu().numpy()
    print(f"N={N} D={D} K={K}: mismatches {bad.size} {bad[:4]} (super-tiles {sorted(set((bad // 256).tolist()))[:8]})  "
          f"[single, rerank, none, overflow, sum nc of rerank] = {dbg}", flush=True)
