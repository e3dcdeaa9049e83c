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
for (N, D, K) in [(1 << 20, 32, 512), (1 << 20, 128, 512), (1 << 20, 32, 4096), (1 << 20, 128, 4096)]:
    g = torch.Generator().manual_seed(1)
    q = Quantizer(K, D, 0.1)
    q.embed.copy_(torch.randn(K, D, generator=g)); q.first_pass.fill_(0)
    q = q.cuda().eval()
    x = torch.randn(1, D, N, 1, 1, generator=g).cuda()
    with torch.no_grad():
        o.vq_tensor_cores = True
        _, _, i1 = q(x)
        torch.cuda.synchronize()
        ws = max(o._ws.values(), key=lambda w: w.numel())
        raw = ws[:256].cpu().numpy()
        dbg = raw.view(np.uint32)[4:9].tolist()
        clk = raw[16:16 + 16 * 8].view(np.uint64)          # slots 4..11 (64-bit) after the five 32-bit counters
        tot = float(clk[11]) or 1.0
        names = {4: "issuer:a_full", 5: "issuer:d_empty", 6: "sweep:r_empty", 7: "sweep:a_full", 8: "sweep:d_full", 9: "le:a_empty", 10: "le:r_full",
                 12: "le:merge", 13: "le:rerank", 14: "le:gather", 15: "le:stage"}
        nw = {4: 1, 5: 1, 6: 16, 7: 16, 8: 16, 9: 8, 10: 8, 12: 8, 13: 8, 14: 8, 15: 8}
        waits = {names[k]: round(float(clk[k]) / nw[k] / tot, 3) for k in names}
        o.vq_tensor_cores = False
        _, _, i0 = q(x)
    bad = (i1 != i0).flatten().nonzero().flatten().cpu().numpy()
    print(f"N={N} D={D} K={K}: mismatches {bad.size} {bad[:4]} (super-tiles {sorted(set((bad // 256).tolist()))[:8]})  "
          f"[single, rerank, none, overflow, sum nc of rerank] = {dbg}  wait fraction of the kernel per warp of the role: {waits}", flush=True)
