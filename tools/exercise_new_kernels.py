#!/usr/bin/env python
"""Small invocations (tail tiles, padded tiles, partial boxes) of the kernels added on the last day of round 2 -- a quick crash / launch-error
check on a GPU box (the parity tests cover their results; compute-sanitizer is not available on this pool):
the hi/lo + group-store quantizer (D 32) and the residual-norm margin path (D 64), the k4 s2 / k2 s2 tensor-core gathers, the 9-channel
tiled convolution, the row-sliding weight gradient, the templated conv1x1 backward, the median radix select."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import torch
from vqvae import _ops
from vqvae.layers import Quantizer, PreActFixupResBlock

dev = "cuda"
o = _ops.default()
g = torch.Generator().manual_seed(0)
for D, K, N in ((32, 300, 33000), (32, 1100, 40000), (64, 512, 33000)):          # tail tiles, padded codebook tiles, streamed ring
    q = Quantizer(K, D, 0.1)
    q.embed.copy_(torch.randn(K, D, generator=g)); q.first_pass.fill_(0)
    q = q.to(dev).eval()
    x = torch.randn(1, D, N, 1, 1, generator=g).to(dev)
    with torch.no_grad():
        _, _, idx = q(x)
    print("quantizer", D, K, N, int(idx.sum()))
o.precision = "bf16"
for (c1, co, k, s, circ, shape) in ((16, 16, 4, 2, True, (1, 16, 16, 8)), (16, 16, 4, 2, False, (2, 6, 10, 6)), (16, 32, 2, 2, False, (1, 8, 8, 8))):
    x = torch.randn(shape[0], c1, *shape[1:], generator=g).to(dev)
    w = (torch.randn(co, c1, k, k, k, generator=g) * 0.1).to(dev)
    y = o.conv3d(x, w, stride=s, pad=1 if k > 2 else 0, circular=circ)
    print("conv3d_tc", k, s, float(y.abs().sum()))
o.precision = "fp32"
for (cin, cout, shape) in ((18, 18, (1, 18, 9, 10, 40)), (8, 8, (2, 8, 5, 6, 33)), (2, 2, (1, 2, 6, 5, 32))):
    blk = PreActFixupResBlock(cin, cout, "same").to(dev)
    with torch.no_grad():
        for p_ in blk.parameters():
            p_.add_(torch.randn(p_.shape, generator=g).to(dev) * 0.1)
    x = torch.randn(shape, generator=g).to(dev).requires_grad_(True)
    y = blk(x)
    y.backward(torch.ones_like(y))
    print("block fwd+bwd", cin, float(x.grad.abs().sum()))
dec = (torch.rand(2, 1, 12, 10, 9, generator=g) * 3 - 0.5).to(dev)
xx = (torch.rand(2, 1, 12, 10, 9, generator=g) * 4.5 - 0.5).to(dev)
log = o.huber_metrics(dec, xx, torch.tensor([9, 5], dtype=torch.int32, device=dev), None)
print("medians", float(log["loc_median"]), float(log["recon_loss_median"]))
torch.cuda.synchronize()
print("ok")
