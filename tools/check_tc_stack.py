#!/usr/bin/env python
"""GPU check + timing (CUDA-graph replay, warm clocks) of the tensor-core stack kernel vs the fp32 SIMT stack."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import torch
from vqvae import _ops
from vqvae.layers import BlockSequence, PreActFixupResBlock

o = _ops.default()
dev = torch.device("cuda", 0)
cases = [(18, 1, (16, 16, 32)), (18, 4, (128, 128, 32)), (72, 4, (32, 32, 8)), (32, 8, (8, 8, 2)), (8, 8, (32, 32, 8)),
         (16, 3, (128, 128, 32)), (8, 3, (256, 256, 64)), (18, 50, (128, 128, 32)), (72, 50, (32, 32, 8)), (32, 50, (8, 8, 2)),
         (8, 50, (32, 32, 8))]
sel = [a for a in sys.argv[1:] if not a.startswith("-")]
if sel:
    cases = [cases[int(a)] for a in sel]
precs = ("bf16",) if "--tc-only" in sys.argv else ("fp32", "bf16")
# warm the clocks
a = torch.randn(4096, 4096, device=dev)
t0 = time.time()
while time.time() - t0 < 0.3:
    (a @ a).sum().item()
for C, n, sp in cases:
    torch.manual_seed(C + n)
    blocks = [PreActFixupResBlock(C, C, "same") for _ in range(n)]
    with torch.no_grad():
        for b in blocks:
            b.initialize_weights(num_layers=max(n, 2))
            for p in b.parameters():
                p.add_(torch.randn(p.shape) * 0.05)
    seq = BlockSequence(*blocks).to(dev).eval()
    x = torch.randn(1, C, *sp, device=dev)
    res = {}
    with torch.no_grad():
        for prec in precs:
            o.precision = prec
            try:
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    y = seq(x)
                torch.cuda.synchronize()
                if "--no-graph" in sys.argv:
                    res[prec] = (y.clone(), 0.0)
                    continue
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    y = seq(x)
                for _ in range(3):
                    g.replay()
                torch.cuda.synchronize()
                reps = 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    g.replay()
                e1.record(); torch.cuda.synchronize()
                res[prec] = (y.clone(), e0.elapsed_time(e1) / reps * 1e3)
            except Exception as ex:
                print(f"C={C} n={n} {sp} {prec}: FAILED {ex}")
                res[prec] = None
    if res.get("fp32") and res.get("bf16"):
        ref, got = res["fp32"][0], res["bf16"][0]
        br = ref - x
        print(f"C={C} n={n} {sp}: fp32 {res['fp32'][1]:.1f} us ({res['fp32'][1] / n:.1f}/block)  tc {res['bf16'][1]:.1f} us ({res['bf16'][1] / n:.1f}/block)  "
              f"max err {float((got - ref).abs().max()):.3e} (branch max {float(br.abs().max()):.3e})  "
              f"mean err {float((got - ref).abs().mean()):.3e} (branch mean {float(br.abs().mean()):.3e})", flush=True)
    elif res.get("bf16"):
        print(f"C={C} n={n} {sp}: tc {res['bf16'][1]:.1f} us ({res['bf16'][1] / n:.1f}/block)", flush=True)
