#!/usr/bin/env python
"""Runs ONE named op of the hot path a few times on cuda:0 (short command lines for ncu).

    python tools/prof_case.py stack18 [--reps 3] [--precision bf16|fp32]

Cases are the layers of the Full model at 512x512x128 (SURVEY.md 8a sizes).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from vqvae import _ops  # noqa: E402
from vqvae.layers import BlockSequence, PreActFixupResBlock, Conv3d, Quantizer  # noqa: E402

CASES = {
    # name: (kind, channels_in, channels_out, mode, n_blocks, spatial)
    "stack18": ("stack", 18, 18, "same", 4, (128, 128, 32)),
    "stack72": ("stack", 72, 72, "same", 4, (32, 32, 8)),
    "stack32": ("stack", 32, 32, "same", 8, (8, 8, 2)),
    "stack8": ("stack", 8, 8, "same", 4, (32, 32, 8)),
    "stack2": ("stack", 2, 2, "same", 4, (128, 128, 32)),
    "stack4_512": ("stack", 4, 4, "same", 3, (512, 512, 128)),
    "stack8_256": ("stack", 8, 8, "same", 3, (256, 256, 64)),
    "up8_256": ("block", 8, 4, "up", 1, (256, 256, 64)),
    "down4_512": ("block", 4, 8, "down", 1, (512, 512, 128)),
    "down16_128": ("block", 16, 32, "down", 1, (128, 128, 32)),
    "down8_256": ("block", 8, 16, "down", 1, (256, 256, 64)),
    "up18_128": ("block", 18, 8, "up", 1, (128, 128, 32)),
    "up32_64": ("block", 32, 16, "up", 1, (64, 64, 16)),
    "out4_512": ("conv1", 4, 1, None, 1, (512, 512, 128)),
    "in1_512": ("conv1", 1, 4, None, 1, (512, 512, 128)),
    "vq0": ("vq", 2, 128, None, 1, (128, 128, 32)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case", choices=sorted(CASES))
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=1, help="volumes stacked along B (bench.py default: 8)")
    a = ap.parse_args()
    kind, cin, cout, mode, n, sp = CASES[a.case]
    dev = torch.device("cuda", 0)
    o = _ops.default()
    o.precision = a.precision
    torch.manual_seed(0)
    if kind == "stack":
        m = BlockSequence(*(PreActFixupResBlock(cin, cout, mode) for _ in range(n)))
    elif kind == "block":
        m = PreActFixupResBlock(cin, cout, mode)
    elif kind == "conv1":
        m = Conv3d(cin, cout, kernel_size=1)
    else:
        m = Quantizer(cout, cin, 0.1)
        m.first_pass.fill_(0)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    m = m.to(dev).eval()
    x = torch.rand(a.batch, cin, *sp, device=dev) * 2 - 0.5
    with torch.no_grad():
        for _ in range(a.reps):
            y = m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            y = m(x)
        e1.record()
        torch.cuda.synchronize()
    print(f"{a.case}: {e0.elapsed_time(e1) / a.reps * 1e3:.1f} us per call ({n} block(s), batch {a.batch})")


if __name__ == "__main__":
    main()
