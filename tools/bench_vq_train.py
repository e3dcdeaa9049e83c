#!/usr/bin/env python
"""Training-mode (EMA statistics on) vs eval timing of the quantizer at a few sweep points."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_quantizer import run
for (N, D, K) in [(1 << 20, 32, 512), (1 << 20, 64, 512), (1 << 20, 128, 512), (1 << 20, 128, 4096)]:
    te, tt = run(N, D, K, False), run(N, D, K, True)
    print(f"N={N} D={D} K={K}: eval {te * 1e3:.3f} ms, train {tt * 1e3:.3f} ms", flush=True)
