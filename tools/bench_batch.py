#!/usr/bin/env python
"""Forward throughput per volume for batch 1 / 2 / 4 (independent volumes stacked along B), CUDA-graph replay."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
dev = torch.device("cuda", 0)
for B in (1, 2, 4):
    model = bench.build_model("full").to(dev)
    x = torch.cat([bench.synthetic_volume((1, 1, 512, 512, 128), 42 + i) for i in range(B)]).to(dev)
    with torch.no_grad():
        model.enable_cuda_graphs()
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = model(x)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"batch {B}: {ms:.3f} ms per step, {ms / B:.3f} ms per volume, {B / ms * 1e3:.1f} volumes/s, mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del model, x, out
    torch.cuda.empty_cache()
