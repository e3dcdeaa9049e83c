#!/usr/bin/env python
"""Training-step time (forward in training mode + backward + gradient all-reduce + fused Adam) of the published configs:

    python tools/bench_train.py --workload downscaled_256x256x128 [--steps 3]        (BASELINE.json configs[1])
    python tools/bench_train.py --workload full_512x512x128                          (configs[2], per GPU)

The backward runs on the shape-generic fp32 kernels of csrc/backward_kernels.cu (correctness-first); this tool
exists to put an honest number next to the forward benchmark, not as the headline metric."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import torch
import bench
from vqvae.parallel import training_step


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="downscaled_256x256x128", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1, help="volumes per step (the reference's --batch-size; it trained with 1 on 24 GB cards)")
    ap.add_argument("--graph", action="store_true", help="capture the whole step in one CUDA graph (vqvae.parallel.GraphedTrainingStep)")
    a = ap.parse_args()
    kind, shape = bench.WORKLOADS[a.workload]
    dev = torch.device("cuda", 0)
    m = bench.build_model(kind).to(dev).train()
    for q in m.encoder.quantize:
        q.first_pass.fill_(1)
    x = torch.cat([bench.synthetic_volume(shape, 42 + i) for i in range(a.batch)]).to(dev)
    opt = m.configure_optimizers()
    batch = (x, [shape[4]] * a.batch)
    torch.cuda.reset_peak_memory_stats()
    loss = training_step(m, opt, batch)          # warm-up (first-pass codebook init)
    torch.cuda.synchronize()
    step_fn = (lambda b: training_step(m, opt, b))
    if a.graph:
        from vqvae.parallel import GraphedTrainingStep
        step_fn = GraphedTrainingStep(m, opt, batch, warmup=2)
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step_fn(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"workload": a.workload, "train_step_ms": ms, "batch": a.batch, "volumes_per_s": a.batch * 1e3 / ms, "loss": float(loss),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "steps": a.steps, "cuda_graph": bool(a.graph),
                      "note": "fwd composed (tensor-core convs where GEMM-shaped) + generic fp32 backward + fused Adam"}))


if __name__ == "__main__":
    main()
