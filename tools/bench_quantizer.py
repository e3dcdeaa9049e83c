#!/usr/bin/env python
"""Quantizer sweep (BASELINE.json configs[4]): Gcodes/s of Quantizer.forward (eval and training mode) for
N latent vectors x codebook K x embedding dim D, against the HBM / tensor rooflines of SURVEY.md 8d.

    python tools/bench_quantizer.py [--quick] [--out profiles/x.tsv]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    sys.path.insert(0, p)
import torch
from vqvae import _ops
from vqvae.layers import Quantizer


def run(N, D, K, training, reps=5):
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(N % 1000 + D + K)
    q = Quantizer(K, D, 0.1)
    q.embed.copy_(torch.randn(K, D, generator=g)); q.embed_avg.copy_(q.embed); q.first_pass.fill_(0); q.cluster_size.fill_(1.0)
    q = q.to(dev).train(training)
    x = torch.randn(1, D, N // 4096, 64, 64, device=dev)
    with torch.no_grad():
        for _ in range(2):
            q(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            q(x)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--simt", action="store_true", help="force the exact SIMT scan (no tensor-core candidate pass)")
    a = ap.parse_args()
    if a.simt:
        _ops.default().vq_tensor_cores = False
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    Ns = [1 << 20] if a.quick else [1 << 20, 1 << 22, 1 << 24, 1 << 26]      # BASELINE.json configs[4]: 1 M ... 64 M
    rows = []
    for N in Ns:
        for K in (512, 1024, 4096):
            for D in (32, 64, 128):
                if N * D * 4 * 2 > 80e9:            # x + quantized (fp32) must fit comfortably: 64 M x D128 = 69 GB
                    continue
                for training in (False, True):
                    if training and (a.quick or N > (1 << 22)):
                        continue
                    t = run(N, D, K, training, reps=2 if N >= (1 << 26) else 3 if N >= (1 << 24) else 5)
                    gc = N / t / 1e9
                    hbm = pk["hbm_gbs"] * 1e9 / (8 * D + 8) / 1e9
                    tens = pk["bf16_tflops_sustained"] * 1e12 / (2.0 * K * D) / 1e9
                    roof = min(hbm, tens)
                    rows.append((N, K, D, "train" if training else "eval", t * 1e3, gc, hbm, tens, gc / roof))
                    print(f"N={N:>9d} K={K:4d} D={D:3d} {'train' if training else 'eval ':5s} {t * 1e3:9.3f} ms  {gc:7.3f} Gcodes/s  "
                          f"roofline min(hbm {hbm:.2f}, tensor {tens:.2f}) -> {100 * gc / roof:5.1f}%", flush=True)
    if a.out:
        with open(a.out, "w") as f:
            f.write("N\tK\tD\tmode\tms\tGcodes_per_s\thbm_roof_Gcodes\ttensor_roof_Gcodes\tfrac_of_roofline\n")
            for r in rows:
                f.write("\t".join(str(round(v, 4)) if isinstance(v, float) else str(v) for v in r) + "\n")


if __name__ == "__main__":
    main()
