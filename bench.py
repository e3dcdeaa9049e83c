#!/usr/bin/env python
"""bench.py -- encode -> quantize -> decode throughput of the Full 3-level 3D VQ-VAE-2 on
synthetic 512x512x128 volumes (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference ...                     (the oracle on the host cores)

A "step" = one forward pass (Encoder2 -> 3 quantizers -> Decoder, model.py:79-83) over one batch of
--batch independent volumes per GPU (default 8, stacked along B); volumes are independent, so ranks
shard them with no data-path collective (weak scaling).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # name: (config fn name, volume shape)
    "full_512x512x128": ("full", (1, 1, 512, 512, 128)),
    "downscaled_256x256x128": ("down", (1, 1, 256, 256, 128)),
    "downscaled_128x128x64": ("down", (1, 1, 128, 128, 64)),
}
CPU_SAMPLE_SHAPE = (1, 1, 128, 128, 64)     # bounded sample for the CPU arm: 1/32 of a 512x512x128 volume


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def build_model(kind, seed=42):
    """Reference constructors' RNG stream under seed 42 + Fixup init, then every parameter
    += N(0, 0.02) (at pure Fixup init branch_conv3 = 0 and all branches are dead)."""
    from vqvae.model import VQVAE, downscaled_config_args, full_config_args
    torch.manual_seed(seed)
    m = VQVAE(full_config_args() if kind == "full" else downscaled_config_args())
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.02)
        for q in m.encoder.quantize:
            q.first_pass.fill_(0)
    return m.eval()


def volume_seed(rank, i=0, batch=1):
    """Volumes are independent (train_vqvae_3d.job:76): rank r works on its own synthetic volumes."""
    return 42 + rank * batch + i


def aggregate(world, steps, elapsed_ms_max, batch=1):
    """Whole-job volumes/s: every rank did `steps` batches of `batch` volumes, the job took the slowest rank's time."""
    return world * steps * batch / (elapsed_ms_max * 1e-3)


def synthetic_volume(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * 4.5 - 0.5        # reference value range, SURVEY.md 8d


# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.err = index, [], set(), False, None, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = get(h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.01)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), **({"error": self.err} if self.err else {})}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


_ORIGINAL_AFFINITY = None


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to its GPU before any pinned host buffer is allocated (first touch
    puts the pages on that NUMA node): with 8 ranks the end-to-end leg moves ~200 GB/s through host memory, and remote-node
    staging buffers were what capped it.  Best effort: returns the CPU count bound to, or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and w * 64 + b < ncpu}
        original = set(os.sched_getaffinity(0))
        allowed = cpus & original
        if allowed:
            global _ORIGINAL_AFFINITY
            _ORIGINAL_AFFINITY = original
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def quantizer_point(dev, pk, N=1 << 20, D=32, K=512, reps=10):
    """Quantizer.forward (eval) on N latent vectors presented as (1, D, N/4096, 64, 64): Gcodes/s and the
    fraction of min(HBM, tensor) roofline (8D+8 bytes, 2KD flops per vector)."""
    from vqvae.layers import Quantizer
    g = torch.Generator().manual_seed(7)
    q = Quantizer(K, D, 0.1)
    q.embed.copy_(torch.randn(K, D, generator=g)); q.first_pass.fill_(0)
    q = q.to(dev).eval()
    x = torch.randn(1, D, N // 4096, 64, 64, generator=g).to(dev)
    with torch.no_grad():
        for _ in range(3):
            q(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            q(x)
        e1.record()
        torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps * 1e-3
    gc = N / t / 1e9
    hbm = pk["hbm_gbs"] / (8 * D + 8)
    tens = pk.get("bf16_tflops_sustained", 1400.0) * 1e3 / (2.0 * K * D)
    # third ceiling of this design: every score has to leave TMEM (N*K*4 bytes) at 64 B/clk per SM (guide figure, 148 SMs)
    tmem = 148 * 64 * 1.965 / (4.0 * K)
    return {"value": gc, "unit": "Gcodes/s", "config": {"N": N, "D": D, "K": K, "mode": "eval", "path": "tcgen05 bf16 candidate pass + exact fp32 re-rank"},
            "ms": t * 1e3, "roofline": {"bound": "hbm" if hbm < tens else "tensor", "peak": min(hbm, tens), "unit": "Gcodes/s", "frac": gc / min(hbm, tens),
                                        "tmem_read_ceiling": tmem, "frac_of_tmem_read_ceiling": gc / tmem},
            "sweep": "profiles/r01_final_quantizer_sweep_tc_v2.tsv (tools/bench_quantizer.py)"}


def train_point(dev, steps=3):
    from vqvae.parallel import GraphedTrainingStep, training_step
    kind, shape = WORKLOADS["downscaled_256x256x128"]
    torch.cuda.empty_cache()
    m = build_model(kind).to(dev).train()
    for q in m.encoder.quantize:
        q.first_pass.fill_(1)
    x = synthetic_volume(shape, 42).to(dev)
    opt = m.configure_optimizers()
    batch = (x, [shape[4]])
    training_step(m, opt, batch)                      # data-dependent codebook init (layers.py:665-683)
    step = GraphedTrainingStep(m, opt, batch, warmup=2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"workload": "downscaled_256x256x128 (BASELINE.json configs[1])", "ms_per_step": ms, "volumes_per_s": 1e3 / ms, "batch": 1,
           "loss": float(loss), "cuda_graph": True, "steps": steps,
           "what": "forward (training mode, EMA codebook update) + backward + fused Adam(amsgrad); fp32 master weights, bf16 tcgen05 convolutions where GEMM-shaped"}
    del step, m, opt
    torch.cuda.empty_cache()
    return out


def ncu_traffic(kernel_tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    --set full capture (profiles/ncu_traffic.json, written by hand from profiles/*.summary.txt)."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None
    for key, val in table.items():
        if key in kernel_tag:
            return val
    return None


# ---------------------------------------------------------------------------------------
def cpu_arm(kind, steps, warmup):
    """The reference's CPU path (oracle port, ATen fp32, all host cores) on a bounded sample."""
    from oracle import vqvae_oracle as O
    if _ORIGINAL_AFFINITY is not None:          # the CPU arm uses every host core, not only the GPU-local ones
        os.sched_setaffinity(0, _ORIGINAL_AFFINITY)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = build_model(kind)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    cfg = O.FULL if kind == "full" else O.DOWNSCALED
    x = synthetic_volume(CPU_SAMPLE_SHAPE, 42)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.vqvae_forward(sd, cfg, x)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    kind, shape = WORKLOADS[args.workload]
    frac = (shape[2] * shape[3] * shape[4]) / (CPU_SAMPLE_SHAPE[2] * CPU_SAMPLE_SHAPE[3] * CPU_SAMPLE_SHAPE[4])
    steps = min(args.steps, 5)
    times, cores = cpu_arm(kind, steps, min(args.warmup, 1))
    total = sum(times)
    value = steps / (total * frac)
    sample = (f"{kind} model forward on a {CPU_SAMPLE_SHAPE[2]}x{CPU_SAMPLE_SHAPE[3]}x{CPU_SAMPLE_SHAPE[4]} crop "
              f"(1/{frac:g} of the volume's voxels) per step; volumes/s = crops/s / {frac:g}")
    print(json.dumps({
        "impl": "reference", "metric": "volumes_per_s_encode_vq_decode", "value": value, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / steps * frac,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "volume": list(shape), "batch_per_gpu": args.batch},
        "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from vqvae import _ops
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa_node(physical_gpu_index(local))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    kind, shape = WORKLOADS[args.workload]
    steps, warmup = args.steps, max(args.warmup, 3)
    ops = _ops.default()

    model = build_model(kind).to(dev)
    # a step = one forward over a batch of `batch` independent volumes stacked along B (the tiny top-level layers are
    # launch/latency bound at one volume; stacking amortises them -- every kernel takes B as a size, nothing is skipped)
    batch = max(1, args.batch)
    bshape = (batch,) + tuple(shape[1:])
    x_host = torch.cat([synthetic_volume(shape, volume_seed(rank, i, batch)) for i in range(batch)]).pin_memory()
    x_dev = x_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # one eager profiled pass: per-C-call CUDA events (launching stream) -> kernel shares + dominant kernel
        model(x_dev)
        torch.cuda.synchronize()
        ops.profile = []
        l0 = ops.launches
        model(x_dev)
        torch.cuda.synchronize()
        launches_per_step = ops.launches - l0
        prof, ops.profile = ops.profile, None
        groups = {}
        for name, tag, nbytes, flops, e0, e1 in prof:
            g = groups.setdefault((name, tag), [0, 0.0, 0, 0])
            g[0] += 1; g[1] += e0.elapsed_time(e1); g[2] += nbytes; g[3] += flops
        by_kernel = {}
        for (name, tag), g in groups.items():
            k = by_kernel.setdefault(name, [0, 0.0])
            k[0] += g[0]; k[1] += g[1]
        prof_total = sum(g[1] for g in groups.values())

        # the measured path: the public API with CUDA graphs on
        model.enable_cuda_graphs()
        for _ in range(warmup):
            model(x_dev)
        barrier()
        sampler = ClockSampler(physical_gpu_index(local))
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            out = model(x_dev)
        ev1.record()
        barrier()
        elapsed_ms = ev0.elapsed_time(ev1)
        sampler.stop_flag = True
        sampler.join(timeout=2)

        # dominant kernel, timed live in isolation on the launching stream (L2-cold: the workload's
        # working set is far larger than the 126 MB L2 and a 256 MB scratch is rewritten between repeats)
        dom_key = max(groups, key=lambda k: groups[k][1])
        dom = groups[dom_key]

        # end to end through the public API with HOST buffers: every step copies its volume from pinned host
        # memory, runs the forward and copies the reconstruction + the three code-index tensors back to pinned host
        # memory.  Copies ride on their own streams (double-buffered input, staged outputs) so that the H2D of
        # volume i+1 and the D2H of result i-1 overlap the forward of volume i; the clock runs until the last
        # result is in host memory.
        dec_host = torch.empty(bshape, dtype=torch.float32).pin_memory()
        e2e_steps = max(3, min(steps, 10))
        main = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        x_bufs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
        ev_in = [torch.cuda.Event(), torch.cuda.Event()]
        ev_fwd_read = [torch.cuda.Event(), torch.cuda.Event()]     # forward has consumed x_bufs[i]
        ev_staged, ev_out = torch.cuda.Event(), torch.cuda.Event()
        dec_stage = idx_stage = idx_host = None
        for it in range(2 + e2e_steps):
            if it == 2:
                barrier()
                t0 = time.perf_counter()
            b = it & 1
            with torch.cuda.stream(s_in):
                if it >= 2:
                    s_in.wait_event(ev_fwd_read[b])
                x_bufs[b].copy_(x_host, non_blocking=True)
                ev_in[b].record(s_in)
            main.wait_event(ev_in[b])
            dec, (_, _, idxs) = model(x_bufs[b])
            ev_fwd_read[b].record(main)
            if dec_stage is None:
                dec_stage = torch.empty_like(dec)
                idx_stage = [torch.empty_like(i) for i in idxs]
                idx_host = [torch.empty(i.shape, dtype=i.dtype).pin_memory() for i in idxs]
            if it > 0:
                main.wait_event(ev_out)              # the previous result has left the staging buffers
            dec_stage.copy_(dec, non_blocking=True)
            for st, d in zip(idx_stage, idxs):
                st.copy_(d, non_blocking=True)
            ev_staged.record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_staged)
                dec_host.copy_(dec_stage, non_blocking=True)
                for h, d in zip(idx_host, idx_stage):
                    h.copy_(d, non_blocking=True)
                ev_out.record(s_out)
        torch.cuda.synchronize()
        barrier()
        e2e_s = time.perf_counter() - t0
        # the pipelined loop returned the same bits as a plain call on the same volume
        chk, (_, _, chk_idx) = model(x_dev)
        # (split-K layers reduce with fp32 atomics, so two runs agree to rounding, not bit for bit; a latent that sits on a
        # near-tie can then pick the other code, which changes the reconstruction around it -- allow a small fraction)
        bad = float(((chk.cpu() - dec_host).abs() > 1e-3 + 1e-3 * dec_host.abs()).float().mean())
        assert bad < 1e-2, f"e2e pipeline result mismatch ({bad:.2e} of the voxels)"
        assert all(float((a.cpu() != h).float().mean()) < 1e-2 for a, h in zip(chk_idx, idx_host)), "e2e pipeline index mismatch"
    h2d = x_host.numel() * 4
    d2h = dec_host.numel() * 4 + sum(h.numel() * 8 for h in idx_host)

    t = torch.tensor([elapsed_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_s = float(t[0]), float(t[1])

    if rank == 0:
        pk, pk_src = peaks()
        n_dom, t_dom, b_dom, f_dom = dom
        achieved = b_dom / (t_dom * 1e-3) / 1e9            # GB/s, algorithmic bytes / event time
        intensity = f_dom / max(b_dom, 1)
        line = {
            "metric": "volumes_per_s_encode_vq_decode", "value": aggregate(world, steps, elapsed_ms, batch), "unit": "volumes/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": elapsed_ms / steps,
            "ms_per_volume": elapsed_ms / steps / batch,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "volume": list(shape), "batch_per_gpu": batch,
                       "model": "3-level Full (train_vqvae_3d.job flags)" if kind == "full" else "2-level downscaled",
                       "weights": "reference ctor RNG seed 42 + Fixup init + N(0,0.02) perturbation",
                       "l2": f"inputs larger than L2 ({134 * batch} MB of volumes, GBs of activations per step)",
                       "cuda_graph": True, "parallelism": f"dp{world} (independent volumes, no collective)",
                       "e2e": "pinned H2D / forward / D2H on three streams, double-buffered",
                       "host_cpus_bound": numa_cpus},
            "e2e": {"value": world * e2e_steps * batch / e2e_s, "unit": "volumes/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / e2e_steps},
            "gpu_launches": launches_per_step * steps,
            "clocks": sampler.result(),
            "roofline": {"bound": "hbm" if intensity < 210 else "tensor", "achieved": achieved, "peak": pk["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": None, "peak_source": pk_src,
                         "kernel": f"{dom_key[0]} [{dom_key[1]}]", "launches_per_step": n_dom,
                         "avg_launch_us": 1e3 * t_dom / n_dom, "share_of_step": t_dom / prof_total,
                         "algorithmic_bytes_per_launch": b_dom / n_dom, "flop_per_byte": intensity},
            "kernel_shares": {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / prof_total, 4)}
                              for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1][1])},
            "top_ops": [{"op": f"{k[0]} [{k[1]}]", "n": g[0], "ms": round(g[1], 3),
                         "GBps": round(g[2] / max(g[1], 1e-9) / 1e6, 1), "TFLOPs": round(g[3] / max(g[1], 1e-9) / 1e9, 2)}
                        for k, g in sorted(groups.items(), key=lambda kv: -kv[1][1])[:12]],
        }
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                f.write("op\ttag\tlaunches\tms\tshare\tGB/s(algorithmic)\tTFLOP/s\n")
                for k, g in sorted(groups.items(), key=lambda kv: -kv[1][1]):
                    f.write(f"{k[0]}\t{k[1]}\t{g[0]}\t{g[1]:.3f}\t{g[1] / prof_total:.4f}\t{g[2] / max(g[1], 1e-9) / 1e6:.1f}\t{g[3] / max(g[1], 1e-9) / 1e9:.2f}\n")
                f.write(f"TOTAL\t\t{sum(g[0] for g in groups.values())}\t{prof_total:.3f}\n")
        # second half of BASELINE.json's metric: quantizer Gcodes/s on a sweep point (configs[4]), inputs resident in HBM
        try:
            line["quantizer"] = quantizer_point(dev, pk)
        except Exception as ex:  # pragma: no cover
            line["quantizer"] = {"error": repr(ex)}
        # BASELINE.json configs[3]: extract_embeddings = encode + quantize only (hierarchical code indices), same volume
        try:
            with torch.no_grad():
                for _ in range(3):
                    list(model.encode(x_dev))
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(steps):
                    list(model.encode(x_dev))
                a1.record()
                torch.cuda.synchronize()
            enc_ms = a0.elapsed_time(a1) / steps
            line["extract"] = {"value": world * batch * 1e3 / enc_ms, "unit": "volumes/s", "ms_per_step": enc_ms,
                               "what": "VQVAE.encode (Encoder2 + 3 quantizers -> code indices), CUDA-graph replay, per-rank time of rank 0"}
        except Exception as ex:  # pragma: no cover
            line["extract"] = {"error": repr(ex)}
        # BASELINE.json configs[1]: one training step of the 2-level downscaled model on a 256x256x128 volume (forward in training
        # mode with EMA updates, backward, fused Adam), captured as one CUDA graph; reported next to the headline, not part of it
        if not args.no_train and world == 1:
            try:
                line["train_step"] = train_point(dev)
            except Exception as ex:  # pragma: no cover
                line["train_step"] = {"error": repr(ex)}
        traffic = ncu_traffic(f"{dom_key[0]} [{dom_key[1]}]")
        if traffic is not None:
            line["roofline"]["traffic"] = traffic["dram_bytes_per_launch"]
            line["roofline"]["traffic_source"] = traffic["source"]
        if not args.no_cpu_baseline and world == 1:
            frac = (shape[2] * shape[3] * shape[4]) / (CPU_SAMPLE_SHAPE[2] * CPU_SAMPLE_SHAPE[3] * CPU_SAMPLE_SHAPE[4])
            times, cores = cpu_arm(kind, 3, 1)
            line["cpu_baseline"] = {
                "value": 1.0 / (min(times) * frac), "unit": "volumes/s", "cores": cores, "kind": "port",
                "median_value": 1.0 / (sorted(times)[len(times) // 2] * frac),
                "sample": f"oracle (ATen fp32) forward of the same model on a {CPU_SAMPLE_SHAPE[2]}x{CPU_SAMPLE_SHAPE[3]}x"
                          f"{CPU_SAMPLE_SHAPE[4]} crop = 1/{frac:g} of the voxels; best of 3, scaled by 1/{frac:g}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="full_512x512x128", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=8, help="independent volumes per step and GPU (stacked along B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step point (configs[1])")
    ap.add_argument("--profile-out", default=None, help="write the per-op CUDA-event table of one eager step here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
