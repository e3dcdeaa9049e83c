#!/usr/bin/env python
"""bench.py -- encode -> quantize -> decode throughput of the Full 3-level 3D VQ-VAE-2 on
synthetic 512x512x128 volumes (BASELINE.json metric), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun)
    python bench.py --impl reference ...                     (the oracle on the host cores)

A "step" = one forward pass (Encoder2 -> 3 quantizers -> Decoder, model.py:79-83) over one batch of
--batch independent volumes per GPU (default 16, stacked along B); volumes are independent, so ranks
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
# bounded samples for the CPU arm (the Full model needs extents divisible by 64): the whole volume, a quarter, 1/32 of it
CPU_SAMPLES = [(512, 512, 128), (256, 256, 128), (128, 128, 64)]
DTYPE = "bf16 operands / fp32 accumulate (GEMM-shaped convolutions, tcgen05) + fp32 (thin convolutions, quantizer)"


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
    # third ceiling of this design: every score has to leave TMEM (N*K*4 bytes).  tcgen05.ld measured on this part
    # (tools/microbench/tc_microbench.cu -> profiles/r02_tc_microbench.txt): 460 B/clk per SM with the kernel's 16 sweep warps
    # (32x32b.x64), 390 B/clk with its x16 loads -- not the 64 B/clk that round 1 assumed
    tmem = 148 * 390 * 1.965 / (4.0 * K)
    return {"value": gc, "unit": "Gcodes/s", "config": {"N": N, "D": D, "K": K, "mode": "eval", "path": "tcgen05 bf16 candidate pass + exact fp32 re-rank"},
            "ms": t * 1e3, "roofline": {"bound": "hbm" if hbm < tens else "tensor", "peak": min(hbm, tens), "unit": "Gcodes/s", "frac": gc / min(hbm, tens),
                                        "tmem_read_ceiling": tmem, "frac_of_tmem_read_ceiling": gc / tmem,
                                        "tmem_read_ceiling_source": "measured tcgen05.ld throughput, profiles/r02_tc_microbench.txt"},
            "sweep": "profiles/r02_quantizer_sweep.tsv (tools/bench_quantizer.py, N up to 64 M)"}


def train_point(dev, workload, steps=3, world=1, rank=0):
    """One optimisation step (forward in training mode with EMA codebook updates, Huber + commitment loss, backward, gradient
    all-reduce, fused Adam(amsgrad)) of `workload`, batch 1 per GPU like the reference (train_vqvae_3d.job:76), captured as
    ONE CUDA graph.  With world > 1 every rank calls this: the flat gradient all-reduce and the three flat [counts | dw] EMA
    all-reduces (layers.py:645-647) are NCCL calls inside the captured step; the step time is the max over ranks."""
    import torch.distributed as dist
    from vqvae.parallel import GraphedTrainingStep, training_step
    kind, shape = WORKLOADS[workload]
    torch.cuda.empty_cache()
    m = build_model(kind).to(dev).train()
    for q in m.encoder.quantize:
        q.first_pass.fill_(1)
    x = synthetic_volume(shape, volume_seed(rank)).to(dev)
    opt = m.configure_optimizers()
    batch = (x, [shape[4]])
    torch.cuda.reset_peak_memory_stats()
    training_step(m, opt, batch)                      # data-dependent codebook init (layers.py:665-683)
    graphed = True
    try:
        step = GraphedTrainingStep(m, opt, batch, warmup=2)
    except Exception as ex:                           # pragma: no cover - capture refused (e.g. NCCL build without graph support)
        graphed = repr(ex)[:200]
        step = lambda b: training_step(m, opt, b)
        step(batch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    grad_bytes = opt.flat_grad.numel() * 4
    ema_bytes = sum(4 * q.num_embeddings * (q.embedding_dim + 1) for q in m.encoder.quantize)
    ar = None
    if world > 1:                                     # the two exchanges on their own (same buffers, NCCL over NVLink)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        g = torch.zeros_like(opt.flat_grad)
        stats = [torch.zeros(q.num_embeddings * (q.embedding_dim + 1), device=dev) for q in m.encoder.quantize]
        for _ in range(3):
            dist.all_reduce(g)
        torch.cuda.synchronize()
        a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a0.record()
        for _ in range(10):
            dist.all_reduce(g)
        a1.record()
        for _ in range(10):
            for st in stats:
                dist.all_reduce(st)
        a2.record()
        torch.cuda.synchronize()
        ar = {"gradient_allreduce_us": 100.0 * a0.elapsed_time(a1), "gradient_bytes": grad_bytes,
              "gradient_busbw_GBps": grad_bytes * 2 * (world - 1) / world / (a0.elapsed_time(a1) * 1e-4) / 1e9,
              "ema_allreduce_us": 100.0 * a1.elapsed_time(a2), "ema_bytes": ema_bytes, "ema_calls_per_step": len(stats),
              "backend": "nccl", "inside_timed_step": True}
    out = {"workload": workload, "ms_per_step": ms, "volumes_per_s": world * 1e3 / ms, "batch_per_gpu": 1, "n_gpus": world,
           "loss": float(loss), "cuda_graph": graphed, "steps": steps, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
           "parameters": opt.flat_grad.numel(), "collectives": ar,
           "what": "forward (training mode, EMA codebook update) + backward + gradient/EMA all-reduce + fused Adam(amsgrad); fp32 master "
                   "weights, bf16 tcgen05 convolutions where GEMM-shaped"}
    del step, m, opt
    torch.cuda.empty_cache()
    return out


def ncu_traffic(kernel_tag):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture of that very launch (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the capture's raw page)."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return None
    for key, val in table.items():
        if key in kernel_tag:
            return val
    return None


# ---------------------------------------------------------------------------------------
def cpu_forward_fn(kind):
    """The reference's CPU path for this workload: the oracle port (ATen fp32 kernels, the same F.* calls in the same order as
    the reference modules) on all host cores.  Returns run(sample_hwd) -> (seconds, outputs)."""
    from oracle import vqvae_oracle as O
    if _ORIGINAL_AFFINITY is not None:          # the CPU arm uses every host core, not only the GPU-local ones
        os.sched_setaffinity(0, _ORIGINAL_AFFINITY)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = build_model(kind)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    cfg = O.FULL if kind == "full" else O.DOWNSCALED

    def run(hwd, seed=42):
        x = synthetic_volume((1, 1) + tuple(hwd), seed)
        with torch.no_grad():
            t0 = time.perf_counter()
            out = O.vqvae_forward(sd, cfg, x)
            return time.perf_counter() - t0, out
    return run, cores


def run_reference(args):
    """`--impl reference`: every step is ONE oracle forward of a bounded sample of the workload's volume -- the whole volume if
    K + W of them fit the budget, else a quarter or 1/32 of it (picked from one untimed probe).  `ms_per_step` is the measured
    time of a step as run; `value` converts sample forwards to whole volumes by the voxel ratio, named in `config`."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    kind, shape = WORKLOADS[args.workload]
    run, cores = cpu_forward_fn(kind)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    full_vox = shape[2] * shape[3] * shape[4]
    samples = [hwd for hwd in CPU_SAMPLES if hwd[0] * hwd[1] * hwd[2] <= full_vox] or [tuple(shape[2:])]
    probe_hwd = samples[-1]
    probe_s, _ = run(probe_hwd)                                    # untimed probe (also warms the allocator / thread pool)
    probe_s, _ = run(probe_hwd)
    budget_s = float(os.environ.get("VQ3D_REFERENCE_BUDGET_S", "150"))
    pick = probe_hwd
    for hwd in samples:                                            # largest first
        est = probe_s * (hwd[0] * hwd[1] * hwd[2]) / (probe_hwd[0] * probe_hwd[1] * probe_hwd[2])
        if est * (steps + warmup) <= budget_s:
            pick = hwd
            break
    frac = full_vox / (pick[0] * pick[1] * pick[2])
    for _ in range(warmup):
        run(pick)
    times = [run(pick)[0] for _ in range(steps)]
    total = sum(times)
    value = steps / (total * frac)
    sample = (f"{kind} model forward (oracle port, ATen fp32, {cores} threads) on a {pick[0]}x{pick[1]}x{pick[2]} "
              + ("volume (the whole workload volume)" if frac == 1 else f"crop = 1/{frac:g} of the volume's voxels; volumes/s = crops/s / {frac:g}")
              + " per step")
    print(json.dumps({
        "impl": "reference", "metric": "volumes_per_s_encode_vq_decode", "value": value, "unit": "volumes/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * total / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "extrapolated": frac != 1, "sample_fraction_of_volume": 1.0 / frac,
        "config": {"workload": args.workload, "volume": list(shape), "batch_per_gpu": args.batch,
                   "workload_sample": list(pick), "volumes_per_step": 1.0 / frac},
        "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ---------------------------------------------------------------------------------------
def pipelined(fwd, host_in, n_steps, barrier, dev, check=None):
    """End-to-end loop with HOST buffers: every step copies `host_in` (pinned) to the device, runs fwd(device input) -> list
    of device tensors, and copies those to pinned host buffers.  Copies ride on their own streams (double-buffered input, staged
    outputs) so that the H2D of step i+1 and the D2H of step i-1 overlap the compute of step i; the clock (host wall clock
    between barriers, the copies are part of it) runs until the last result is in host memory.  Returns (seconds for n_steps,
    h2d bytes per step, d2h bytes per step, host outputs)."""
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    x_bufs = [torch.empty(host_in.shape, dtype=host_in.dtype, device=dev) for _ in range(2)]
    ev_in = [torch.cuda.Event(), torch.cuda.Event()]
    ev_fwd_read = [torch.cuda.Event(), torch.cuda.Event()]     # fwd has consumed x_bufs[i]
    ev_staged, ev_out = torch.cuda.Event(), torch.cuda.Event()
    stage = host_out = None
    t0 = None
    for it in range(2 + n_steps):
        if it == 2:
            barrier()
            t0 = time.perf_counter()
        b = it & 1
        with torch.cuda.stream(s_in):
            if it >= 2:
                s_in.wait_event(ev_fwd_read[b])
            x_bufs[b].copy_(host_in, non_blocking=True)
            ev_in[b].record(s_in)
        main.wait_event(ev_in[b])
        outs = fwd(x_bufs[b])
        ev_fwd_read[b].record(main)
        if stage is None:
            stage = [torch.empty_like(o) for o in outs]
            host_out = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
        if it > 0:
            main.wait_event(ev_out)              # the previous result has left the staging buffers
        for st, o in zip(stage, outs):
            st.copy_(o, non_blocking=True)
        ev_staged.record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_staged)
            for h, st in zip(host_out, stage):
                h.copy_(st, non_blocking=True)
            ev_out.record(s_out)
    torch.cuda.synchronize()
    barrier()
    secs = time.perf_counter() - t0
    h2d = host_in.numel() * host_in.element_size()
    d2h = sum(h.numel() * h.element_size() for h in host_out)
    return secs, h2d, d2h, host_out


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

        # the measured path: the public API with CUDA graphs on (VQ3D_BENCH_NO_GRAPH=1: eager launches, so that an ncu
        # launch list of this command sees the step's kernels one by one -- tools/gpu_launch_list.sh; never a bench value)
        model.enable_cuda_graphs(not os.environ.get("VQ3D_BENCH_NO_GRAPH"))
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

        dom_key = max(groups, key=lambda k: groups[k][1])
        dom = groups[dom_key]

        # ---- end to end through the public API with HOST buffers (pinned): H2D + forward + D2H inside the clock ----
        e2e_steps = max(3, min(steps, 20))

        def fwd_api(xb):                     # VQVAE.forward: fp32 volume in, fp32 reconstruction + int64 code indices out
            dec, (_, _, idxs) = model(xb)
            return [dec] + list(idxs)
        e2e_s, h2d, d2h, host_out = pipelined(fwd_api, x_host, e2e_steps, barrier, dev)
        # the pipelined loop returned the same results as a plain call on the same volumes (the forward has no atomics, so
        # it is bit-reproducible; the tolerance only covers a change of the split-K plan between captures)
        chk, (_, _, chk_idx) = model(x_dev)
        bad = float(((chk.cpu() - host_out[0]).abs() > 1e-3 + 1e-3 * host_out[0].abs()).float().mean())
        assert bad < 1e-2, f"e2e pipeline result mismatch ({bad:.2e} of the voxels)"
        assert all(float((a.cpu() != h).float().mean()) < 1e-2 for a, h in zip(chk_idx, host_out[1:])), "e2e pipeline index mismatch"
        del host_out, chk

        # the same volumes as raw int16 Hounsfield units in, int16 Hounsfield units + code indices out (VQVAE.reconstruct_hu:
        # the data module's clip/scale/shift and decode_embeddings' ELU*1000-1000+rint run on the device): 2 bytes per voxel
        # each way instead of 4
        hu_host = torch.clamp(torch.round((x_host - 1.0) * 1000.0), -32768, 32767).to(torch.int16).pin_memory()

        def fwd_hu(hb):
            hu, idxs = model.reconstruct_hu(hb)
            return [hu] + list(idxs)
        hu_s, hu_h2d, hu_d2h, hu_out = pipelined(fwd_hu, hu_host, e2e_steps, barrier, dev)
        del hu_out

        # BASELINE.json configs[3]: extract_embeddings end to end -- fp32 volume in, ONLY the hierarchical code indices out
        def fwd_extract(xb):
            return [t[2] for t in model.encode(xb)]
        ex_s, ex_h2d, ex_d2h, _ = pipelined(fwd_extract, x_host, e2e_steps, barrier, dev)

        # the box's ceiling for the fp32 API's traffic: the same buffers and streams with NO forward between the copies
        copy_outs = [torch.empty(bshape, dtype=torch.float32, device=dev)] + [torch.empty_like(i) for i in chk_idx]
        cp_s, _, _, _ = pipelined(lambda xb: copy_outs, x_host, e2e_steps, barrier, dev)
        del copy_outs

        # batch 1 (the reference's own operating point, train_vqvae_3d.job:76): latency of one volume, CUDA-graph replay
        x1 = x_dev[:1].contiguous()
        for _ in range(3):
            model(x1)
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(max(5, steps)):
            model(x1)
        b1.record()
        torch.cuda.synchronize()
        batch1_ms = b0.elapsed_time(b1) / max(5, steps)

        # encode + quantize only (extract_embeddings), inputs resident in HBM
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

    t = torch.tensor([elapsed_ms, e2e_s, hu_s, ex_s, cp_s, enc_ms, batch1_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_s, hu_s, ex_s, cp_s, enc_ms, batch1_ms = (float(v) for v in t)

    # BASELINE.json configs[1] / configs[2]: training steps.  configs[2] (Full model, 512x512x128, data parallel) runs on EVERY
    # rank with the NCCL gradient + EMA all-reduces inside the captured step; configs[1] (downscaled, 1 GPU) on rank 0 at N = 1
    train = {}
    if not args.no_train:
        model.enable_cuda_graphs(False)          # drop the inference graphs' static pools before the training steps allocate
        del out
        torch.cuda.empty_cache()
        try:
            train["full_512x512x128_dp"] = train_point(dev, "full_512x512x128", steps=3, world=world, rank=rank)
        except Exception as ex:  # pragma: no cover
            train["full_512x512x128_dp"] = {"error": repr(ex)[:300]}
        if world == 1:
            try:
                train["downscaled_256x256x128"] = train_point(dev, "downscaled_256x256x128", steps=3)
            except Exception as ex:  # pragma: no cover
                train["downscaled_256x256x128"] = {"error": repr(ex)[:300]}
        model.enable_cuda_graphs()

    if rank == 0:
        pk, pk_src = peaks()
        n_dom, t_dom, b_dom, f_dom = dom
        achieved = b_dom / (t_dom * 1e-3) / 1e9            # GB/s, algorithmic bytes / event time
        intensity = f_dom / max(b_dom, 1)
        vps = lambda secs: world * e2e_steps * batch / secs
        line = {
            "metric": "volumes_per_s_encode_vq_decode", "value": aggregate(world, steps, elapsed_ms, batch), "unit": "volumes/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": elapsed_ms / steps,
            "ms_per_volume": elapsed_ms / steps / batch,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": {"workload": args.workload, "volume": list(shape), "batch_per_gpu": batch,
                       "model": "3-level Full (train_vqvae_3d.job flags)" if kind == "full" else "2-level downscaled",
                       "weights": "reference ctor RNG seed 42 + Fixup init + N(0,0.02) perturbation",
                       "l2": f"inputs larger than L2 ({134 * batch} MB of volumes, GBs of activations per step)",
                       "cuda_graph": True, "parallelism": f"dp{world} (independent volumes, no collective on the inference path)",
                       "e2e": "pinned H2D / forward / D2H on three streams, double-buffered",
                       "host_cpus_bound": numa_cpus},
            "e2e": {"value": vps(e2e_s), "unit": "volumes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "api": "VQVAE.forward (fp32 volume in; fp32 reconstruction + int64 code indices out)",
                    "copy_only": {"value": vps(cp_s), "unit": "volumes/s", "GBps_each_way_per_gpu": (h2d + d2h) / 2 * e2e_steps / cp_s / 1e9,
                                  "what": "the same pinned buffers, streams and bytes with no forward in between: the host<->device ceiling of this box at this N"}},
            "e2e_hu_int16": {"value": vps(hu_s), "unit": "volumes/s", "h2d_bytes_per_step": hu_h2d, "d2h_bytes_per_step": hu_d2h,
                             "ms_per_step": 1e3 * hu_s / e2e_steps,
                             "api": "VQVAE.reconstruct_hu (int16 Hounsfield units in; int16 Hounsfield units + int64 code indices out; "
                                    "clip/scale/shift and ELU*1000-1000+rint on the device)"},
            "batch1": {"value": 1e3 / batch1_ms, "unit": "volumes/s", "ms_per_volume": batch1_ms,
                       "what": "one 512x512x128 volume per step (the reference's batch size), CUDA-graph replay, per GPU"},
            "gpu_launches": launches_per_step * steps,
            "clocks": sampler.result(),
            "roofline": {"bound": "hbm" if intensity < 210 else "tensor", "achieved": achieved, "peak": pk["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": None, "peak_source": pk_src,
                         "kernel": f"{dom_key[0]} [{dom_key[1]}]", "launches_per_step": n_dom,
                         "avg_launch_us": 1e3 * t_dom / n_dom, "share_of_step": t_dom / prof_total,
                         "algorithmic_bytes_per_launch": b_dom / n_dom, "flop_per_byte": intensity,
                         "whole_step": {"algorithmic_GB": 12.77 * batch if kind == "full" else None,
                                        "achieved_GBps": (12.77 * batch / (elapsed_ms / steps * 1e-3)) if kind == "full" else None,
                                        "frac_of_hbm": (12.77 * batch / (elapsed_ms / steps * 1e-3) / pk["hbm_gbs"]) if kind == "full" else None,
                                        "what": "SURVEY 8d block-fused fp32 byte count (12.77 GB per volume) / step time"}},
            "kernel_shares": {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / prof_total, 4)}
                              for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1][1])},
            "top_ops": [{"op": f"{k[0]} [{k[1]}]", "n": g[0], "ms": round(g[1], 3),
                         "GBps": round(g[2] / max(g[1], 1e-9) / 1e6, 1), "TFLOPs": round(g[3] / max(g[1], 1e-9) / 1e9, 2)}
                        for k, g in sorted(groups.items(), key=lambda kv: -kv[1][1])[:12]],
            "extract": {"value": world * batch * 1e3 / enc_ms, "unit": "volumes/s", "ms_per_step": enc_ms,
                        "what": "VQVAE.encode (Encoder2 + 3 quantizers -> code indices), CUDA-graph replay, inputs resident in HBM, max over ranks",
                        "e2e": {"value": vps(ex_s), "unit": "volumes/s", "h2d_bytes_per_step": ex_h2d, "d2h_bytes_per_step": ex_d2h,
                                "what": "fp32 volume from pinned host memory in, only the three int64 code-index tensors out (extract_embeddings.py:62-76)"}},
        }
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                f.write("op\ttag\tlaunches\tms\tshare\tGB/s(algorithmic)\tTFLOP/s\n")
                for k, g in sorted(groups.items(), key=lambda kv: -kv[1][1]):
                    f.write(f"{k[0]}\t{k[1]}\t{g[0]}\t{g[1]:.3f}\t{g[1] / prof_total:.4f}\t{g[2] / max(g[1], 1e-9) / 1e6:.1f}\t{g[3] / max(g[1], 1e-9) / 1e9:.2f}\n")
                f.write(f"TOTAL\t\t{sum(g[0] for g in groups.values())}\t{prof_total:.3f}\n")
        # second half of BASELINE.json's metric: quantizer Gcodes/s on two sweep points (configs[4]), inputs resident in HBM
        try:
            line["quantizer"] = quantizer_point(dev, pk)
            line["quantizer"]["second_point"] = quantizer_point(dev, pk, N=1 << 22, D=128, K=4096, reps=3)
        except Exception as ex:  # pragma: no cover
            line["quantizer"] = {"error": repr(ex)}
        if train:
            line["train_step"] = train
        traffic = ncu_traffic(f"{dom_key[0]} [{dom_key[1]}]")
        if traffic is not None:
            line["roofline"]["traffic"] = traffic["dram_bytes_per_launch"]
            line["roofline"]["traffic_source"] = traffic["source"]
        if not args.no_cpu_baseline and world == 1:
            # the oracle on ONE whole volume of this workload (volume 0 of rank 0) on all host cores: the CPU baseline, and the
            # checker for the GPU's code indices on the benchmarked configuration (both precision modes)
            run_cpu, cores = cpu_forward_fn(kind)
            secs, (ref_dec, (_, _, ref_idx)) = run_cpu(tuple(shape[2:]), seed=volume_seed(rank, 0, batch))
            line["cpu_baseline"] = {"value": 1.0 / secs, "unit": "volumes/s", "cores": cores, "kind": "port",
                                    "sample": f"oracle (ATen fp32, {cores} threads) forward of the same model on ONE whole "
                                              f"{shape[2]}x{shape[3]}x{shape[4]} volume (volume 0 of the batch), {secs:.1f} s, no extrapolation"}
            mism = {}
            model.enable_cuda_graphs(False)
            with torch.no_grad():
                for mode in ("bf16", "fp32"):
                    prev, ops.precision = ops.precision, mode
                    try:
                        dec1, (_, _, idx1) = model(x_dev[:1].contiguous())
                        mism[mode] = {"index_mismatch_bottom_to_top": [float((a.cpu() != b).float().mean()) for a, b in zip(idx1, ref_idx)],
                                      "decoded_mean_abs_err_rel_to_range": float((dec1.cpu() - ref_dec).abs().mean() / ref_dec.abs().max())}
                    finally:
                        ops.precision = prev
            line["extract"]["index_mismatch_vs_oracle"] = mism
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="full_512x512x128", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=16, help="independent volumes per step and GPU (stacked along B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step points (configs[1], configs[2])")
    ap.add_argument("--profile-out", default=None, help="write the per-op CUDA-event table of one eager step here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
