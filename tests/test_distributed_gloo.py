"""N > 1 host logic on CPU: two processes, gloo backend, 127.0.0.1 rendezvous.

The product's kernels only run on a GPU, so each rank drives the SAME kernel sources through the host
emulator (tests/emu) -- what is under test here is the data-parallel logic of layers.py:645-647,670-676 as
rebuilt in `Quantizer._forward_impl`: ONE flat [counts | dw] all-reduce per level (the reference issues
two), the step-0 mean/std average, and that every rank ends the step with bit-identical codebook buffers
that equal the oracle run on the concatenated batch.  Also checks bench.py's volume sharding helper.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, first_pass, out):
    for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from emu.emu_ops import use_emulator
        from vqvae import layers as L
        K, D = 24, 6
        g = torch.Generator().manual_seed(123)
        embed = torch.randn(K, D, generator=g)
        x_all = torch.randn(world, D, 6, 5, 4, generator=g) * 1.5 + 0.3       # one (D, H, W, Z) latent grid per rank
        with use_emulator(), torch.no_grad():
            q = L.Quantizer(K, D, 0.1)
            q.embed.copy_(embed); q.embed_avg.copy_(embed); q.cluster_size.zero_(); q.first_pass.fill_(first_pass)
            q.train()
            loss, quant, idx = q(x_all[rank:rank + 1])
        out[rank] = dict(embed=q.embed.numpy().copy(), embed_avg=q.embed_avg.numpy().copy(), cluster_size=q.cluster_size.numpy().copy(),
                         first_pass=int(q.first_pass), idx=idx.numpy().copy(), loss=float(loss))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("first_pass", [0, 1])
def test_quantizer_data_parallel_two_ranks_gloo(first_pass):
    sys.path.insert(0, ROOT)
    from oracle import vqvae_oracle as O
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), first_pass, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    # ranks agree bit for bit after the collective (so DDP's per-forward buffer broadcast is unnecessary)
    for k in ("embed", "embed_avg", "cluster_size"):
        assert np.array_equal(r0[k], r1[k]), k
    assert r0["first_pass"] == 0 and r1["first_pass"] == 0
    # ... and equal the reference semantics on the concatenated batch
    K, D = 24, 6
    g = torch.Generator().manual_seed(123)
    embed = torch.randn(K, D, generator=g)
    x_all = torch.randn(world, D, 6, 5, 4, generator=g) * 1.5 + 0.3
    e = embed.numpy().astype(np.float64).copy()
    flat = [x_all[r].permute(1, 2, 3, 0).reshape(-1, D).numpy() for r in range(world)]
    n_tot = sum(f.shape[0] for f in flat)
    cs = np.zeros(K)
    if first_pass:      # layers.py:665-683: per-rank mean / unbiased std, SUM all-reduce, divide by world
        mean = np.mean([f.astype(np.float64).mean(0) for f in flat], axis=0)
        std = np.mean([f.astype(np.float64).std(0, ddof=1) for f in flat], axis=0)
        e = e * std + mean
        cs += n_tot / K
    ea = e.copy()
    e32 = e.astype(np.float32)
    n = np.zeros(K)
    dw = np.zeros((K, D))
    for r, f in enumerate(flat):
        idx, _ = O.vq_assign_c(f, e32)
        assert np.array_equal(idx, out[r]["idx"].reshape(-1)), f"rank {r} indices"
        nr, dwr = O.vq_stats_c(f, idx, K)
        n += nr
        dw += dwr
    cs = 0.99 * cs + 0.01 * n
    ea = 0.99 * ea + 0.01 * dw
    tot = cs.sum()
    sm = tot * (cs + 1e-5) / (tot + K * 1e-5)
    assert np.allclose(r0["cluster_size"], cs, rtol=1e-5, atol=1e-7)
    assert np.allclose(r0["embed_avg"], ea, rtol=1e-4, atol=1e-5)
    assert np.allclose(r0["embed"], ea / sm[:, None], rtol=1e-4, atol=1e-5)


def _grad_worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "3d-vq-vae-2_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vqvae.parallel import allreduce_gradients
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(5))]
        for i, p in enumerate(params[:2]):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        allreduce_gradients(params)          # params[2] has no grad: skipped consistently on every rank
        out[rank] = [None if p.grad is None else p.grad.numpy().copy() for p in params]
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_grad_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        assert np.allclose(out[r][0], 1.5) and np.allclose(out[r][1], 3.0) and out[r][2] is None   # mean of (1, 2) and (2, 4)


def test_bench_shards_volumes_across_ranks():
    sys.path.insert(0, ROOT)
    import bench
    seen = []
    for world in (1, 2, 4, 8):
        seeds = [bench.volume_seed(rank) for rank in range(world)]
        assert len(set(seeds)) == world                  # every rank works on its own volume(s)
        seen.append(seeds)
    assert seen[3][:2] == seen[1]                        # weak scaling: rank r's work does not depend on the world size
    line = bench.aggregate(world=4, steps=5, elapsed_ms_max=100.0)
    assert abs(line - 4 * 5 / 0.1) < 1e-9                # whole-job volumes/s over the slowest rank's time
