"""GPU parity of the backward kernels (csrc/backward_kernels.cu) through the module API: gradients of blocks and of a
whole training step against torch.autograd on the CPU oracle, and the fused Adam(amsgrad) step against torch.optim.Adam."""
import numpy as np
import pytest
import torch

from oracle import vqvae_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _fp32():
    from vqvae import _ops
    o = _ops.default()
    prev = o.precision
    o.precision = "fp32"
    yield
    o.precision = prev


@pytest.mark.parametrize("cin,cout,mode,shape", [
    (4, 4, "same", (1, 4, 12, 10, 16)),
    (18, 2, "same", (2, 18, 6, 5, 7)),
    (4, 8, "down", (1, 4, 12, 8, 16)),
    (16, 32, "down", (1, 16, 8, 8, 4)),
    (8, 4, "up", (1, 8, 6, 5, 4)),
    (18, 8, "up", (1, 18, 4, 4, 2)),
])
def test_preact_block_gradients(cin, cout, mode, shape):
    from vqvae import layers as L
    torch.manual_seed(cin * 13 + cout)
    blk = L.PreActFixupResBlock(cin, cout, mode)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.2))
        blk.scale.fill_(0.9)
    x = torch.randn(shape)
    sd = {"b." + k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    yr = O.preact_block(sd, "b.", xr, mode)
    r = torch.randn(yr.shape, generator=torch.Generator().manual_seed(3))
    (yr * r).sum().backward()
    blk = blk.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y = blk(xg)
    (y * r.to(DEV)).sum().backward()
    assert torch.allclose(y.detach().cpu(), yr.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(xg.grad.cpu(), xr.grad, rtol=2e-4, atol=2e-5), float((xg.grad.cpu() - xr.grad).abs().max())
    for k, p in blk.named_parameters():
        ref = sd["b." + k].grad
        assert torch.allclose(p.grad.cpu(), ref, rtol=1e-3, atol=1e-4), (k, float((p.grad.cpu() - ref).abs().max()), float(ref.abs().max()))


def test_training_steps_match_oracle_autograd_and_adam():
    """Two optimisation steps of a small 2-level model (loss = Huber + commitment, Adam amsgrad, eval-mode codebooks so the
    oracle sees the same quantizers): gradients, then parameters after each step."""
    from vqvae.model import VQVAE
    cfg = dict(n_bottleneck_blocks=2, num_embeddings=[16, 24], n_pre_quantization_blocks=2, n_post_quantization_blocks=2,
               n_post_upscale_blocks=1, n_post_downscale_blocks=1)
    torch.manual_seed(42)
    m = VQVAE(VQVAE.default_args(extract_center_cylinder=True, base_lr=1e-3, **cfg))
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
        for q in m.encoder.quantize:
            q.first_pass.fill_(0)
    m.eval()
    x = O.synthetic_volume((1, 1, 32, 32, 32))
    sd = {k: (v.detach().clone().requires_grad_(True) if v.dtype.is_floating_point and ".quantize." not in k else v.detach().clone())
          for k, v in m.state_dict().items()}
    leaves = {k: v for k, v in sd.items() if v.requires_grad}
    ref_opt = torch.optim.Adam(list(leaves.values()), lr=1e-3, amsgrad=True)
    m = m.to(DEV)
    opt = m.configure_optimizers()
    assert type(opt).__name__ == "FusedAdamAMSGrad"
    xd = x.to(DEV)
    for step in range(2):
        ref_opt.zero_grad()
        dec_r, (loss_r, _, idx_r) = O.vqvae_forward(sd, O.ModelConfig(**cfg), x)
        total_r, _ = O.huber_epilogue(dec_r, x, [30], list(loss_r), cylinder=True)
        total_r.backward()
        opt.zero_grad()
        loss, _ = m.huber((xd, [30]))
        loss.backward()
        assert abs(float(loss.detach()) - float(total_r.detach())) < 2e-4 * abs(float(total_r.detach())) + 1e-6
        for k, p in m.named_parameters():
            ref = leaves[k].grad
            err, scale = float((p.grad.cpu() - ref).abs().max()), float(ref.abs().max()) + 1e-6
            assert err <= 5e-3 * scale + 1e-6, (step, k, err, scale)
        ref_opt.step()
        opt.step()
        for k, p in m.named_parameters():
            assert torch.allclose(p.detach().cpu(), leaves[k].detach(), rtol=2e-4, atol=2e-5), (step, k)


def test_downscaled_model_training_step_runs():
    """BASELINE.json configs[1] at a reduced volume: the 2-level downscaled model (662 blocks) takes a full training step
    (forward in training mode with first-pass codebook init + EMA updates, backward, fused Adam); losses and gradients stay
    finite and every parameter moves."""
    from vqvae.model import VQVAE, downscaled_config_args
    from vqvae import _ops
    _ops.default().precision = "bf16"
    torch.manual_seed(0)
    args = downscaled_config_args()
    args.base_lr = 1e-4
    m = VQVAE(args).to(DEV).train()
    x = O.synthetic_volume((1, 1, 32, 32, 32)).to(DEV)
    opt = m.configure_optimizers()
    before = [p.detach().clone() for p in m.parameters()]
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss, _ = m.huber((x, [32]))
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)), losses
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    assert all(not torch.equal(a, p.detach()) for a, p in zip(before, m.parameters()))
    assert all(int(q.first_pass) == 0 for q in m.encoder.quantize)


def _gpu_grads_vs_oracle(blk, oracle_fn, mode, x, seed, tol):
    sd = {"b." + k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    yr = oracle_fn(sd, "b.", xr, mode)
    r = torch.randn(yr.shape, generator=torch.Generator().manual_seed(seed))
    (yr * r).sum().backward()
    blk = blk.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y = blk(xg)
    (y * r.to(DEV)).sum().backward()
    assert torch.allclose(y.detach().cpu(), yr.detach(), rtol=1e-4, atol=1e-5)
    gx = xg.grad.cpu()
    assert float((gx - xr.grad).abs().max()) <= tol * float(xr.grad.abs().max()) + 1e-6
    for k, p in blk.named_parameters():
        ref = sd["b." + k].grad
        assert float((p.grad.cpu() - ref).abs().max()) <= tol * float(ref.abs().max()) + 1e-6, (k, float(ref.abs().max()))


@pytest.mark.parametrize("cin,cout,mode,shape", [
    (3, 5, "same", (1, 3, 10, 9, 12)),
    (4, 6, "down", (1, 4, 12, 8, 16)),
    (6, 3, "up", (1, 6, 5, 4, 6)),
    (4, 2, "out", (2, 4, 6, 4, 5)),
])
def test_fixup_block_gradients(cin, cout, mode, shape):
    """FixupResBlock (`--block-type regular`, layers.py:219-303) incl. the trailing ELU's backward (vq3d_elu_backward)."""
    from vqvae import layers as L
    torch.manual_seed(cin * 7 + cout)
    blk = L.FixupResBlock(cin, cout, mode)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.2))
        blk.scale.fill_(0.9)
    _gpu_grads_vs_oracle(blk, O.fixup_block, mode, torch.randn(shape), 5, 1e-3)


@pytest.mark.parametrize("cin,cout,mode,shape", [
    (8, 8, "same", (1, 8, 10, 9, 12)),
    (16, 8, "same", (1, 16, 6, 8, 8)),
    (8, 16, "down", (1, 8, 8, 8, 12)),
    (16, 8, "up", (1, 16, 4, 5, 4)),
])
def test_evonorm_block_gradients(cin, cout, mode, shape):
    """EvonormResBlock / EvoNorm3DS0 backward (vq3d_evonorm_s0_backward_sums/_apply; evonorm.py:12-47,59-76 under autograd)."""
    from vqvae import layers as L
    torch.manual_seed(cin * 3 + cout)
    blk = L.EvonormResBlock(cin, cout, mode)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.endswith(".gamma"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape))
            elif n.endswith(".v") or n.endswith(".beta"):
                p.add_(0.2 * torch.randn(p.shape))
            else:
                p.copy_(torch.randn(p.shape) * 0.3)
    _gpu_grads_vs_oracle(blk, O.evonorm_block, mode, torch.randn(shape), 9, 3e-3)


def test_silu_velocity_known_answer_backward():
    """The reference's own test of this path (evonorm.py:79-98, test_silu_velocity): SiLU-velocity forward/backward of
    x * sigmoid(v * x) against autograd -- here through EvoNorm3DS0 with gamma/std factored out."""
    from vqvae.evonorm import EvoNorm3DS0
    torch.manual_seed(0)
    m = EvoNorm3DS0(8)
    with torch.no_grad():
        m.gamma.fill_(1.0)
        m.v.copy_(torch.randn(m.v.shape))
    x = torch.randn(1, 8, 6, 5, 7)
    xr = x.clone().requires_grad_(True)
    vr = m.v.detach().clone().requires_grad_(True)
    std = torch.sqrt(x.reshape(1, 1, -1).var(dim=-1, unbiased=True) + 1e-5).reshape(1, 1, 1, 1, 1)
    num = xr * torch.sigmoid(vr * xr)
    yr = num / torch.sqrt(xr.reshape(1, 1, -1).var(dim=-1, unbiased=True) + 1e-5).reshape(1, 1, 1, 1, 1)
    r = torch.randn(yr.shape, generator=torch.Generator().manual_seed(1))
    (yr * r).sum().backward()
    m = m.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    y = m(xg)
    (y * r.to(DEV)).sum().backward()
    assert torch.allclose(y.detach().cpu(), yr.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(xg.grad.cpu(), xr.grad, rtol=1e-3, atol=1e-4)
    assert torch.allclose(m.v.grad.cpu(), vr.grad, rtol=1e-3, atol=1e-4)


def test_graphed_training_step_matches_eager_steps():
    """vqvae.parallel.GraphedTrainingStep (whole step in one CUDA graph, device-side Adam step counter) against the same
    number of eager steps: same parameters and EMA buffers up to the atomics' summation order."""
    import copy
    from vqvae.model import VQVAE
    from vqvae.parallel import GraphedTrainingStep, training_step
    cfg = dict(n_bottleneck_blocks=2, n_downscales_per_bottleneck=1, num_embeddings=[16, 24], n_pre_quantization_blocks=2,
               n_post_quantization_blocks=2, n_post_upscale_blocks=1, n_post_downscale_blocks=1)
    torch.manual_seed(42)
    base = VQVAE(VQVAE.default_args(extract_center_cylinder=False, base_lr=1e-3, **cfg))
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in base.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    x = O.synthetic_volume((1, 1, 16, 16, 8)).to(DEV)
    batch = (x, [8])
    eager = copy.deepcopy(base).to(DEV).train()
    opt_e = eager.configure_optimizers()
    for _ in range(5):
        loss_e = training_step(eager, opt_e, batch)
    graphed = copy.deepcopy(base).to(DEV).train()
    opt_g = graphed.configure_optimizers()
    step = GraphedTrainingStep(graphed, opt_g, batch, warmup=2)
    for _ in range(3):
        loss_g = step(batch)
    torch.cuda.synchronize()
    assert abs(float(loss_g) - float(loss_e)) <= 1e-3 * abs(float(loss_e)) + 1e-6
    for (k, a), (_, b) in zip(eager.state_dict().items(), graphed.state_dict().items()):
        if a.dtype.is_floating_point:
            assert torch.allclose(a, b, rtol=2e-3, atol=2e-5), (k, float((a - b).abs().max()))
        else:
            assert torch.equal(a, b), k
    assert float(opt_g._step_state[0]) == 5.0


@pytest.mark.parametrize("cin,cout,k,circ,shape", [
    (36, 36, 3, True, (1, 36, 6, 5, 40)),      # accumulators split over input-channel chunks, partial tiles on every axis
    (72, 36, 1, False, (2, 72, 5, 4, 33)),     # k = 1, batch 2, halo tile bounded by shared memory
    (5, 7, 3, False, (1, 5, 9, 6, 35)),        # zero padding, odd channel counts
    (9, 9, 3, True, (1, 9, 9, 10, 40)),        # the row-sliding weight-gradient kernel (k3 s1, 9 / 8 / 4 / 2 / 1 output channels, depth >= 32)
    (18, 9, 3, True, (2, 18, 5, 6, 33)),       # 162 (ci, kh, kw) items, one row group, batch 2
    (30, 8, 3, False, (1, 30, 4, 7, 32)),      # input channels in two chunks (28 + 2), zero padding
    (4, 4, 3, True, (1, 4, 8, 8, 64)),         # 7 row groups
    (2, 2, 3, True, (1, 2, 6, 4, 32)),
    (1, 1, 3, True, (1, 1, 5, 5, 32)),         # 9 items, 16 row groups
])
def test_conv_gradients_tiled_wgrad_and_forward_dgrad(cin, cout, k, circ, shape):
    """The shared-memory tiled weight gradient and the forward-convolution input gradient against torch.autograd."""
    import torch.nn.functional as F
    from vqvae import _ops
    o = _ops.default()
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(shape, generator=g)
    w = torch.randn(cout, cin, k, k, k, generator=g) * 0.2
    a, b, s, pb = (torch.randn(1, generator=g) * 0.3 for _ in range(4))
    s = s + 1.0
    leaves = [t.clone().requires_grad_(True) for t in (x, w, a, b, s, pb)]
    xr, wr, ar, br, sr, pbr = leaves
    u = F.elu(xr + ar) + br
    pad = (k - 1) // 2
    if pad:
        u = F.pad(u, (pad,) * 6, mode="circular") if circ else F.pad(u, (pad,) * 6)
    yr = F.conv3d(u, wr) * sr + pbr
    r = torch.randn(yr.shape, generator=g)
    (yr * r).sum().backward()
    dl = [t.detach().clone().to(DEV).requires_grad_(True) for t in (x, w, a, b, s, pb)]
    y = o.conv3d(dl[0], dl[1], pad=pad, circular=circ, pre_act=True, pre_a=dl[2], pre_b=dl[3], post_scale=dl[4], post_b=dl[5])
    (y * r.to(DEV)).sum().backward()
    assert torch.allclose(y.detach().cpu(), yr.detach(), rtol=1e-4, atol=1e-4)
    for got, ref, name in zip(dl, leaves, ("x", "w", "pre_a", "pre_b", "scale", "post_b")):
        err, scale = float((got.grad.cpu() - ref.grad).abs().max()), float(ref.grad.abs().max())
        assert err <= 1e-3 * scale + 1e-5, (name, err, scale)


# ---- bf16 tensor-core mode (the product default, and what bench.py's train_step runs) --------------------------------
def _rel_grad_err(got, ref):
    """max |got - ref| relative to the reference tensor's max (north_star: 1e-2-level on BF16 paths; asserted at 2e-2)."""
    return float((got - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)


def _bf16_grad_violations(named_got_ref):
    """The stated bf16-mode gradient tolerance.  Weight tensors: max-abs error <= 2e-2 of the tensor's own max (+ 1e-3 of the
    largest weight gradient of the module, for tensors whose whole gradient is tiny).  Fixup scalars (bias*, scale): each is ONE
    number, a sum over every voxel of terms of both signs, so bf16 operand rounding gives it an absolute error set by the size
    of the terms, not of the (cancelled) sum: <= 2e-2 of its own magnitude + 1e-2 of the largest scalar gradient of the module
    (measured on the B200: worst 0.6e-2 of that scale).  Returns the offenders."""
    items = list(named_got_ref)
    wmax = max([float(r.abs().max()) for _, g, r in items if r.numel() > 1] + [0.0])
    smax = max([float(r.abs().max()) for _, g, r in items if r.numel() == 1] + [0.0])
    bad = {}
    for k, g, r in items:
        err, own = float((g - r).abs().max()), float(r.abs().max())
        tol = 2e-2 * own + (1e-2 * smax if r.numel() == 1 else 1e-3 * wmax)
        if err > tol:
            bad[k] = (err, own, tol)
    return bad


@pytest.mark.parametrize("cin,cout,mode,shape", [
    (32, 32, "same", (1, 32, 8, 8, 6)),        # conv2 16 -> 16 k3 on tcgen05, dgrad as a forward tensor-core convolution
    (72, 72, "same", (1, 72, 6, 5, 4)),
    (16, 32, "down", (1, 16, 8, 8, 4)),        # k4 s2 + k2 s2 skip
    (64, 128, "down", (1, 64, 4, 4, 4)),
    (18, 8, "up", (1, 18, 4, 4, 4)),
    (72, 32, "up", (1, 72, 4, 4, 2)),
])
def test_preact_block_gradients_bf16_mode(cin, cout, mode, shape):
    """Gradients of a recorded block with the GEMM-shaped convolutions (forward AND the forward-form input gradients) on the
    bf16 tensor-core kernels, against fp32 autograd on the oracle: every tensor within 2e-2 of its own max."""
    from vqvae import layers as L, _ops
    o = _ops.default()
    torch.manual_seed(cin * 13 + cout)
    blk = L.PreActFixupResBlock(cin, cout, mode)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 / np.sqrt(max(p.shape[1] if p.dim() > 1 else 1, 1) / 4) if p.dim() > 1 else 0.2))
        blk.scale.fill_(0.9)
    x = torch.randn(shape)
    sd = {"b." + k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    yr = O.preact_block(sd, "b.", xr, mode)
    r = torch.randn(yr.shape, generator=torch.Generator().manual_seed(3))
    (yr * r).sum().backward()
    blk = blk.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    o.precision = "bf16"
    o.profile = []
    try:
        y = blk(xg)
        (y * r.to(DEV)).sum().backward()
        torch.cuda.synchronize()
        names = [e[0] for e in o.profile]
    finally:
        o.profile = None
    assert "conv3d_tc" in names, names                       # the tensor-core path really ran
    assert _rel_grad_err(y.detach().cpu(), yr.detach()) <= 2e-2
    assert _rel_grad_err(xg.grad.cpu(), xr.grad) <= 2e-2
    bad = _bf16_grad_violations((k, p.grad.cpu(), sd["b." + k].grad) for k, p in blk.named_parameters())
    assert not bad, bad


def test_training_step_gradients_bf16_mode_vs_oracle_autograd():
    """Whole-model gradients in bf16 mode (loss = Huber + commitment, eval-mode codebooks): every parameter gradient within
    2e-2 of its tensor's max against fp32 autograd on the oracle.  bf16 operand rounding can move a latent across a near-tie;
    a flipped code changes the function being differentiated, so the comparison uses the first seed whose code indices agree
    with the oracle's (the mismatch rate itself is pinned in tests/test_gpu_full_config.py)."""
    from vqvae.model import VQVAE
    from vqvae import _ops
    o = _ops.default()
    cfg = dict(n_bottleneck_blocks=2, num_embeddings=[16, 24], n_pre_quantization_blocks=2, n_post_quantization_blocks=2,
               n_post_upscale_blocks=1, n_post_downscale_blocks=1, base_network_channels=8)
    chosen = None
    for seed in range(42, 50):
        torch.manual_seed(seed)
        m = VQVAE(VQVAE.default_args(extract_center_cylinder=True, base_lr=1e-3, **cfg))
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for p in m.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
            for q in m.encoder.quantize:
                q.first_pass.fill_(0)
        m.eval()
        x = O.synthetic_volume((1, 1, 32, 32, 32), seed=seed)
        sd = {k: (v.detach().clone().requires_grad_(True) if v.dtype.is_floating_point and ".quantize." not in k else v.detach().clone())
              for k, v in m.state_dict().items()}
        dec_r, (loss_r, _, idx_r) = O.vqvae_forward(sd, O.ModelConfig(**cfg), x)
        total_r, _ = O.huber_epilogue(dec_r, x, [30], list(loss_r), cylinder=True)
        m = m.to(DEV)
        o.precision = "bf16"
        o.profile = []
        try:
            loss, _ = m.huber((x.to(DEV), [30]))
            with torch.no_grad():
                _, (_, _, idx) = m(x.to(DEV))
            same = all(torch.equal(a.cpu(), b) for a, b in zip(idx, idx_r))
            if same:
                loss.backward()
                torch.cuda.synchronize()
            names = {e[0] for e in o.profile}
        finally:
            o.profile = None
        if same:
            chosen = seed
            break
    assert chosen is not None, "no seed in 42..49 gave oracle-identical code indices in bf16 mode"
    assert "conv3d_tc" in names
    total_r.backward()
    assert abs(float(loss.detach()) - float(total_r.detach())) <= 1e-2 * abs(float(total_r.detach()))
    errs = {k: _rel_grad_err(p.grad.cpu(), sd[k].grad) for k, p in m.named_parameters()}
    w_worst = max(v for k, v in errs.items() if k.endswith("weight"))
    print(f"\nbf16 training gradients (seed {chosen}): {len(errs)} tensors, worst weight-tensor error {w_worst:.3e} of the tensor's max")
    bad = _bf16_grad_violations((k, p.grad.cpu(), sd[k].grad) for k, p in m.named_parameters())
    assert not bad, bad


def test_optimizer_state_dict_resume_matches_uninterrupted_run(tmp_path):
    """FusedAdamAMSGrad (flat buffers, device-side step counter): save after 3 steps, load into a fresh optimizer, 2 more
    steps == 5 uninterrupted steps, and both equal torch.optim.Adam(amsgrad=True) on the same gradients."""
    from vqvae.optim import FusedAdamAMSGrad
    torch.manual_seed(0)
    shapes = [(7, 3, 3), (1,), (5, 4), (1,), (33,)]
    init = [torch.randn(s) for s in shapes]
    grads = [[torch.randn(s, generator=torch.Generator().manual_seed(100 * t + i)) for i, s in enumerate(shapes)] for t in range(5)]

    def run(opt, params, steps):
        for t in steps:
            for p, g in zip(params, grads[t]):
                p.grad.copy_(g.to(p.device)) if p.grad is not None else setattr(p, "grad", g.to(p.device).clone())
            opt.step()

    ref_p = [torch.nn.Parameter(t.clone()) for t in init]
    ref = torch.optim.Adam(ref_p, lr=1e-2, amsgrad=True)
    run(ref, ref_p, range(5))

    a_p = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    a = FusedAdamAMSGrad(a_p, lr=1e-2)
    assert a.flat_grad is not None
    run(a, a_p, range(3))
    assert a.current_step() == 3
    torch.save({"opt": a.state_dict(), "params": [p.detach().cpu() for p in a_p]}, tmp_path / "ck.pt")
    ck = torch.load(tmp_path / "ck.pt", weights_only=False)
    assert all(int(st["step"]) == 3 for st in ck["opt"]["state"].values())
    b_p = [torch.nn.Parameter(t.clone().to(DEV)) for t in ck["params"]]
    b = FusedAdamAMSGrad(b_p, lr=1e-2)
    b.load_state_dict(ck["opt"])
    assert b.current_step() == 3
    assert b.state[b_p[0]]["exp_avg"].data_ptr() == b._m.data_ptr()            # the entries alias the flat buffers again
    run(a, a_p, range(3, 5))
    run(b, b_p, range(3, 5))
    for pa, pb, pr in zip(a_p, b_p, ref_p):
        assert torch.equal(pa, pb)
        assert torch.allclose(pa.detach().cpu(), pr.detach(), rtol=1e-5, atol=1e-6)
    # a torch.optim.Adam state_dict of the same parameters loads too
    c_p = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
    c = FusedAdamAMSGrad(c_p, lr=1e-2)
    ref3_p = [torch.nn.Parameter(t.clone()) for t in init]
    ref3 = torch.optim.Adam(ref3_p, lr=1e-2, amsgrad=True)
    run(ref3, ref3_p, range(3))
    c.load_state_dict(ref3.state_dict())
    with torch.no_grad():
        for p, r in zip(c_p, ref3_p):
            p.copy_(r.to(DEV))
    run(c, c_p, range(3, 5))
    for pc, pr in zip(c_p, ref_p):
        assert torch.allclose(pc.detach().cpu(), pr.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("C,shape", [(4, (1, 4, 12, 10, 16)), (18, (2, 18, 6, 5, 7)), (2, (1, 2, 8, 8, 2)), (8, (1, 8, 20, 9, 33)), (32, (1, 32, 6, 6, 4))])
def test_fused_same_block_backward_vs_oracle_autograd(C, shape):
    """vq3d_preact_same_backward (one fused forward that keeps only x + two tiled backward kernels + the tiled d W2): every
    gradient of a 'same' block against autograd on the oracle.  (The path is off by default -- slower than the composed one inside
    a CUDA graph, DESIGN.md 8 -- and enabled here.)"""
    from vqvae import layers as L, _ops
    o = _ops.default()
    torch.manual_seed(C)
    blk = L.PreActFixupResBlock(C, C, "same")
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.2))
        blk.scale.fill_(0.9)
    x = torch.randn(shape)
    sd = {"b." + k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    yr = O.preact_block(sd, "b.", xr, "same")
    r = torch.randn(yr.shape, generator=torch.Generator().manual_seed(3))
    (yr * r).sum().backward()
    blk = blk.to(DEV)
    xg = x.to(DEV).requires_grad_(True)
    prev, o.fused_block_bwd = o.fused_block_bwd, True
    o.profile = []
    try:
        y = blk(xg)
        (y * r.to(DEV)).sum().backward()
        torch.cuda.synchronize()
        names = [e[0] for e in o.profile]
    finally:
        o.fused_block_bwd = prev
        o.profile = None
    assert names == ["preact_block", "preact_same_backward"], names
    assert torch.allclose(y.detach().cpu(), yr.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(xg.grad.cpu(), xr.grad, rtol=2e-4, atol=2e-5), float((xg.grad.cpu() - xr.grad).abs().max())
    for k, p in blk.named_parameters():
        ref = sd["b." + k].grad
        assert torch.allclose(p.grad.cpu(), ref, rtol=1e-3, atol=1e-4), (k, float((p.grad.cpu() - ref).abs().max()), float(ref.abs().max()))
