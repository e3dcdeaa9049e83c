"""Gradients of the differentiable generic ops (csrc/backward_kernels.cu) against torch.autograd run on the oracle
(CPU, host emulator of the same kernel sources).  GPU counterpart: tests/test_gpu_backward.py."""
import numpy as np
import pytest
import torch

from emu.emu_ops import use_emulator
from oracle import vqvae_oracle as O
from vqvae import layers as L


def _block_grads(blk, x, r, emu):
    """(dx, {param: grad}) of sum(block(x) * r)."""
    for p in blk.parameters():
        p.grad = None
    xg = x.clone().requires_grad_(True)
    y = blk(xg)
    (y * r).sum().backward()
    return y.detach(), xg.grad, {k: p.grad.clone() for k, p in blk.named_parameters()}


@pytest.mark.parametrize("cin,cout,mode,shape", [
    (4, 4, "same", (1, 4, 5, 6, 4)),
    (6, 3, "same", (2, 6, 4, 3, 5)),        # skip conv, batch 2
    (4, 8, "down", (1, 4, 6, 4, 8)),
    (8, 4, "up", (1, 8, 3, 4, 2)),
    (8, 8, "same", (1, 8, 5, 3, 33)),       # depth >= 32: conv2's weight gradient on the row-sliding kernel (4 channels), partial tiles
])
def test_preact_block_gradients_vs_oracle_autograd(cin, cout, mode, shape):
    torch.manual_seed(cin * 13 + cout)
    blk = L.PreActFixupResBlock(cin, cout, mode)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.2))
        blk.scale.fill_(0.9)
    x = torch.randn(shape)
    with use_emulator():
        xg = x.clone().requires_grad_(True)
        y = blk(xg)
        r = torch.randn(y.shape, generator=torch.Generator().manual_seed(3))
        (y * r).sum().backward()
        got_dx = xg.grad.clone()
        got = {k: p.grad.clone() for k, p in blk.named_parameters()}
    # oracle: the same block as pure torch functions of a state dict whose leaves require grad
    sd = {"b." + k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    yr = O.preact_block(sd, "b.", xr, mode)
    assert torch.allclose(y.detach(), yr.detach(), rtol=1e-4, atol=1e-5)
    (yr * r).sum().backward()
    assert torch.allclose(got_dx, xr.grad, rtol=2e-4, atol=2e-5), float((got_dx - xr.grad).abs().max())
    for k, g in got.items():
        ref = sd["b." + k].grad
        assert ref is not None, k
        assert torch.allclose(g, ref, rtol=5e-4, atol=5e-5), (k, float((g - ref).abs().max()), float(ref.abs().max()))


def test_tiny_model_training_step_gradients_vs_oracle():
    """One training step of a tiny 2-level pre-activation model: loss = Huber + sum(commitment); gradients of every
    parameter against autograd on the oracle; then one fused-Adam-free sanity check that the loss is finite."""
    from vqvae.model import VQVAE
    cfg = dict(n_bottleneck_blocks=2, n_downscales_per_bottleneck=1, num_embeddings=[8, 12], n_pre_quantization_blocks=1,
               n_post_quantization_blocks=1, n_post_upscale_blocks=0, n_post_downscale_blocks=1)
    torch.manual_seed(42)
    m = VQVAE(VQVAE.default_args(extract_center_cylinder=False, **cfg))
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
        for q in m.encoder.quantize:
            q.first_pass.fill_(0)
    m.eval()            # eval: the EMA buffers stay fixed, so the oracle sees the same codebooks
    x = O.synthetic_volume((1, 1, 8, 8, 8))
    sd = {k: (v.detach().clone().requires_grad_(True) if v.dtype.is_floating_point and ".quantize." not in k else v.detach().clone())
          for k, v in m.state_dict().items()}
    dec_r, (loss_r, _, idx_r) = O.vqvae_forward(sd, O.ModelConfig(**cfg), x)
    total_r, recon_r = O.huber_epilogue(dec_r, x, [8], list(loss_r), cylinder=False)
    total_r.backward()
    with use_emulator():
        loss, log = m.huber((x, [8]))
        loss.backward()
    assert abs(float(loss.detach()) - float(total_r.detach())) < 1e-4 * abs(float(total_r)) + 1e-6
    worst = 0.0
    for k, p in m.named_parameters():
        ref = sd[k].grad
        assert p.grad is not None and ref is not None, k
        err = float((p.grad - ref).abs().max())
        scale = float(ref.abs().max()) + 1e-6
        worst = max(worst, err / scale)
        assert err <= 2e-3 * scale + 1e-6, (k, err, scale)


def _grads_vs_oracle(blk, oracle_fn, mode, x, seed, tol=5e-4):
    """Gradients of sum(block(x) * r) wrt x and every parameter: emulator kernels vs torch.autograd on the oracle."""
    with use_emulator():
        xg = x.clone().requires_grad_(True)
        y = blk(xg)
        r = torch.randn(y.shape, generator=torch.Generator().manual_seed(seed))
        (y * r).sum().backward()
        got_dx = xg.grad.clone()
        got = {k: p.grad.clone() for k, p in blk.named_parameters()}
    sd = {"b." + k: v.detach().clone().requires_grad_(True) for k, v in blk.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    yr = oracle_fn(sd, "b.", xr, mode)
    assert torch.allclose(y.detach(), yr.detach(), rtol=1e-4, atol=1e-5)
    (yr * r).sum().backward()
    assert torch.allclose(got_dx, xr.grad, rtol=tol, atol=tol * 0.1 * float(xr.grad.abs().max())), float((got_dx - xr.grad).abs().max())
    for k, g in got.items():
        ref = sd["b." + k].grad
        assert ref is not None, k
        assert float((g - ref).abs().max()) <= tol * float(ref.abs().max()) + 1e-6, (k, float((g - ref).abs().max()), float(ref.abs().max()))


@pytest.mark.parametrize("cin,cout,mode,shape", [
    (3, 5, "same", (1, 3, 4, 5, 4)),
    (4, 6, "down", (1, 4, 6, 4, 8)),
    (6, 3, "up", (1, 6, 3, 4, 2)),
    (4, 2, "out", (2, 4, 3, 4, 5)),        # no trailing ELU, batch 2
])
def test_fixup_block_gradients_vs_oracle_autograd(cin, cout, mode, shape):
    """FixupResBlock (`--block-type regular`, layers.py:219-303) incl. the trailing ELU's backward."""
    torch.manual_seed(cin * 7 + cout)
    blk = L.FixupResBlock(cin, cout, mode)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.2))
        blk.scale.fill_(0.9)
    _grads_vs_oracle(blk, O.fixup_block, mode, torch.randn(shape), 5)


@pytest.mark.parametrize("cin,cout,mode,shape", [
    (8, 8, "same", (1, 8, 4, 5, 4)),        # one group, no skip conv
    (16, 8, "same", (1, 16, 3, 4, 4)),      # two groups, skip conv
    (8, 16, "down", (1, 8, 4, 4, 6)),
    (16, 8, "up", (1, 16, 2, 3, 2)),
])
def test_evonorm_block_gradients_vs_oracle_autograd(cin, cout, mode, shape):
    """EvonormResBlock / EvoNorm3DS0 backward (evonorm.py:12-47,59-76 under autograd)."""
    torch.manual_seed(cin * 3 + cout)
    blk = L.EvonormResBlock(cin, cout, mode)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.endswith(".gamma"):
                p.copy_(1.0 + 0.2 * torch.randn(p.shape))       # reference init gamma = 0 would zero every branch gradient
            elif n.endswith(".v") or n.endswith(".beta"):
                p.add_(0.2 * torch.randn(p.shape))
            else:
                p.copy_(torch.randn(p.shape) * 0.3)
    _grads_vs_oracle(blk, O.evonorm_block, mode, torch.randn(shape), 9, tol=2e-3)


@pytest.mark.parametrize("C,shape", [(4, (1, 4, 5, 6, 7)), (18, (2, 18, 4, 5, 3)), (2, (1, 2, 8, 8, 2))])
def test_fused_same_block_backward_on_the_emulator(C, shape):
    """vq3d_preact_same_backward (fused forward keeping only x + two tiled recompute kernels + tiled d W2; off by default) against
    autograd on the oracle: partial tiles, batch 2, wrap on a size-2 axis."""
    from vqvae import _ops
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
    with use_emulator():
        o = _ops.default()
        prev, o.fused_block_bwd = o.fused_block_bwd, True
        try:
            l0 = o.launches
            xg = x.clone().requires_grad_(True)
            y = blk(xg)
            (y * r).sum().backward()
            assert o.launches - l0 == 4                      # 1 forward + 3 backward launches
        finally:
            o.fused_block_bwd = prev
    assert torch.allclose(y.detach(), yr.detach(), rtol=1e-4, atol=1e-5)
    assert torch.allclose(xg.grad, xr.grad, rtol=2e-4, atol=2e-5)
    for k, p in blk.named_parameters():
        ref = sd["b." + k].grad
        assert torch.allclose(p.grad, ref, rtol=1e-3, atol=1e-4), (k, float((p.grad - ref).abs().max()))


@pytest.mark.parametrize("mode,cin,cout,shape", [("same", 6, 3, (1, 6, 4, 3, 5)), ("down", 4, 8, (1, 4, 6, 4, 8)), ("up", 8, 4, (1, 8, 3, 4, 2))])
def test_parameter_gradients_accumulate_in_place_like_autograd(mode, cin, cout, shape):
    """Ops.grad_inplace: with existing contiguous fp32 `.grad` buffers (the flat buffer of FusedAdamAMSGrad) the kernels add the
    weight / bias gradients straight into them and the Fixup scalars arrive through one multi-tensor add -- the same values as
    autograd's AccumulateGrad path, accumulated on top of what the buffers held, over two backward passes."""
    from vqvae import _ops
    torch.manual_seed(cin + cout)
    blk = L.PreActFixupResBlock(cin, cout, mode)
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn(p.shape) * 0.2)
    x = torch.randn(shape)
    grads = {}
    with use_emulator():
        o = _ops.default()
        for inplace in (False, True):
            prev, o.grad_inplace = o.grad_inplace, inplace
            try:
                for p in blk.parameters():
                    p.grad = torch.full_like(p, 0.25)              # pre-existing content must be kept
                for _ in range(2):
                    xi = x.clone().requires_grad_(True)
                    y = blk(xi)
                    (y * y).sum().backward()
                grads[inplace] = [p.grad.clone() for p in blk.parameters()] + [xi.grad.clone()]
            finally:
                o.grad_inplace = prev
    for a, b in zip(grads[False], grads[True]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), float((a - b).abs().max())
