"""CPU-only logic checks of the product's kernel SOURCES, compiled for a host-thread SIMT
emulator (tests/emu), against the reference's golden vectors.  The GPU parity tests proper
are tests/test_gpu_*.py (-m gpu); these exist so that indexing / fusion bugs are caught
without GPU time."""
import json
import os

import numpy as np
import pytest
import torch

from common import GOLDEN_DIR, golden_state_dict, portable_randn, portable_volume
from emu.emu_ops import use_emulator
from vqvae import layers as L

CASES = json.load(open(os.path.join(GOLDEN_DIR, "manifest.json")))["cases"]


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def by_kind(kind, pred=lambda c: True):
    return sorted(k for k, v in CASES.items() if v["kind"] == kind and pred(v))


def close(a, b, rtol=2e-5, atol=2e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.allclose(a, b, rtol=rtol, atol=atol)


def make_quantizer(c, embed, first_pass, training):
    q = L.Quantizer(c["K"], c["D"], 0.1)
    q.embed.copy_(embed); q.embed_avg.copy_(embed); q.cluster_size.zero_(); q.first_pass.fill_(first_pass)
    return q.train(training)


@pytest.mark.parametrize("name", by_kind("quantizer"))
def test_emu_quantizer(name):
    c, g = CASES[name], load(name)
    x = portable_randn((c["B"], c["D"]) + tuple(c["spatial"]), c["seed"])
    embed = portable_randn((c["K"], c["D"]), c["seed"] + 1)
    with use_emulator():
        q = make_quantizer(c, embed, 0, False)
        loss, quant, idx = q(x)
        assert idx.dtype == torch.int64 and np.array_equal(idx.numpy(), g["eval_idx"])      # bit-exact indices
        assert np.array_equal(quant.numpy(), g["eval_quantized"])                            # exact gather
        assert close(loss.numpy(), g["eval_loss"], rtol=1e-6, atol=0)
        q = make_quantizer(c, embed, 1, True)
        l1, q1, i1 = q(x)
        assert int(q.first_pass) == 0
        assert close(q.cluster_size.numpy(), g["t1_cluster_size"], rtol=1e-5)
        same_idx = np.array_equal(i1.numpy(), g["t1_idx"])
        # the data-dependent init rescales the codebook in fp32; a last-bit difference in
        # mean/std may flip a near-tie, so indices are compared only when the init is identical
        if same_idx:
            assert close(q.embed.numpy(), g["t1_embed"], rtol=1e-4, atol=1e-6)
            assert close(q.embed_avg.numpy(), g["t1_embed_avg"], rtol=1e-4, atol=1e-6)
            assert close(l1.numpy(), g["t1_loss"], rtol=1e-5)
        else:
            assert (i1.numpy() != g["t1_idx"]).mean() < 0.02
        x2 = portable_randn((c["B"], c["D"]) + tuple(c["spatial"]), c["seed"] + 2)
        q(x2)
        assert close(q.cluster_size.numpy().sum(), g["t2_cluster_size"].sum(), rtol=1e-5)
        # straight-through backward
        q = make_quantizer(c, embed, 0, False)
        xg = x.clone().requires_grad_(True)
        l, qq, _ = q(xg)
        (l * 1.7 + (qq * torch.from_numpy(g["bwd_grad_q"])).sum()).backward()
        assert close(xg.grad.numpy(), g["bwd_grad_x"], rtol=1e-5, atol=1e-7)


def test_emu_quantizer_ties():
    g = load("q_ties")
    with use_emulator():
        for e, x, idx in ((g["embed"], g["x"], g["idx"]), (g["embed2"], g["x2"], g["idx2"])):
            q = L.Quantizer(e.shape[0], e.shape[1], 0.1).eval()
            q.embed.copy_(torch.from_numpy(e)); q.first_pass.fill_(0)
            _, _, got = q(torch.from_numpy(x))
            assert np.array_equal(got.numpy(), idx)


def test_emu_embed_code():
    with use_emulator():
        q = L.Quantizer(16, 3, 0.1)
        idx = torch.from_numpy(np.random.RandomState(0).randint(0, 16, size=(2, 3, 4, 5)))
        assert torch.equal(q.embed_code(idx), q.embed[idx])


@pytest.mark.parametrize("name", by_kind("block"))
def test_emu_block(name):
    c, g = CASES[name], load(name)
    with use_emulator(), torch.no_grad():
        m = getattr(L, c["cls"])(c["cin"], c["cout"], c["mode"]).eval()
        assert [k for k, _, _ in c["spec"]] == list(m.state_dict().keys())       # key names AND order
        m.load_state_dict(golden_state_dict(c["spec"], c["seed"]))
        x = portable_randn(c["shape"], c["seed"] + 7)
        y = m(x)
        assert close(y.numpy(), g["y"]), np.abs(y.numpy() - g["y"]).max()
        if c["cls"] == "PreActFixupResBlock":
            y2 = m.forward_composed(x)
            assert close(y2.numpy(), g["y"])


def _build_model(cfg):
    rb = {"regular": L.FixupResBlock, "pre-activation": L.PreActFixupResBlock, "evonorm": L.EvonormResBlock}[cfg["block_type"]]
    enc = L.Encoder2(in_channels=1, base_network_channels=cfg["base_network_channels"], n_enc=cfg["n_bottleneck_blocks"],
                     n_down_per_enc=2, n_pre_q_blocks=cfg["n_pre_quantization_blocks"],
                     n_post_downscale_blocks=cfg["n_post_downscale_blocks"], n_post_upscale_blocks=cfg["n_post_upscale_blocks"],
                     num_embeddings=cfg["num_embeddings"], resblock=rb)
    dec = L.Decoder(out_channels=1, base_network_channels=cfg["base_network_channels"], n_enc=cfg["n_bottleneck_blocks"],
                    n_up_per_enc=2, n_post_q_blocks=cfg["n_post_quantization_blocks"],
                    n_post_upscale_blocks=cfg["n_post_upscale_blocks"], resblock=rb)
    m = torch.nn.Module()
    m.encoder, m.decoder = enc, dec
    return m


@pytest.mark.parametrize("name", ["tiny2_preact", "tiny2_regular", "tiny2_evonorm"])
def test_emu_model(name):
    c, g = CASES[name], load(name)
    with use_emulator(), torch.no_grad():
        m = _build_model(c["cfg"]).eval()
        assert [k for k, _, _ in c["spec"]] == list(m.state_dict().keys())
        m.load_state_dict(golden_state_dict(c["spec"], c["seed"]))
        x = portable_volume(c["shape"], c["seed"] + 11)
        losses, quants, idxs = zip(*m.encoder(x))
        dec = m.decoder(quants)
        n = c["cfg"]["n_bottleneck_blocks"]
        mism = [float((idxs[i].numpy() != g[f"eval_idx_{i}"]).mean()) for i in range(n)]
        assert max(mism) <= 0.01, mism
        if max(mism) == 0.0:
            assert close(dec.numpy(), g["eval_decoded"], rtol=1e-4, atol=1e-5), np.abs(dec.numpy() - g["eval_decoded"]).max()
            for i in range(n):
                assert close(losses[i].numpy(), g[f"eval_loss_{i}"], rtol=1e-4)
        # teacher-forced decoder: the reference's own quantised latents -> reference decoded volume
        dec_tf = m.decoder([torch.from_numpy(g[f"eval_quantized_{i}"]) for i in range(n)])
        assert close(dec_tf.numpy(), g["eval_decoded"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("cin,cout,mode,shape", [
    (4, 4, "same", (1, 4, 20, 12, 40)),      # several tiles per axis, ragged in H and Z
    (18, 18, "same", (1, 18, 9, 5, 3)),      # odd extents smaller than a tile
    (18, 2, "same", (2, 18, 10, 9, 33)),
    (4, 8, "down", (1, 4, 20, 18, 68)),
    (8, 16, "down", (2, 8, 6, 4, 2)),
    (8, 4, "up", (1, 8, 10, 5, 17)),         # hi-res 20 x 10 x 34: ragged tiles + wrap on every axis
    (18, 8, "up", (1, 18, 3, 1, 2)),
    (4, 2, "up", (2, 4, 2, 6, 20)),
])
def test_emu_fused_equals_composed_multitile(cin, cout, mode, shape):
    torch.manual_seed(cin * 100 + cout)
    with use_emulator() as o, torch.no_grad():
        m = L.PreActFixupResBlock(cin, cout, mode).eval()
        for p in m.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.2))
        m.scale.fill_(0.9)
        x = torch.randn(shape)
        n0 = o.launches
        y = m(x)
        assert o.launches - n0 == 1, "expected the single fused launch"
        ref = m.forward_composed(x)
        assert close(y.numpy(), ref.numpy(), rtol=1e-4, atol=1e-5), float((y - ref).abs().max())


def test_emu_stack_ping_pong():
    torch.manual_seed(5)
    with use_emulator() as o, torch.no_grad():
        for n in (1, 2, 3):
            seq = L.BlockSequence(*(L.PreActFixupResBlock(4, 4, "same") for _ in range(n))).eval()
            for p in seq.parameters():
                p.copy_(torch.randn(p.shape) * 0.3)
            x = torch.randn(1, 4, 6, 5, 9)
            ref = x
            for blk in seq:
                ref = blk.forward_composed(ref)
            y = seq(x)
            assert close(y.numpy(), ref.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("c,shape,n,tail", [
    (4, (1, 4, 10, 12, 16), 1, False),     # row kernel: 4 z per thread, full-depth tiles, ragged in H
    (8, (2, 8, 5, 6, 8), 2, False),        # batch 2, two blocks (ping-pong)
    (2, (1, 2, 9, 7, 32), 3, False),
    (4, (1, 4, 6, 9, 4), 2, True),         # Z = 4: one thread per row holds both circular halos; fused `out` conv
    (4, (1, 4, 9, 10, 16), 1, True),
])
def test_emu_row_kernel(c, shape, n, tail):
    """preact_row_kernels.cu (the 512x512x128 / 256x256x64 layers) against the composed generic path."""
    torch.manual_seed(c * 7 + n)
    with use_emulator() as o, torch.no_grad():
        seq = L.BlockSequence(*(L.PreActFixupResBlock(c, c, "same") for _ in range(n))).eval()
        for p in seq.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.2))
        out = L.Conv3d(c, 1, kernel_size=1)
        x = torch.randn(shape)
        ref = x
        for blk in seq:
            ref = blk.forward_composed(ref)
        if tail:
            ref = out(ref)
        n0 = o.launches
        y = seq(x, tail=out) if tail else seq(x)
        assert o.launches - n0 == n, "expected one fused launch per block (trailing conv included)"
        assert y.shape == ref.shape
        assert close(y.numpy(), ref.numpy(), rtol=1e-4, atol=1e-5), float((y - ref).abs().max())


@pytest.mark.parametrize("cin,cout,shape,parse", [
    (4, 8, (1, 4, 8, 12, 16), False),      # down row kernel: 2x2 output-row tiles, wrap on every axis
    (4, 8, (2, 1, 6, 4, 8), True),         # fused parse_input (1 -> 4), batch 2, Z/8 = 1 output lane
    (8, 16, (1, 8, 4, 8, 32), False),
    (4, 8, (1, 1, 10, 6, 64), True),       # ragged number of tiles in H (5 output rows)
])
def test_emu_down_row_kernel(cin, cout, shape, parse):
    """preact_row_kernels.cu::preact_down_row_kernel (the encoder's 512^3 / 256^3 down blocks) vs the composed path."""
    torch.manual_seed(cin + cout + shape[-1])
    with use_emulator() as o, torch.no_grad():
        blk = L.PreActFixupResBlock(cin, cout, "down").eval()
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.2))
        pre = L.Conv3d(1, cin, kernel_size=1) if parse else None
        x = torch.randn(shape)
        ref = blk.forward_composed(pre(x) if parse else x)
        seq = L.BlockSequence(blk)
        n0 = o.launches
        y = seq(x, pre=pre) if parse else seq(x)
        assert o.launches - n0 == 1, "expected one fused launch"
        assert y.shape == ref.shape
        assert close(y.numpy(), ref.numpy(), rtol=1e-4, atol=1e-5), float((y - ref).abs().max())


@pytest.mark.parametrize("cin,cout,shape", [
    (8, 4, (1, 8, 4, 6, 8)),        # up row kernel: hi-res 8 x 12 x 16, several tiles, wrap + clamp on every axis
    (4, 2, (2, 4, 2, 3, 4)),        # smallest extents (tile rows clamp to the same low-res row), odd W, batch 2
    (8, 4, (1, 8, 3, 2, 8)),        # odd H
    (4, 2, (1, 4, 5, 4, 4)),        # ragged tiles in H (10 hi rows / 4)
])
def test_emu_up_row_kernel(cin, cout, shape):
    """preact_row_kernels.cu::preact_up_row_kernel (decoder up blocks at 256^3 -> 512^3) vs the composed path."""
    torch.manual_seed(cin + cout + shape[-1])
    with use_emulator() as o, torch.no_grad():
        blk = L.PreActFixupResBlock(cin, cout, "up").eval()
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape) * (0.3 if p.dim() > 1 else 0.2))
        x = torch.randn(shape)
        ref = blk.forward_composed(x)
        n0 = o.launches
        y = blk(x)
        assert o.launches - n0 == 1, "expected one fused launch"
        assert y.shape == ref.shape
        assert close(y.numpy(), ref.numpy(), rtol=1e-4, atol=1e-5), float((y - ref).abs().max())


@pytest.mark.parametrize("cin,cout,circ,shape,pre_act,res", [
    (2, 2, True, (1, 2, 24, 22, 33), True, False),        # partial tiles on every axis, circular wrap
    (5, 3, False, (1, 5, 21, 24, 35), True, True),        # zero padding, residual, CO_T = 4 with a ragged last chunk
    (9, 9, True, (1, 9, 10, 9, 40), True, False),         # the 9-channel output block
    (3, 4, True, (1, 3, 40, 36, 12), True, False),        # depth < 32: one partial z tile
    (4, 18, False, (2, 4, 9, 7, 33), False, True),        # two 9-channel blocks, batch 2, no pre-activation
])
def test_emu_tiled_conv_matches_torch(cin, cout, circ, shape, pre_act, res):
    """conv3d_tiled_kernel (k3 s1 p1, few input channels, >= 16384 output voxels) against torch's conv3d."""
    import torch.nn.functional as F
    from vqvae import _ops
    g = torch.Generator().manual_seed(cin * 5 + cout)
    x = torch.randn(shape, generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) * 0.2
    a, b, s, pb = (torch.randn(1, generator=g) * 0.3 for _ in range(4))
    r = torch.randn(shape[0], cout, *shape[2:], generator=g) if res else None
    u = (F.elu(x + a) + b) if pre_act else x + b
    u = F.pad(u, (1,) * 6, mode="circular") if circ else F.pad(u, (1,) * 6)
    ref = F.conv3d(u, w) * s + pb + (r if res else 0)
    with use_emulator():
        got = _ops.default().conv3d(x, w, pad=1, circular=circ, pre_act=pre_act, pre_a=a, pre_b=b, post_scale=s, post_b=pb, residual=r)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4), float((got - ref).abs().max())


@pytest.mark.parametrize("C,shape", [(4, (1, 4, 8, 10, 8)), (8, (2, 8, 9, 8, 8)), (2, (1, 2, 8, 8, 8))])
def test_emu_row_kernel_two_rows_per_thread(C, shape):
    """preact_row_kernel with H, W >= 8 (the two-output-rows-per-thread path of the even-width variants, partial
    tiles included) against the block composed from the generic convolution kernels."""
    with use_emulator(), torch.no_grad():
        torch.manual_seed(C)
        m = L.PreActFixupResBlock(C, C, "same").eval()
        for p in m.parameters():
            p.add_(torch.randn(p.shape) * 0.1)
        x = portable_randn(shape, 11 + C)
        assert close(m(x).numpy(), m.forward_composed(x).numpy())
