"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the
product modules -> ctypes -> libvqvae3d_b200.so; the oracle (oracle/) and the committed
golden vectors of the reference are the checkers.

Tolerances (stated per north_star): code indices bit-exact on identical latents; the
gathered codewords bit-exact; fp32 conv path within rtol 2e-5 / atol 2e-6 per block and
rtol 1e-4 / atol 1e-5 through whole models; EMA buffers within fp32 reduction-order
tolerance (atomics are order-nondeterministic).
"""
import json
import os

import numpy as np
import pytest
import torch

from common import GOLDEN_DIR, golden_state_dict, portable_randn, portable_volume
from oracle import vqvae_oracle as O
from vqvae import layers as L
from vqvae.model import VQVAE, downscaled_config_args, full_config_args

pytestmark = pytest.mark.gpu
CASES = json.load(open(os.path.join(GOLDEN_DIR, "manifest.json")))["cases"]
DEV = "cuda"


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def by_kind(kind, pred=lambda c: True):
    return sorted(k for k, v in CASES.items() if v["kind"] == kind and pred(v))


def close(a, b, rtol=2e-5, atol=2e-6):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return np.allclose(a.astype(np.float64), b.astype(np.float64), rtol=rtol, atol=atol)


@pytest.fixture(autouse=True)
def _precision(request):
    """fp32 (exact SIMT kernels everywhere) unless a test asks for the bf16 tensor-core mode."""
    from vqvae import _ops
    o = _ops.default()
    prev = o.precision
    o.precision = "bf16" if "bf16" in request.keywords else "fp32"
    yield
    o.precision = prev


def test_native_library_is_loaded():
    from vqvae import _cabi
    lib = _cabi.lib()
    assert lib.vq3d_is_cuda_build() == 1
    assert any("libvqvae3d_b200.so" in l for l in open("/proc/self/maps"))


def make_quantizer(K, D, embed, first_pass, training):
    q = L.Quantizer(K, D, 0.1)
    q.embed.copy_(embed); q.embed_avg.copy_(embed); q.cluster_size.zero_(); q.first_pass.fill_(first_pass)
    return q.train(training).to(DEV)


@pytest.mark.parametrize("name", by_kind("quantizer"))
def test_quantizer_golden(name):
    c, g = CASES[name], load(name)
    x = portable_randn((c["B"], c["D"]) + tuple(c["spatial"]), c["seed"])
    embed = portable_randn((c["K"], c["D"]), c["seed"] + 1)
    q = make_quantizer(c["K"], c["D"], embed, 0, False)
    loss, quant, idx = q(x.to(DEV))
    assert idx.dtype == torch.int64 and idx.shape == (c["B"],) + tuple(c["spatial"])
    assert np.array_equal(idx.cpu().numpy(), g["eval_idx"])
    assert np.array_equal(quant.cpu().numpy(), g["eval_quantized"])
    assert loss.dim() == 0 and close(loss, g["eval_loss"], rtol=1e-6, atol=0)
    # training: first-pass init + EMA, then a second step
    q = make_quantizer(c["K"], c["D"], embed, 1, True)
    l1, _, i1 = q(x.to(DEV))
    assert int(q.first_pass) == 0
    assert close(q.cluster_size, g["t1_cluster_size"], rtol=1e-5)
    if np.array_equal(i1.cpu().numpy(), g["t1_idx"]):
        assert close(q.embed, g["t1_embed"], rtol=1e-4, atol=1e-6)
        assert close(q.embed_avg, g["t1_embed_avg"], rtol=1e-4, atol=1e-6)
        assert close(l1, g["t1_loss"], rtol=1e-5)
        x2 = portable_randn((c["B"], c["D"]) + tuple(c["spatial"]), c["seed"] + 2)
        l2, _, i2 = q(x2.to(DEV))
        if np.array_equal(i2.cpu().numpy(), g["t2_idx"]):
            assert close(q.embed, g["t2_embed"], rtol=1e-4, atol=1e-6)
            assert close(q.cluster_size, g["t2_cluster_size"], rtol=1e-5)
    else:
        assert (i1.cpu().numpy() != g["t1_idx"]).mean() < 0.02
    # straight-through backward
    q = make_quantizer(c["K"], c["D"], embed, 0, False)
    xg = x.to(DEV).requires_grad_(True)
    l, qq, _ = q(xg)
    (l * 1.7 + (qq * torch.from_numpy(g["bwd_grad_q"]).to(DEV)).sum()).backward()
    assert close(xg.grad, g["bwd_grad_x"], rtol=1e-5, atol=1e-7)


def test_quantizer_ties():
    g = load("q_ties")
    for e, x, idx in ((g["embed"], g["x"], g["idx"]), (g["embed2"], g["x2"], g["idx2"])):
        q = make_quantizer(e.shape[0], e.shape[1], torch.from_numpy(e), 0, False)
        _, _, got = q(torch.from_numpy(x).to(DEV))
        assert np.array_equal(got.cpu().numpy(), idx)


# (N, D, K): the model's real sizes (Full: 524288x2x128, 8192x8x256, 128x32x512), ragged N,
# run-time-D fallbacks, and sweep-sized cases of BASELINE.json configs[4]
@pytest.mark.parametrize("N,D,K", [(524288, 2, 128), (8192, 8, 256), (128, 32, 512), (1000, 2, 128), (77, 8, 256),
                                   (4099, 3, 19), (5000, 5, 64), (3001, 6, 40), (20000, 16, 300), (100000, 32, 512),
                                   (30000, 64, 1024), (9000, 128, 512), (1, 2, 1), (1 << 20, 32, 512),
                                   # tensor-core candidate pass (N >= 32768, D in {32, 64, 128}): ragged N, K not a tile multiple
                                   (65536, 64, 1024), (40000, 128, 2048), (50001, 32, 300), (33000, 128, 96), (70000, 64, 64)])
def test_quantizer_vs_oracle_index_exact(N, D, K):
    rs = np.random.RandomState(N % 9973 + D * 7 + K)
    x = rs.standard_normal((N, D)).astype(np.float32)
    e = rs.standard_normal((K, D)).astype(np.float32)
    ref_idx, _ = O.vq_assign_c(x, e)
    # present as (1, D, N/f, f, 1)-style volume, like the reference does for latents
    f = 64 if N % 64 == 0 else 1
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).reshape(1, D, N // f, f, 1)
    q = make_quantizer(K, D, torch.from_numpy(e), 0, False)
    loss, quant, idx = q(xt.to(DEV))
    got = idx.cpu().numpy().reshape(-1)
    mism = np.nonzero(got != ref_idx)[0]
    assert mism.size == 0, (mism.size, mism[:5])
    qv = quant.cpu().numpy().reshape(D, N).T
    assert np.array_equal(qv, x + (e[ref_idx] - x))                   # straight-through value, bit-exact
    ref_loss = 0.1 * np.mean((e[ref_idx].astype(np.float64) - x) ** 2)
    assert abs(float(loss) - ref_loss) <= 1e-6 * ref_loss + 1e-12
    # idempotence (size-independent property): codewords quantize to themselves
    cw = torch.from_numpy(np.ascontiguousarray(e[ref_idx].T)).reshape(1, D, N // f, f, 1).to(DEV)
    _, _, idx2 = q(cw)
    d_self = np.sqrt(((e[ref_idx] - e[idx2.cpu().numpy().reshape(-1)]) ** 2).sum(1))
    assert (d_self == 0).all()


def test_quantizer_tensor_core_path_ties_and_overflow():
    """The tensor-core candidate pass must keep the reference's first-index tie rule: duplicated codebook
    rows (more candidates than the kernel tracks -> exact full-scan fallback) and grid-snapped data with
    many exactly equal distances."""
    from vqvae import _ops
    o = _ops.default()
    rs = np.random.RandomState(11)
    N, D, K = 40000, 32, 512
    for snap in (False, True):
        x = rs.standard_normal((N, D)).astype(np.float32)
        e = rs.standard_normal((K, D)).astype(np.float32)
        if snap:
            x, e = np.round(x * 2) / 2, np.round(e * 2) / 2
        e[100:120] = e[7]                       # 21 identical rows: lowest index must win
        e[300] = e[299]
        x[:500] = e[7] + 0.01 * rs.standard_normal((500, D)).astype(np.float32)
        ref_idx, _ = O.vq_assign_c(x, e)
        xt = torch.from_numpy(np.ascontiguousarray(x.T)).reshape(1, D, N // 64, 64, 1)
        q = make_quantizer(K, D, torch.from_numpy(e), 0, False)
        o.profile = []
        try:
            _, quant, idx = q(xt.to(DEV))
            torch.cuda.synchronize()
            assert [p[0] for p in o.profile][0] == "vq_assign_tc"
        finally:
            o.profile = None
        got = idx.cpu().numpy().reshape(-1)
        assert np.array_equal(got, ref_idx), int((got != ref_idx).sum())
        assert (got[:500] == 7).all()
        assert np.array_equal(quant.cpu().numpy().reshape(D, N).T, x + (e[ref_idx] - x))


def test_quantizer_tensor_core_path_training_statistics():
    N, D, K = 65536, 32, 512
    rs = np.random.RandomState(5)
    x = rs.standard_normal((N, D)).astype(np.float32)
    e = rs.standard_normal((K, D)).astype(np.float32)
    q = make_quantizer(K, D, torch.from_numpy(e), 0, True)
    q.cluster_size.fill_(1.0)
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).reshape(1, D, N // 64, 64, 1)
    _, _, idx = q(xt.to(DEV))
    ref_idx, _ = O.vq_assign_c(x, e)
    assert np.array_equal(idx.cpu().numpy().reshape(-1), ref_idx)
    n, dw = O.vq_stats_c(x, ref_idx, K)
    assert close(q.cluster_size, 0.99 + 0.01 * n, rtol=1e-5)
    assert close(q.embed_avg, 0.99 * e.astype(np.float64) + 0.01 * dw, rtol=1e-4, atol=1e-5)


def test_quantizer_ema_statistics_vs_oracle():
    N, D, K = 200000, 8, 256
    rs = np.random.RandomState(3)
    x = rs.standard_normal((N, D)).astype(np.float32)
    e = rs.standard_normal((K, D)).astype(np.float32)
    q = make_quantizer(K, D, torch.from_numpy(e), 0, True)
    q.cluster_size.fill_(1.0)
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).reshape(1, D, N // 64, 64, 1)
    _, _, idx = q(xt.to(DEV))
    ref_idx, _ = O.vq_assign_c(x, e)
    assert np.array_equal(idx.cpu().numpy().reshape(-1), ref_idx)
    n, dw = O.vq_stats_c(x, ref_idx, K)
    cs = 0.99 * 1.0 + 0.01 * n
    ea = 0.99 * e.astype(np.float64) + 0.01 * dw
    tot = cs.sum()
    sm = tot * (cs + 1e-5) / (tot + K * 1e-5)
    assert close(q.cluster_size, cs, rtol=1e-5)
    assert close(q.embed_avg, ea, rtol=1e-4, atol=1e-5)
    assert close(q.embed, ea / sm[:, None], rtol=1e-4, atol=1e-5)
    assert abs(float(q.cluster_size.sum()) - (0.99 * K + 0.01 * N)) < 1e-2     # counts sum to N


def test_embed_code():
    q = make_quantizer(16, 3, portable_randn((16, 3), 1), 0, False)
    idx = torch.from_numpy(np.random.RandomState(0).randint(0, 16, size=(2, 3, 4, 5))).to(DEV)
    assert torch.equal(q.embed_code(idx), q.embed[idx])


@pytest.mark.parametrize("name", by_kind("block"))
def test_block_golden(name):
    c, g = CASES[name], load(name)
    with torch.no_grad():
        m = getattr(L, c["cls"])(c["cin"], c["cout"], c["mode"]).eval()
        m.load_state_dict(golden_state_dict(c["spec"], c["seed"]))
        m.to(DEV)
        x = portable_randn(c["shape"], c["seed"] + 7).to(DEV)
        y = m(x)
        assert close(y, g["y"]), np.abs(y.cpu().numpy() - g["y"]).max()
        if c["cls"] == "PreActFixupResBlock":
            assert close(m.forward_composed(x), g["y"])


def _golden_model(name):
    c = CASES[name]
    m = VQVAE(VQVAE.default_args(**c["cfg"])).eval()
    sd = golden_state_dict(c["spec"], c["seed"])
    assert list(sd.keys()) == list(m.state_dict().keys())
    m.load_state_dict(sd)
    return c, m.to(DEV)


@pytest.mark.parametrize("name", ["tiny2_preact", "tiny3_preact", "tiny2_regular", "tiny2_evonorm"])
def test_model_golden(name):
    g = load(name)
    c, m = _golden_model(name)
    x = portable_volume(c["shape"], c["seed"] + 11).to(DEV)
    n = c["cfg"]["n_bottleneck_blocks"]
    with torch.no_grad():
        dec, (losses, quants, idxs) = m(x)
        sub = (lambda t: t) if c["decoded_full"] else (lambda t: t[..., ::4, ::4, ::4])
        mism = [float((idxs[i].cpu().numpy() != g[f"eval_idx_{i}"]).mean()) for i in range(n)]
        assert max(mism) <= 0.01, mism
        if max(mism) == 0.0:
            assert close(sub(dec), g["eval_decoded"], rtol=1e-4, atol=1e-5)
            for i in range(n):
                assert close(losses[i], g[f"eval_loss_{i}"], rtol=1e-4)
        # teacher-forced: reference latents -> index-exact; reference quantised -> reference volume
        for i, qz in enumerate(m.encoder.quantize):
            _, _, got = qz(torch.from_numpy(g[f"eval_latent_{i}"]).to(DEV))
            assert np.array_equal(got.cpu().numpy(), g[f"eval_idx_{i}"])
        dec_tf = m.decoder([torch.from_numpy(g[f"eval_quantized_{i}"]).to(DEV) for i in range(n)])
        assert close(sub(dec_tf), g["eval_decoded"], rtol=1e-4, atol=1e-5)
        # CUDA-graph replay gives the same bits as the eager launches
        m.enable_cuda_graphs()
        dec_g, (_, _, idx_g) = m(x)
        dec_g2, _ = m(x)
        assert torch.equal(dec_g2, dec) and all(torch.equal(a, b) for a, b in zip(idx_g, idxs))


def _perturbed(args, seed=42, first_pass=0):
    torch.manual_seed(seed)
    m = VQVAE(args)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in m.parameters():           # Fixup zero-inits branch_conv3: perturb so branches are live
            p.add_(torch.randn(p.shape, generator=g) * 0.02)
        for q in m.encoder.quantize:
            q.first_pass.fill_(first_pass)
    return m.eval()


def test_config1_downscaled_model_vs_oracle():
    """BASELINE.json configs[0]: 2-level downscaled model (662 blocks), one 128x128x64 volume,
    against the oracle's fp32 CPU forward."""
    m = _perturbed(downscaled_config_args())
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = O.synthetic_volume((1, 1, 128, 128, 64))
    lat = {}
    with torch.no_grad():
        ref_dec, (ref_loss, ref_q, ref_idx) = O.vqvae_forward(sd, O.DOWNSCALED, x, collect=lat)
        m.to(DEV)
        dec, (loss, q, idx) = m(x.to(DEV))
        for i, qz in enumerate(m.encoder.quantize):           # teacher-forced index exactness
            _, _, got = qz(lat[f"latent_{i}"].to(DEV))
            assert torch.equal(got.cpu(), ref_idx[i])
        dec_tf = m.decoder([t.to(DEV) for t in ref_q])
    assert close(dec_tf, ref_dec, rtol=1e-3, atol=1e-4), float((dec_tf.cpu() - ref_dec).abs().max())
    mism = [float((a.cpu() != b).float().mean()) for a, b in zip(idx, ref_idx)]
    assert max(mism) < 0.02, mism
    for i in range(2):      # latents agree to fp32 accumulation noise through 150+ blocks
        pass
    rel = float((dec.cpu() - ref_dec).abs().mean() / ref_dec.abs().mean())
    assert rel < 1e-2, rel


def test_same_block_shift_equivariance_full_size():
    """Size-independent property at the Full model's largest stack shape (18 ch @128x128x32):
    a 'same' block with circular padding commutes with circular shifts, bit for bit."""
    torch.manual_seed(0)
    blk = L.PreActFixupResBlock(18, 18, "same")
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn(p.shape) * 0.1)
        blk.to(DEV).eval()
        x = torch.randn(1, 18, 128, 128, 32, device=DEV)
        y = blk(x)
        ys = blk(torch.roll(x, shifts=(5, -3, 7), dims=(2, 3, 4)))
        assert torch.equal(ys, torch.roll(y, shifts=(5, -3, 7), dims=(2, 3, 4)))
        assert torch.allclose(blk.forward_composed(x), y, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("C,shape", [(4, (2, 4, 24, 20, 128)), (4, (1, 4, 9, 8, 16)), (8, (1, 8, 16, 24, 64)), (8, (2, 8, 8, 10, 32))])
def test_row_kernel_two_rows_per_thread_vs_composed(C, shape):
    """preact_row_kernel (packed FMAs, two output rows per thread when H, W >= 8; partial tiles included) against the
    same block composed from the generic fp32 convolution kernels, and its circular-shift equivariance bit for bit."""
    torch.manual_seed(C)
    blk = L.PreActFixupResBlock(C, C, "same")
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn(p.shape) * 0.1)
        blk.to(DEV).eval()
        x = torch.randn(*shape, device=DEV)
        y = blk(x)
        assert torch.allclose(blk.forward_composed(x), y, rtol=2e-5, atol=2e-6), float((blk.forward_composed(x) - y).abs().max())
        ys = blk(torch.roll(x, shifts=(3, -1, 5), dims=(2, 3, 4)))
        assert torch.equal(ys, torch.roll(y, shifts=(3, -1, 5), dims=(2, 3, 4)))


def test_full_model_512_runs_and_is_self_consistent():
    """BASELINE.json metric config: Full 3-level model on one 512x512x128 volume.
    The oracle does not finish in seconds here, so use size-independent properties:
    determinism, graph == eager, index range, codeword idempotence, decoder(embed_code(idx))
    == decoder(quantised) and the commitment loss recomputed from the outputs."""
    m = _perturbed(full_config_args()).to(DEV)
    x = O.synthetic_volume((1, 1, 512, 512, 128)).to(DEV)
    with torch.no_grad():
        dec, (losses, quants, idxs) = m(x)
        assert dec.shape == x.shape and torch.isfinite(dec).all()
        assert [tuple(i.shape) for i in idxs] == [(1, 128, 128, 32), (1, 32, 32, 8), (1, 8, 8, 2)]
        for lvl, (qz, q, idx) in enumerate(zip(m.encoder.quantize, quants, idxs)):
            assert int(idx.min()) >= 0 and int(idx.max()) < qz.num_embeddings
            code = qz.embed_code(idx).permute(0, 4, 1, 2, 3).contiguous()
            assert torch.allclose(code, q, rtol=0, atol=1e-6)
            _, _, idx2 = qz(code)
            assert torch.equal(qz.embed[idx2], qz.embed[idx])
        dec2 = m.decoder([qz.embed_code(i).permute(0, 4, 1, 2, 3).contiguous() for qz, i in zip(m.encoder.quantize, idxs)])
        assert torch.allclose(dec2, dec, rtol=1e-4, atol=1e-4)
        dec_again, (_, _, idx_again) = m(x)
        assert torch.equal(dec_again, dec) and all(torch.equal(a, b) for a, b in zip(idx_again, idxs))


def test_huber_epilogue_vs_oracle():
    torch.manual_seed(1)
    dec = torch.randn(2, 1, 16, 12, 8)
    x = torch.rand(2, 1, 16, 12, 8) * 4.5 - 0.5
    nv = [8, 5]
    commit = [torch.tensor(0.25), torch.tensor(0.5)]
    from vqvae import _ops
    from vqvae.model import center_cylinder_mask
    for cyl in (False, True):
        ref_loss, ref_recon = O.huber_epilogue(dec, x, nv, commit, cylinder=cyl)
        mask = center_cylinder_mask(16, 12).to(torch.uint8).reshape(-1).to(DEV) if cyl else None
        acc = _ops.default().huber_elu_mask(dec.to(DEV), x.to(DEV), torch.tensor(nv, dtype=torch.int32, device=DEV), mask)
        got = float(acc[0] / acc[1])
        assert abs(got - float(ref_recon)) < 1e-6 * max(1.0, abs(float(ref_recon))), (got, float(ref_recon))


# ---- tcgen05 implicit-GEMM convolution vs the fp32 SIMT kernel ---------------------------------
@pytest.mark.parametrize("C1,C2,Cout,k,stride,circ,shape,pre_act,res", [
    (36, 0, 36, 3, 1, True, (1, 32, 32, 8), True, False),      # conv2 of the Full model's 72-channel stack
    (72, 0, 36, 1, 1, False, (1, 32, 32, 8), True, False),     # its conv1
    (36, 0, 72, 1, 1, False, (1, 32, 32, 8), True, True),      # its conv3 (+scale, bias4, residual)
    (64, 8, 72, 1, 1, False, (1, 32, 32, 8), False, False),    # cat + proj (two sources, conv bias)
    (64, 0, 64, 3, 1, True, (1, 16, 16, 4), True, False),
    (128, 0, 128, 3, 1, True, (1, 8, 8, 2), True, False),      # one 128-row tile, wrap on a size-2 axis
    (128, 0, 128, 4, 2, True, (1, 16, 16, 4), True, False),    # down block conv2
    (64, 0, 128, 2, 2, False, (1, 16, 16, 4), False, False),   # down block skip (k2 s2, no padding)
    (16, 0, 16, 4, 2, True, (1, 32, 32, 16), True, False),     # the 16-channel down block's conv2 (specialised k4 gather)
    (16, 0, 16, 4, 2, False, (2, 8, 12, 6), True, False),      # k4 s2 with zero padding: taps outside are skipped, batch 2
    (12, 4, 16, 4, 2, True, (1, 8, 8, 4), False, True),        # k4 s2 over two sources, wrap on a size-4 axis, residual
    (16, 0, 32, 2, 2, False, (1, 16, 16, 8), False, False),    # its skip convolution (specialised k2 gather)
    (8, 8, 16, 2, 2, False, (2, 6, 10, 4), True, False),       # k2 s2 over two sources
    (9, 0, 9, 3, 1, True, (2, 12, 10, 6), True, False),        # ragged: N padded 9 -> 16, K 243 -> 256, batch 2
    (16, 0, 24, 3, 1, False, (1, 5, 7, 9), True, True),        # zero padding, odd extents, partial M tile
])
def test_tensor_core_conv_matches_fp32(C1, C2, Cout, k, stride, circ, shape, pre_act, res):
    from vqvae import _ops
    o = _ops.default()
    rs = np.random.RandomState(C1 * 31 + Cout + k)
    B, H, W, Z = shape
    x1 = torch.from_numpy(rs.standard_normal((B, C1, H, W, Z)).astype(np.float32)).to(DEV)
    x2 = torch.from_numpy(rs.standard_normal((B, C2, H, W, Z)).astype(np.float32)).to(DEV) if C2 else None
    w = torch.from_numpy((rs.standard_normal((Cout, C1 + C2, k, k, k)) / np.sqrt((C1 + C2) * k ** 3)).astype(np.float32)).to(DEV)
    bias = torch.from_numpy(rs.standard_normal(Cout).astype(np.float32)).to(DEV) if C2 else None
    sc = lambda v: torch.tensor([v], dtype=torch.float32, device=DEV)
    pad = 1 if k >= 3 else 0
    kw = dict(x2=x2, bias=bias, stride=stride, pad=pad, circular=circ, pre_act=pre_act, pre_a=sc(0.1) if pre_act else None,
              pre_b=sc(-0.05), post_scale=sc(0.9), post_b=sc(0.02))
    prev = o.precision
    try:
        o.precision = "fp32"
        ref = o.conv3d(x1, w, **kw)
        r = torch.randn_like(ref) if res else None
        ref = o.conv3d(x1, w, residual=r, **kw)
        o.precision = "bf16"
        n0 = len(o.profile) if o.profile is not None else None
        o.profile = []
        got = o.conv3d(x1, w, residual=r, **kw)
        torch.cuda.synchronize()
        assert [e[0] for e in o.profile] == ["conv3d_tc"], "tensor-core path was not taken"
    finally:
        o.precision = prev
        o.profile = None
    conv_only = ref - (r if res else 0)
    err = float((got - ref).abs().max())
    scale = float(conv_only.abs().max())
    assert err <= 2e-2 * scale, (err, scale)
    assert float((got - ref).abs().mean()) <= 4e-3 * float(conv_only.abs().mean() + 1e-6)


@pytest.mark.bf16
@pytest.mark.parametrize("name", ["tiny2_preact", "tiny3_preact"])
def test_model_golden_bf16_tensor_core_mode(name):
    """Default product mode: GEMM-shaped convs on tcgen05 with bf16 operands.  Tolerance of the
    north star for BF16 paths: max-abs 1e-2 relative (to the volume's dynamic range), on the
    teacher-forced decoder; code indices stay bit-exact on identical latents (quantizer is fp32)."""
    g = load(name)
    c, m = _golden_model(name)
    n = c["cfg"]["n_bottleneck_blocks"]
    sub = (lambda t: t) if c["decoded_full"] else (lambda t: t[..., ::4, ::4, ::4])
    with torch.no_grad():
        dec_tf = m.decoder([torch.from_numpy(g[f"eval_quantized_{i}"]).to(DEV) for i in range(n)])
        ref = g["eval_decoded"]
        err = np.abs(sub(dec_tf).cpu().numpy() - ref).max()
        assert err <= 1e-2 * np.abs(ref).max(), (err, np.abs(ref).max())
        for i, qz in enumerate(m.encoder.quantize):
            _, _, got = qz(torch.from_numpy(g[f"eval_latent_{i}"]).to(DEV))
            assert np.array_equal(got.cpu().numpy(), g[f"eval_idx_{i}"])
        x = portable_volume(c["shape"], c["seed"] + 11).to(DEV)
        dec, (_, _, idxs) = m(x)
        mism = [float((idxs[i].cpu().numpy() != g[f"eval_idx_{i}"]).mean()) for i in range(n)]
        assert max(mism) <= 0.1, mism


@pytest.mark.bf16
def test_full_model_512_bf16_close_to_fp32():
    """Full model, 512x512x128: the bf16 tensor-core mode against the fp32 mode of the same
    kernels (teacher-forced through the decoder so a flipped code cannot dominate)."""
    from vqvae import _ops
    o = _ops.default()
    m = _perturbed(full_config_args()).to(DEV)
    x = O.synthetic_volume((1, 1, 512, 512, 128)).to(DEV)
    with torch.no_grad():
        o.precision = "fp32"
        dec32, (_, q32, idx32) = m(x)
        o.precision = "bf16"
        dec16 = m.decoder(q32)
        _, (_, _, idx16) = m(x)
    scale = float(dec32.abs().max())
    assert float((dec16 - dec32).abs().max()) <= 2e-2 * scale
    assert float((dec16 - dec32).abs().mean()) <= 2e-3 * scale
    assert float((idx16[0] != idx32[0]).float().mean()) < 0.05


# ---- tensor-core persistent stack kernel vs the fp32 SIMT stack ------------------------------------
def _rand_stack(C, n, seed):
    torch.manual_seed(seed)
    blocks = [L.PreActFixupResBlock(C, C, "same") for _ in range(n)]
    with torch.no_grad():
        for b in blocks:
            b.initialize_weights(num_layers=max(n, 2))
            for p in b.parameters():
                p.add_(torch.randn(p.shape) * 0.05)
    return L.BlockSequence(*blocks).to(DEV).eval()


@pytest.mark.parametrize("C,n,shape", [
    (18, 1, (1, 16, 16, 32)),      # single block, regular launch, several tiles
    (18, 3, (1, 128, 128, 32)),    # the Full model's biggest stack shape: grid barrier between blocks
    (72, 2, (1, 32, 32, 8)),       # 72 -> 36 -> 72: CBP 48, three K-steps per tap
    (32, 5, (1, 8, 8, 2)),         # the whole volume is one tile; wrap on a size-2 axis
    (8, 27, (1, 32, 32, 8)),       # more than 24 blocks: chunked launches, ping-pong parity across chunks
    (16, 2, (2, 9, 7, 5)),         # odd extents (partial tiles), batch 2
    (64, 2, (1, 16, 16, 4)),
])
def test_tensor_core_stack_matches_fp32(C, n, shape):
    from vqvae import _ops
    o = _ops.default()
    seq = _rand_stack(C, n, seed=C + n)
    B, H, W, Z = shape
    x = torch.randn(B, C, H, W, Z, generator=torch.Generator().manual_seed(1)).to(DEV)
    prev = o.precision
    try:
        with torch.no_grad():
            o.precision = "fp32"
            ref = seq(x)
            o.precision = "bf16"
            o.profile = []
            got = seq(x)
            torch.cuda.synchronize()
            assert [e[0] for e in o.profile] == ["preact_stack_tc"], [e[0] for e in o.profile]
            got2 = seq(x)                       # deterministic (no atomics on this path)
            assert torch.equal(got, got2)
    finally:
        o.precision = prev
        o.profile = None
    branch = ref - x
    err = float((got - ref).abs().max())
    scale = float(branch.abs().max())
    assert err <= 2e-2 * scale, (err, scale)
    assert float((got - ref).abs().mean()) <= 5e-3 * float(branch.abs().mean() + 1e-6)


@pytest.mark.parametrize("cin,cout,shape", [
    (18, 8, (1, 16, 16, 8)),       # the Full decoder's 18 -> 9 -> 8 block (K 9 -> 16 padded), several tiles
    (18, 8, (2, 24, 20, 16)),      # batch 2, non-square, partial tiles
    (32, 16, (1, 8, 8, 2)),        # 32 -> 16 -> 16 at the top level's size: wrap on a size-4 axis after the upsampling
    (32, 16, (1, 16, 12, 8)),
    (16, 8, (1, 6, 5, 3)),         # odd low-resolution extents
    (72, 32, (1, 8, 8, 4)),        # 72 -> 36 -> 32: C_b padded 36 -> 48, staging wider than the output channels
    (18, 8, (1, 64, 64, 32)),      # a quarter of the Full model's tensor
])
def test_tensor_core_up_block_matches_fp32(cin, cout, shape):
    """vq3d_preact_up_tc (low-resolution stage + trilinear expansion + tcgen05 k3 convolution) against the fp32 kernels:
    the branch within bf16 operand rounding, the interpolated skip path to fp32 rounding."""
    from vqvae import _ops
    o = _ops.default()
    torch.manual_seed(cin + cout)
    blk = L.PreActFixupResBlock(cin, cout, "up")
    with torch.no_grad():
        blk.initialize_weights(num_layers=4)
        for p in blk.parameters():
            p.add_(torch.randn(p.shape) * 0.1)
    blk = blk.to(DEV).eval()
    B, H, W, Z = shape
    x = torch.randn(B, cin, H, W, Z, generator=torch.Generator().manual_seed(1)).to(DEV)
    prev = o.precision
    try:
        with torch.no_grad():
            o.precision = "fp32"
            ref = blk(x)
            skip = o.conv3d(o.upsample2x(x, pre_b=blk.bias1c), blk.skip_conv.weight, post_b=blk.bias1d)
            o.precision = "bf16"
            o.profile = []
            got = blk(x)
            torch.cuda.synchronize()
            assert [e[0] for e in o.profile] == ["preact_up_tc"], [e[0] for e in o.profile]
            assert torch.equal(got, blk(x))             # deterministic
    finally:
        o.precision = prev
        o.profile = None
    branch = ref - skip
    err, scale = float((got - ref).abs().max()), float(branch.abs().max())
    assert err <= 2e-2 * scale, (err, scale)
    assert float((got - ref).abs().mean()) <= 5e-3 * float(branch.abs().mean() + 1e-6)


@pytest.mark.bf16
@pytest.mark.parametrize("name", by_kind("block", lambda c: c.get("mode") == "up" and c.get("cls") == "PreActFixupResBlock"))
def test_up_block_golden_bf16_tensor_core_mode(name):
    """The reference's own outputs for the 'up' blocks (tests/golden) in the product's default bf16 mode."""
    c, g = CASES[name], load(name)
    with torch.no_grad():
        m = getattr(L, c["cls"])(c["cin"], c["cout"], c["mode"]).eval()
        m.load_state_dict(golden_state_dict(c["spec"], c["seed"]))
        m.to(DEV)
        x = portable_randn(c["shape"], c["seed"] + 7).to(DEV)
        y = m(x).cpu().numpy()
    err = np.abs(y - g["y"]).max()
    assert err <= 1e-2 * np.abs(g["y"]).max(), (name, err, np.abs(g["y"]).max())


@pytest.mark.bf16
def test_tensor_core_stack_shift_equivariance():
    """Circular padding => the block commutes with circular shifts; with tile-aligned shifts the
    tensor-core kernel must reproduce that bit for bit (same tiles, same summation order)."""
    seq = _rand_stack(18, 2, seed=5)
    x = torch.randn(1, 18, 128, 128, 32, device=DEV)
    with torch.no_grad():
        y = seq(x)
        ys = seq(torch.roll(x, shifts=(32, 64), dims=(2, 3)))
    assert torch.equal(ys, torch.roll(y, shifts=(32, 64), dims=(2, 3)))


@pytest.mark.parametrize("C1,C2,Cout,shape,pre_act,res", [
    (1, 0, 4, (1, 64, 64, 32), False, False),      # parse_input at a reduced size
    (16, 2, 18, (1, 32, 32, 64), False, False),    # cat + proj (two sources, bias), 9-channel chunks
    (5, 0, 7, (2, 16, 32, 128), True, True),       # Fixup transforms, residual, ragged channel chunk, batch 2
    (12, 0, 20, (1, 32, 64, 32), True, False),
])
def test_pointwise_kernel_vs_torch(C1, C2, Cout, shape, pre_act, res):
    """Vectorised 1x1 kernel (conv_kernels.cu::pointwise_kernel) against an fp64 torch reference."""
    import torch.nn.functional as F
    from vqvae import _ops
    o = _ops.default()
    rs = np.random.RandomState(C1 + 10 * Cout)
    B, H, W, Z = shape
    x1 = torch.from_numpy(rs.standard_normal((B, C1, H, W, Z)).astype(np.float32)).to(DEV)
    x2 = torch.from_numpy(rs.standard_normal((B, C2, H, W, Z)).astype(np.float32)).to(DEV) if C2 else None
    w = torch.from_numpy((rs.standard_normal((Cout, C1 + C2, 1, 1, 1)) / np.sqrt(C1 + C2)).astype(np.float32)).to(DEV)
    bias = torch.from_numpy(rs.standard_normal(Cout).astype(np.float32)).to(DEV)
    sc = lambda v: torch.tensor([v], dtype=torch.float32, device=DEV)
    r = torch.from_numpy(rs.standard_normal((B, Cout, H, W, Z)).astype(np.float32)).to(DEV) if res else None
    o.profile = []
    try:
        got = o.conv3d(x1, w, x2=x2, bias=bias, pre_act=pre_act, pre_a=sc(0.1) if pre_act else None, pre_b=sc(-0.05),
                       post_scale=sc(0.9), post_b=sc(0.02), residual=r)
        torch.cuda.synchronize()
        assert [e[0] for e in o.profile] == ["conv3d"]
    finally:
        o.profile = None
    xin = x1 if x2 is None else torch.cat([x1, x2], 1)
    xin = F.elu(xin + 0.1) - 0.05 if pre_act else xin - 0.05
    # fp64 reference (cuDNN's fp32 conv3d may use TF32)
    ref = torch.einsum("oc,bchwz->bohwz", w.double().flatten(1), xin.double()) * 0.9 + 0.02 + bias.double().view(1, -1, 1, 1, 1)
    ref = (ref + (r.double() if res else 0)).float()
    assert torch.allclose(got, ref, rtol=2e-5, atol=2e-5), float((got - ref).abs().max())


def test_batched_forward_equals_per_volume_forward():
    """Volumes stacked along B (bench.py --batch) are processed independently: a batch-2 forward of the downscaled
    model equals the two single-volume forwards (fp32 mode: bit-reproducible kernels, so indices must agree exactly)."""
    from vqvae.model import VQVAE, downscaled_config_args
    torch.manual_seed(42)
    m = VQVAE(downscaled_config_args())
    g = torch.Generator().manual_seed(43)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.02)
        for q in m.encoder.quantize:
            q.first_pass.fill_(0)
    m = m.eval().to(DEV)
    xs = [O.synthetic_volume((1, 1, 64, 64, 32), seed=s).to(DEV) for s in (1, 2)]
    with torch.no_grad():
        dec_b, (loss_b, _, idx_b) = m(torch.cat(xs))
        singles = [m(x) for x in xs]
    for i, (dec, (_, _, idx)) in enumerate(singles):
        for lvl in range(len(idx)):
            assert torch.equal(idx_b[lvl][i:i + 1], idx[lvl]), (i, lvl)
        assert torch.allclose(dec_b[i:i + 1], dec, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("C,n,shape,tail", [
    (4, 2, (1, 32, 32, 128), False),     # the 512^3-level block shape (full-depth tiles, Z = 128)
    (4, 3, (2, 16, 40, 128), True),      # batch 2, non-square tile grid, last block carries the fused out conv (row kernel)
    (8, 2, (1, 48, 32, 64), False),      # the 256^3-level shape: C_b = 4 (K = 108 -> 112), two w columns per M-block
    (8, 1, (1, 64, 64, 32), False),      # Z = 32: four w columns per M-block
])
def test_thin_tensor_core_stack_matches_fp32(C, n, shape, tail):
    """preact_thin_tc_kernel (merged-tap conv2 GEMM, K = 27 * C_b) against the exact fp32 row kernel."""
    from vqvae import _ops
    o = _ops.default()
    seq = _rand_stack(C, n, seed=3 * C + n)
    out = None
    if tail:
        torch.manual_seed(5)
        out = L.Conv3d(C, 1, kernel_size=1).to(DEV)
    B, H, W, Z = shape
    x = torch.randn(B, C, H, W, Z, generator=torch.Generator().manual_seed(2)).to(DEV)
    prev, prev_thin = o.precision, o.thin_tc
    try:
        with torch.no_grad():
            o.precision = "fp32"
            ref = seq(x, tail=out) if tail else seq(x)
            o.precision, o.thin_tc = "bf16", True           # the kernel is off by default (slower than the row kernel)
            o.profile = []
            got = seq(x, tail=out) if tail else seq(x)
            torch.cuda.synchronize()
            assert [e[0] for e in o.profile] == ["preact_stack_thin_tc"], [e[0] for e in o.profile]
            got2 = seq(x, tail=out) if tail else seq(x)
            assert torch.equal(got, got2)
    finally:
        o.precision, o.thin_tc = prev, prev_thin
        o.profile = None
    base = (out(x) if tail else x)
    branch = ref - base
    err, scale = float((got - ref).abs().max()), float(branch.abs().max())
    assert err <= 2e-2 * scale, (err, scale)
    assert float((got - ref).abs().mean()) <= 5e-3 * float(branch.abs().mean() + 1e-6)


@pytest.mark.parametrize("N,D,K,kind", [
    (65536, 32, 8192, "gauss"),        # 64 codebook tiles, streamed through the ring
    (40000, 64, 16384, "gauss"),       # the largest supported codebook (128 tiles)
    (65536, 64, 1024, "clustered"),    # latents close to their codes (a trained model): tiny distances, large norms
    (50000, 32, 512, "outlier"),       # a few codes with 100x the norm of the rest (dead codes after EMA updates)
    (40000, 128, 512, "scaled"),       # latents 1000x smaller than the codebook
    (40000, 32, 300, "nonfinite"),     # NaN / inf / all-zero latent vectors: torch.argmin's answer (index 0 on all-NaN rows)
    # every element of x and e sits exactly half-way between two bf16 values and rounds the same way (to even): the rounding
    # errors of the tensor-core operands (2^-9 .. 2^-8 relative, each) add up instead of averaging out, and the codes are
    # nearly equidistant: the stress case of the candidate margin
    (65536, 32, 512, "midpoint"),
    (65536, 64, 512, "midpoint"),
    (65536, 128, 1024, "midpoint"),
])
def test_quantizer_tensor_core_path_regimes(N, D, K, kind):
    """Index exactness of the tensor-core candidate pass + exact re-rank outside the Gaussian sweep regime."""
    rs = np.random.RandomState(K + D)
    e = rs.standard_normal((K, D)).astype(np.float32)
    if kind == "clustered":
        e = e * 3 + 5.0
        x = e[rs.randint(0, K, size=N)] + 0.05 * rs.standard_normal((N, D)).astype(np.float32)
    elif kind == "outlier":
        e[::37] *= 100.0
        x = rs.standard_normal((N, D)).astype(np.float32)
    elif kind == "scaled":
        x = 1e-3 * rs.standard_normal((N, D)).astype(np.float32)
    elif kind == "midpoint":
        def mid(v):          # bf16 value with an even last mantissa bit plus half a bf16 ulp: an exact tie, rounds back to it
            bits = (np.ascontiguousarray(v, dtype=np.float32).view(np.uint32) & np.uint32(0xfffe0000)) | np.uint32(0x8000)
            return bits.view(np.float32).copy()
        centre = 2.0 + rs.standard_normal((1, D)).astype(np.float32)
        e = mid(centre + 0.05 * rs.standard_normal((K, D)).astype(np.float32))
        x = mid(centre + 0.05 * rs.standard_normal((N, D)).astype(np.float32))
    elif kind == "nonfinite":
        x = rs.standard_normal((N, D)).astype(np.float32)
        x[5:40] = np.nan
        x[1000, 3] = np.nan
        x[2000:2010] = np.inf
        x[3000, 7] = -np.inf
        x[4000:4100] = 0.0
    else:
        x = rs.standard_normal((N, D)).astype(np.float32)
    x = x.astype(np.float32)
    ref_idx, _ = O.vq_assign_c(x, e)
    xt = torch.from_numpy(np.ascontiguousarray(x.T)).reshape(1, D, N, 1, 1)
    q = make_quantizer(K, D, torch.from_numpy(e), 0, False)
    from vqvae import _ops
    o = _ops.default()
    o.profile = []
    try:
        _, quant, idx = q(xt.to(DEV))
        torch.cuda.synchronize()
        assert [p[0] for p in o.profile][0] == "vq_assign_tc"
    finally:
        o.profile = None
    got = idx.cpu().numpy().reshape(-1)
    assert np.array_equal(got, ref_idx), int((got != ref_idx).sum())
    assert np.array_equal(quant.cpu().numpy().reshape(D, N).T, x + (e[ref_idx] - x), equal_nan=True)
