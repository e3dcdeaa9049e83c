"""GPU parity of the caller scripts through the product kernels: extract_embeddings' index stream against a
direct encode, decode_embeddings' Hounsfield volume against the oracle, and two steps of train.py's loop."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import vqvae_oracle as O
from test_callers import FakeEnv

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = dict(n_bottleneck_blocks=2, n_downscales_per_bottleneck=1, num_embeddings=[16, 24], n_pre_quantization_blocks=2,
           n_post_quantization_blocks=2, n_post_upscale_blocks=1, n_post_downscale_blocks=1)


@pytest.fixture(autouse=True)
def _fp32():
    from vqvae import _ops
    o = _ops.default()
    prev = o.precision
    o.precision = "fp32"
    yield
    o.precision = prev


def _model():
    from vqvae.model import VQVAE
    torch.manual_seed(42)
    m = VQVAE(VQVAE.default_args(extract_center_cylinder=False, **CFG))
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
        for q in m.encoder.quantize:
            q.first_pass.fill_(0)
    return m.eval()


def test_extract_embeddings_stream_and_database():
    from utils import open_dataset
    from vqvae.extract_embeddings import extract_samples, write_codes
    m = _model()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ds = open_dataset("synthetic:3:16x16x8")
    loader = torch.utils.data.DataLoader(ds, batch_size=1)
    env = FakeEnv(max_dbs=2)
    n = write_codes(env, m.n_bottleneck_blocks, m.num_embeddings, len(loader), extract_samples(m, loader, DEV))
    assert n == 3
    for i in range(3):
        x = ds[i][0][None]
        _, (_, _, ref_idx) = O.vqvae_forward(sd, O.ModelConfig(**CFG), x)
        for level in range(2):
            arr = pickle.loads(env.store[str(level).encode()][str(i).encode()])
            assert arr.dtype == np.int64 and arr.shape == tuple(ref_idx[level].shape)
            assert np.array_equal(arr, ref_idx[level].numpy()), (i, level)          # index-exact against the oracle


def test_decode_embeddings_hounsfield_volume():
    from vqvae.decode_embeddings import decode_codes
    m = _model()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    rs = np.random.RandomState(3)
    codes = [torch.from_numpy(rs.randint(0, 16, size=(8, 8, 4))), torch.from_numpy(rs.randint(0, 24, size=(4, 4, 2)))]
    got = decode_codes(m.to(DEV), codes).cpu().numpy()
    cfg = O.ModelConfig(**CFG)
    emb = [O.embed_code(sd, f"encoder.quantize.{l}.", c[None]).permute(0, 4, 1, 2, 3) for l, c in enumerate(codes)]
    dec = O.decoder_forward(sd, cfg, emb)
    ref = np.rint(torch.nn.functional.elu(dec).numpy() * 1000 - 1000).astype(np.int64)
    assert got.shape == ref.shape == (1, 1, 16, 16, 8) and got.dtype == np.int64
    assert np.abs(got - ref).max() <= 1                              # fp32 conv rounding can move a value across a .5 boundary
    assert (got != ref).mean() < 1e-2


def test_train_loop_two_steps_and_checkpoint(tmp_path):
    from vqvae import train
    from vqvae.model import VQVAE
    argv = ["synthetic:4:16x16x8", "--batch-size", "2", "--n-bottleneck-blocks", "2", "--n-downscales-per-bottleneck", "1",
            "--num-embeddings", "16", "24", "--n-pre-quantization-blocks", "1", "--n-post-quantization-blocks", "1",
            "--extract-center-cylinder", "False", "--base_lr", "1e-3", "--max-steps", "2", "--num-workers", "0",
            "--default-root-dir", str(tmp_path), "--log-every-n-steps", "1"]
    train.main(train.parse_arguments(argv))
    m = VQVAE.load_from_checkpoint(str(tmp_path / "checkpoints" / "last.ckpt"))
    assert all(torch.isfinite(p).all() for p in m.parameters())
    assert int(m.encoder.quantize[0].first_pass) == 0                 # the EMA init ran (layers.py:665-683)
    assert abs(float(m.encoder.quantize[0].cluster_size.sum()) - float(m.encoder.quantize[0].cluster_size.numel())) > 1e-3


def test_train_resume_from_checkpoint(tmp_path):
    """`--resume-from-checkpoint` (train_vqvae_3d.job:87): weights, EMA codebooks, Adam moments and the step counter come
    back (checked tensor by tensor on a fresh model + optimizer), a resumed run continues the step count, and the
    validation monitor writes best.ckpt (ModelCheckpoint(monitor='val_recon_loss_mean'), train.py:56)."""
    from vqvae import train
    from vqvae.model import VQVAE
    base = ["synthetic:4:16x16x8", "--batch-size", "1", "--n-bottleneck-blocks", "2", "--n-downscales-per-bottleneck", "1",
            "--num-embeddings", "16", "24", "--n-pre-quantization-blocks", "1", "--n-post-quantization-blocks", "1",
            "--extract-center-cylinder", "False", "--base_lr", "1e-3", "--num-workers", "0", "--log-every-n-steps", "1",
            "--default-root-dir", str(tmp_path)]
    args = train.parse_arguments(base + ["--max-steps", "2", "--val-dataset-path", "synthetic:2:16x16x8"])
    train.main(args)
    ck_path = tmp_path / "checkpoints" / "last.ckpt"
    assert (tmp_path / "checkpoints" / "best.ckpt").exists()
    half = torch.load(ck_path, weights_only=False)
    assert half["global_step"] == 2
    osd = half["optimizer_states"][0]
    assert all(int(st["step"]) == 2 for st in osd["state"].values())
    m = VQVAE(args).to(DEV).train()
    opt = m.configure_optimizers()
    train.load_checkpoint(m, opt, ck_path)
    assert opt.current_step() == 2
    for k, v in m.state_dict().items():
        assert torch.equal(v.cpu(), half["state_dict"][k]), k
    for i, p in enumerate(opt.param_groups[0]["params"]):
        for name in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
            assert torch.equal(opt.state[p][name].cpu(), osd["state"][i][name]), (i, name)
    assert float(opt._m.abs().sum()) > 0
    train.main(train.parse_arguments(base + ["--max-steps", "4", "--resume-from-checkpoint", str(ck_path)]))
    done = torch.load(ck_path, weights_only=False)
    assert done["global_step"] == 4
    assert all(int(st["step"]) == 4 for st in done["optimizer_states"][0]["state"].values())


def test_hounsfield_int16_in_and_out_against_the_oracle():
    """VQVAE.reconstruct_hu / encode_hu: raw int16 HU -> (device) clip, * 0.001f, + 1 (utils/load_nrrd_dataset.py:73-81) ->
    forward -> ELU * 1000 - 1000, rint (decode_embeddings.py:43-47) -> int16 HU, against the numpy front end + oracle forward."""
    from utils.volumes import preprocess_hu
    m = _model()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    rs = np.random.RandomState(5)
    hu = rs.randint(-2500, 4000, size=(2, 1, 16, 16, 8)).astype(np.int16)
    x = torch.from_numpy(preprocess_hu(hu))
    ref_dec, (_, _, ref_idx) = O.vqvae_forward(sd, O.ModelConfig(**CFG), x)
    ref_hu = np.clip(np.rint(torch.nn.functional.elu(ref_dec).numpy() * 1000 - 1000), -32768, 32767).astype(np.int16)
    m = m.to(DEV)
    got_hu, idx = m.reconstruct_hu(torch.from_numpy(hu).to(DEV))
    assert got_hu.dtype == torch.int16 and got_hu.shape == hu.shape
    for a, b in zip(idx, ref_idx):
        assert torch.equal(a.cpu(), b)
    diff = np.abs(got_hu.cpu().numpy().astype(np.int32) - ref_hu.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-2            # fp32 conv rounding can move a value across a .5 boundary
    for a, b in zip(m.encode_hu(torch.from_numpy(hu).to(DEV)), ref_idx):
        assert torch.equal(a.cpu(), b)
    from vqvae import _ops
    assert torch.equal(_ops.default().hu_to_network(torch.from_numpy(hu).to(DEV)).cpu(), x)


def test_decoder_tail_with_two_output_channels_falls_back():
    """input_channels = 2: the fused `out` tail only covers C -> 1; with two output channels the decoder must run
    the 1x1 on its own and return both channels, equal to the oracle."""
    from vqvae.model import VQVAE
    cfg = dict(CFG, input_channels=2)
    torch.manual_seed(3)
    m = VQVAE(VQVAE.default_args(extract_center_cylinder=False, **cfg))
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
        for q in m.encoder.quantize:
            q.first_pass.fill_(0)
    m.eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.rand(1, 2, 16, 16, 8, generator=torch.Generator().manual_seed(9)) * 4.5 - 0.5
    ref_dec, (ref_loss, _, ref_idx) = O.vqvae_forward(sd, O.ModelConfig(**cfg), x)
    m = m.to(DEV)
    with torch.no_grad():
        dec, (_, _, idx) = m(x.to(DEV))
    assert dec.shape == (1, 2, 16, 16, 8)
    assert all(torch.equal(a.cpu(), b) for a, b in zip(idx, ref_idx))
    assert torch.allclose(dec.cpu(), ref_dec, rtol=1e-4, atol=1e-5)
    # the Huber epilogue covers both channels (model.py:163)
    loss, log = m.huber((x.to(DEV), [6]))
    ref_total, ref_recon = O.huber_epilogue(ref_dec, x, [6], list(ref_loss), cylinder=False)
    assert abs(float(log["recon_loss_mean"]) - float(ref_recon)) <= 1e-5 * abs(float(ref_recon)) + 1e-7


def test_validation_medians_radix_select_is_exact_at_full_size():
    """vq3d_huber_elu_mask_medians on one 512x512x128 volume with the centre cylinder and 100 valid slices: bit-identical to
    torch.median of the materialised tensors (decoded >= 0, so loc == decoded and only the smooth-L1 formula is restated)."""
    import time
    from vqvae import _ops
    from vqvae.model import center_cylinder_mask
    g = torch.Generator(device=DEV).manual_seed(11)
    shape = (1, 1, 512, 512, 128)
    dec = torch.rand(*shape, generator=g, device=DEV) * 3.0
    x = torch.rand(*shape, generator=g, device=DEV) * 4.5 - 0.5
    nv = torch.tensor([100], dtype=torch.int32, device=DEV)
    keep = center_cylinder_mask(512, 512).to(DEV)
    o = _ops.default()
    got = o.huber_metrics(dec, x, nv, keep.to(torch.uint8).reshape(-1))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = o.huber_metrics(dec, x, nv, keep.to(torch.uint8).reshape(-1))
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    loc = dec.clone()
    loc[..., 100:] = 0.0
    d = (loc - x).abs()
    loss = torch.where(d < 1.0, 0.5 * d * d, d - 0.5)
    loc, loss = loc[:, :, keep], loss[:, :, keep]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ref_loc, ref_loss = loc.median(), loss.median()
    torch.cuda.synchronize()
    t_ref = time.perf_counter() - t0
    print(f"\nvalidation log of one 512x512x128 volume incl. exact medians: {t_all * 1e3:.2f} ms; the two torch.median calls alone: {t_ref * 1e3:.2f} ms")
    assert float(got["loc_median"]) == float(ref_loc) and float(got["recon_loss_median"]) == float(ref_loss)


def test_validation_metrics_fused_pass_vs_oracle():
    """VQVAE.validation_metrics (vq3d_huber_elu_mask_stats + _medians): the reference's validation log -- recon_loss / loc min, max,
    mean, median, std, nmse, psnr (model.py:143-149, metrics/evaluate.py:18-24) -- against the oracle on the oracle's own reconstruction."""
    m = _model()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = O.synthetic_volume((2, 1, 16, 16, 8), seed=5)
    ref_dec, (ref_loss, _, ref_idx) = O.vqvae_forward(sd, O.ModelConfig(**CFG), x)
    ref = O.validation_log(ref_dec, x, [8, 5], cylinder=False)
    m = m.to(DEV)
    got = m.validation_metrics((x.to(DEV), [8, 5]))
    for k, v in ref.items():
        assert abs(float(got[k]) - float(v)) <= 1e-4 * max(1.0, abs(float(v))), (k, float(got[k]), float(v))
    assert abs(float(got["loss"]) - float(ref["recon_loss_mean"] + sum(ref_loss))) <= 1e-4
