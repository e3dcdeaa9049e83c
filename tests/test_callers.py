"""Host logic of the three caller scripts: the LMDB layout of extract_embeddings.py read back the way
the reference's utils/load_lmdb_dataset.py:62-109 does, the NRRD file of decode_embeddings.py, the Lightning-format
checkpoint of train.py, and the fused HU epilogue kernel on the host emulator."""
import os
import pickle
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-vq-vae-2_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


class FakeTxn:
    def __init__(self, env, write):
        self.env, self.write = env, write

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def put(self, key, value, db=None):
        assert self.write and isinstance(key, bytes) and isinstance(value, bytes)
        self.env.store.setdefault(db, {})[key] = value

    def get(self, key, db=None):
        return self.env.store.get(db, {}).get(key)

    def cursor(self, db=None):
        return iter(sorted(self.env.store.get(db, {}).items()))


class FakeEnv:
    """The part of lmdb.Environment that extract_embeddings.py and the reference's reader use."""

    def __init__(self, max_dbs):
        self.max_dbs, self.store, self.names = max_dbs, {}, []

    def open_db(self, name):
        assert isinstance(name, bytes)
        if name not in self.names:
            assert len(self.names) < self.max_dbs
            self.names.append(name)
        return name

    def begin(self, write=False, db=None):
        return FakeTxn(self, write)


def test_lmdb_layout_round_trip_like_the_reference_reader():
    from vqvae.extract_embeddings import write_codes
    rs = np.random.RandomState(0)
    shapes = [(1, 8, 8, 4), (1, 2, 2, 1)]
    samples = [[rs.randint(0, k, size=s).astype(np.int64) for s, k in zip(shapes, (16, 32))] for _ in range(3)]
    env = FakeEnv(max_dbs=2)
    assert write_codes(env, 2, [16, 32], 3, iter(samples)) == 3
    # utils/load_lmdb_dataset.py:70-86: root metadata, then one sub-db per level keyed by str(index)
    with env.begin() as txn:
        num_dbs = int(txn.get(b"num_dbs"))
        length = int(txn.get(b"length"))
        num_embeddings = pickle.loads(txn.get(b"num_embeddings"))
    assert (num_dbs, length) == (2, 3) and np.array_equal(num_embeddings, np.asarray([16, 32]))
    for level in range(num_dbs):
        sub = env.open_db(str(level).encode())
        with env.begin(db=sub) as txn:
            for i in range(length):
                arr = pickle.loads(txn.get(str(i).encode(), db=sub))
                assert arr.dtype == np.int64 and arr.shape == shapes[level]
                assert np.array_equal(arr, samples[i][level])


def test_output_path_rules(tmp_path):
    from vqvae.extract_embeddings import get_output_abspath
    ck = tmp_path / "lightning_logs" / "version_12" / "checkpoints" / "last.ckpt"
    ck.parent.mkdir(parents=True)
    ck.write_bytes(b"x")
    assert get_output_abspath(ck, tmp_path).endswith("version_12_last.lmdb")          # extract_embeddings.py:33-36
    plain = tmp_path / "model.ckpt"
    plain.write_bytes(b"x")
    assert get_output_abspath(plain, tmp_path).endswith("model.lmdb")
    assert get_output_abspath(plain, tmp_path, "codes.lmdb").endswith("codes.lmdb")


def test_nrrd_round_trip(tmp_path):
    from utils import read_nrrd, write_nrrd
    vol = np.arange(4 * 3 * 2, dtype=np.int64).reshape(4, 3, 2) - 7
    path = str(tmp_path / "v.nrrd")
    write_nrrd(path, vol, header={"spacings": (0.976, 0.976, 3)})
    raw = open(path, "rb").read()
    head = raw.split(b"\n\n", 1)[0].decode()
    assert head.splitlines()[0] == "NRRD0004" and "type: int64" in head and "sizes: 4 3 2" in head and "spacings: 0.976 0.976 3" in head
    # pynrrd's default index order: the FIRST axis varies fastest in the payload
    payload = np.frombuffer(raw.split(b"\n\n", 1)[1], dtype="<i8")
    assert payload[1] == vol[1, 0, 0] and payload[4] == vol[0, 1, 0]
    back, fields = read_nrrd(path)
    assert np.array_equal(back, vol) and fields["encoding"] == "raw"


def test_volume_sources():
    from utils import open_dataset
    from utils.volumes import preprocess_hu
    ds = open_dataset("synthetic:3:8x6x4")
    x, nv = ds[1]
    assert len(ds) == 3 and x.shape == (1, 8, 6, 4) and nv == 4 and -0.5 <= float(x.min()) and float(x.max()) <= 4.0
    assert torch.equal(ds[1][0], x)                                                    # seeded
    hu = np.array([-2000.0, -1000.0, 0.0, 3500.0])
    assert np.allclose(preprocess_hu(hu), [-0.5, 0.0, 1.0, 4.0])                       # load_nrrd_dataset.py:73-81


def test_train_arguments_and_checkpoint_round_trip(tmp_path):
    from vqvae import train
    from vqvae.model import VQVAE
    args = train.parse_arguments(["synthetic:2:8x8x8", "--batch-size", "1", "--n-bottleneck-blocks", "2", "--num-embeddings", "8", "12",
                                  "--n-downscales-per-bottleneck", "1", "--n-pre-quantization-blocks", "1", "--base_lr", "1e-4",
                                  "--block-type", "pre-activation", "--extract-center-cylinder", "False"])
    assert args.num_embeddings == [8, 12] and args.base_lr == 1e-4 and args.extract_center_cylinder is False
    torch.manual_seed(0)
    m = VQVAE(args)
    train.save_checkpoint(m, args, tmp_path / "checkpoints" / "last.ckpt", step=3, epoch=0)
    m2 = VQVAE.load_from_checkpoint(str(tmp_path / "checkpoints" / "last.ckpt"))
    assert m2.num_embeddings == [8, 12]
    for (k, a), (_, b) in zip(sorted(m.state_dict().items()), sorted(m2.state_dict().items())):
        assert torch.equal(a, b), k


def test_elu_hu_rint_kernel_on_the_emulator():
    """rint(ELU(x) * 1000 - 1000) as int64, decode_embeddings.py:43-47 (np.rint: half to even)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu.emu_ops import use_emulator
    from vqvae import _ops
    x = torch.tensor([-3.0, -0.5, 0.0, 0.0005, 0.0015, 0.0025, 1.0, 2.4996, 3.99951], dtype=torch.float32).reshape(1, 1, 3, 3, 1)
    ref = np.rint(torch.nn.functional.elu(x).numpy() * 1000 - 1000).astype(np.int64)
    with use_emulator():
        got = _ops.default().elu_hu_rint(x)
    assert got.dtype == torch.int64 and np.array_equal(got.numpy(), ref)


def test_int16_hounsfield_kernels_on_the_emulator():
    """vq3d_hu_to_network (clip / * 0.001f / + 1: utils/load_nrrd_dataset.py:73-81) against the numpy front end, and
    vq3d_elu_hu_rint_i16 against the int64 epilogue incl. saturation and the n % 4 tail."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu.emu_ops import use_emulator
    from utils.volumes import preprocess_hu
    from vqvae import _ops
    rs = np.random.RandomState(0)
    hu = rs.randint(-4000, 6000, size=(1, 1, 5, 7, 9)).astype(np.int16)          # 315 elements: exercises the scalar tail
    hu.reshape(-1)[:4] = [-32768, 32767, -1500, 3000]
    with use_emulator():
        got = _ops.default().hu_to_network(torch.from_numpy(hu))
        assert got.dtype == torch.float32 and np.array_equal(got.numpy(), preprocess_hu(hu))
        assert abs(float(got.min()) + 0.5) < 1e-6 and abs(float(got.max()) - 4.0) < 1e-6      # the reference's value range
        x = torch.from_numpy(rs.standard_normal((1, 1, 5, 7, 9)).astype(np.float32) * 2)
        x.reshape(-1)[:3] = torch.tensor([40.0, -50.0, 0.0005])                    # 39 000 HU saturates to 32 767
        a = _ops.default().elu_hu_rint(x)
        b = _ops.default().elu_hu_rint(x, dtype=torch.int16)
    assert b.dtype == torch.int16 and np.array_equal(b.numpy(), np.clip(a.numpy(), -32768, 32767).astype(np.int16))
    assert int(b.reshape(-1)[0]) == 32767 and int(b.reshape(-1)[1]) == -2000


@pytest.mark.parametrize("cyl", [False, True])
def test_validation_metrics_kernel_on_the_emulator(cyl):
    """vq3d_huber_elu_mask_stats: recon_loss / loc min, max, mean, std + nmse + psnr (model.py:143-149, metrics/evaluate.py:18-24)
    from one pass, against the oracle's restatement."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu.emu_ops import use_emulator
    from oracle import vqvae_oracle as O
    from vqvae import _ops
    from vqvae.model import center_cylinder_mask
    torch.manual_seed(3)
    dec = torch.randn(2, 1, 10, 12, 9) * 1.5
    x = torch.rand(2, 1, 10, 12, 9) * 4.5 - 0.5
    nv = [9, 6]
    ref = O.validation_log(dec, x, nv, cylinder=cyl)
    mask = center_cylinder_mask(10, 12).to(torch.uint8).reshape(-1) if cyl else None
    with use_emulator():
        got = _ops.default().huber_metrics(dec, x, torch.tensor(nv, dtype=torch.int32), mask)
    assert set(got) == set(ref)
    for k in ref:
        assert abs(float(got[k]) - float(ref[k])) <= 2e-5 * max(1.0, abs(float(ref[k]))), (k, float(got[k]), float(ref[k]))


@pytest.mark.parametrize("shape,nv,cyl", [((1, 1, 6, 5, 7), [7], False),      # 210 voxels: even count -> the LOWER middle element
                                          ((1, 1, 5, 5, 7), [4], False),      # 175 voxels, 75 of them masked to exactly 0
                                          ((2, 1, 8, 8, 5), [5, 2], True)])   # centre cylinder: a subset of the voxels
def test_validation_medians_are_the_tensor_elements_torch_median_returns(shape, nv, cyl):
    """vq3d_huber_elu_mask_medians (utils/logging_helpers.py:13): a radix select, so the result must be bit-identical to
    torch.median of the materialised tensors.  decoded >= 0 keeps ELU exact (loc == decoded), so the only arithmetic left is
    the smooth-L1 formula, restated here with the kernel's operations."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu.emu_ops import use_emulator
    from vqvae import _ops
    from vqvae.model import center_cylinder_mask
    g = torch.Generator().manual_seed(sum(shape))
    dec = torch.rand(*shape, generator=g) * 3.0
    dec.reshape(-1)[::7] = 1.25                                               # ties around the middle
    x = torch.rand(*shape, generator=g) * 4.5 - 0.5
    loc = dec.clone()
    for b, n in enumerate(nv):
        loc[b, ..., n:] = 0.0
    d = (loc - x).abs()
    loss = torch.where(d < 1.0, 0.5 * d * d, d - 0.5)
    mask = None
    if cyl:
        keep = center_cylinder_mask(shape[2], shape[3])
        mask = keep.to(torch.uint8).reshape(-1)
        loc, loss = loc[:, :, keep], loss[:, :, keep]
    with use_emulator():
        got = _ops.default().huber_metrics(dec, x, torch.tensor(nv, dtype=torch.int32), mask)
    assert float(got["loc_median"]) == float(loc.median()), (float(got["loc_median"]), float(loc.median()))
    assert float(got["recon_loss_median"]) == float(loss.median()), (float(got["recon_loss_median"]), float(loss.median()))


def test_validation_medians_propagate_nan_like_torch_median():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu.emu_ops import use_emulator
    from vqvae import _ops
    dec = torch.rand(1, 1, 4, 4, 4)
    dec[0, 0, 1, 2, 3] = float("nan")
    x = torch.rand(1, 1, 4, 4, 4)
    with use_emulator():
        got = _ops.default().huber_metrics(dec, x, torch.tensor([4], dtype=torch.int32), None)
    assert torch.isnan(got["loc_median"]) and torch.isnan(got["recon_loss_median"])
    assert torch.isnan(torch.nn.functional.elu(dec).median())


def test_decode_database_iteration_and_names():
    from vqvae.decode_embeddings import iter_samples, output_name
    db = {0: {"a": {"data": torch.ones(2, 2, 2, dtype=torch.long), "condition": "t1"},
              "b": {"data": torch.zeros(2, 2, 2, dtype=torch.long), "condition": "t1"}},
          1: {"t1": {"data": torch.full((1, 1, 1), 3, dtype=torch.long)}}}
    got = list(iter_samples(db))
    assert [(k0, k1, ok) for k0, k1, _, ok in got] == [("a", "t1", True), ("b", "t1", False)]     # all-zero last slice = sampler failure
    assert torch.equal(got[0][2][1], db[1]["t1"]["data"])
    assert output_name("out/vol", True, "t1", "a") == "out/vol_success_t1_a.nrrd"                # decode_embeddings.py:50
    assert output_name("out/vol", False, "t1", "b") == "out/vol_failure_t1_b.nrrd"
