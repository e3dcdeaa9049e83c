"""Drop-in for the reference's vqvae/decode_embeddings.py (:1-60): turn sampled code indices back into CT volumes.

    python decode_embeddings.py db_path ckpt_path out_path

`db_path` is the torch-saved dict the PixelCNN sampler writes: db[level][key] = {'data': LongTensor (h, w, d),
'condition': key of the level above} (decode_embeddings.py:27-33).  Per sample: `Quantizer.embed_code` (gather kernel) for
both levels, `VQVAE.decode` (the decoder's CUDA kernels), then ELU, `* 1000 - 1000`, `np.rint`, `astype(int)`
(:43-47) as ONE kernel (`vq3d_elu_hu_rint`), and an NRRD file with the reference's name pattern and spacings (:50)."""
from __future__ import annotations

from argparse import ArgumentParser, Namespace
from pathlib import Path

import torch

from vqvae import _ops
from vqvae.model import VQVAE

MIN_VAL, MAX_VAL, SCALE_VAL = -1500, 3000, 1000          # decode_embeddings.py:19


@torch.no_grad()
def decode_codes(model: VQVAE, codes, dtype=torch.int64) -> torch.Tensor:
    """codes: per level (bottom -> top) a LongTensor (h, w, d) or (B, h, w, d) -> Hounsfield units (B, 1, H, W, D), int64 like
    the reference's `astype(int)` or int16 (what the NRRD writer needs for CT: a quarter of the device -> host bytes)."""
    embeddings = []
    for idx, quantizer in zip(codes, model.encoder.quantize):
        idx = idx.cuda()
        if idx.dim() == 3:
            idx = idx.unsqueeze(dim=0)
        embeddings.append(quantizer.embed_code(idx).permute(0, 4, 1, 2, 3).contiguous())       # decode_embeddings.py:36-40
    res = model.decode(embeddings)
    return _ops.default().elu_hu_rint(res, SCALE_VAL, SCALE_VAL, dtype=dtype)


def iter_samples(db):
    """The sampler's database (decode_embeddings.py:27-33): every bottom-level entry names the top-level entry it was
    conditioned on.  Yields (bottom key, top key, (bottom codes, top codes), sampled_ok); a bottom map whose last slice is
    all zeros is the PixelCNN failure mode the reference flags in the file name."""
    bottom, top = db[0], db[1]
    for key0, entry0 in bottom.items():
        key1 = entry0["condition"]
        ok = not bool(torch.all(entry0["data"][-1] == 0))
        yield key0, key1, (entry0["data"], top[key1]["data"]), ok


def output_name(out_path, ok: bool, key1, key0) -> str:
    """decode_embeddings.py:50: <out>_<success|failure>_<top key>_<bottom key>.nrrd"""
    return f"{out_path}_{'success' if ok else 'failure'}_{key1}_{key0}.nrrd"


@torch.no_grad()
def main(args: Namespace):
    from utils import write_nrrd
    model = VQVAE.load_from_checkpoint(str(args.ckpt_path)).cuda().eval()
    db = torch.load(args.db_path, weights_only=False)
    for n, (key0, key1, codes, ok) in enumerate(iter_samples(db)):
        hu = decode_codes(model, codes).squeeze().cpu().numpy()
        path = output_name(args.out_path, ok, key1, key0)
        write_nrrd(path, hu, header={"spacings": (0.976, 0.976, 3)})
        print(f"[{n}] {path}: {hu.shape}, HU range [{int(hu.min())}, {int(hu.max())}]")


if __name__ == "__main__":
    cli = ArgumentParser(description=__doc__.splitlines()[0])
    cli.add_argument("db_path", type=Path, help="torch-saved sampler output")
    cli.add_argument("ckpt_path", type=Path, help="Lightning-format VQ-VAE checkpoint")
    cli.add_argument("out_path", type=Path, help="output path without extension")
    main(cli.parse_args())
