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
def decode_codes(model: VQVAE, codes) -> torch.Tensor:
    """codes: per level (bottom -> top) a LongTensor (h, w, d) or (B, h, w, d) -> Hounsfield units, int64 (B, 1, H, W, D)."""
    embeddings = []
    for idx, quantizer in zip(codes, model.encoder.quantize):
        idx = idx.cuda()
        if idx.dim() == 3:
            idx = idx.unsqueeze(dim=0)
        embeddings.append(quantizer.embed_code(idx).permute(0, 4, 1, 2, 3).contiguous())       # decode_embeddings.py:36-40
    res = model.decode(embeddings)
    return _ops.default().elu_hu_rint(res, SCALE_VAL, SCALE_VAL)


@torch.no_grad()
def main(args: Namespace):
    from utils import write_nrrd
    print("- Loading model weights")
    model = VQVAE.load_from_checkpoint(str(args.ckpt_path)).cuda().eval()
    db = torch.load(args.db_path, weights_only=False)
    for embedding_0_key, embedding_0 in db[0].items():
        embedding_1_key = embedding_0["condition"]
        embedding_1 = db[1][embedding_1_key]
        # issue where the pixelcnn samples 0's (decode_embeddings.py:32-33)
        success = "failure" if torch.all(embedding_0["data"][-1] == 0) else "success"
        print("- Performing forward pass")
        res = decode_codes(model, (embedding_0["data"], embedding_1["data"]))
        res = res.squeeze().cpu().numpy()
        print("- Writing to nrrd")
        write_nrrd(str(args.out_path) + f"_{success}_{str(embedding_1_key)}_{str(embedding_0_key)}.nrrd", res,
                   header={"spacings": (0.976, 0.976, 3)})
        print("- Done")


if __name__ == "__main__":
    parser = ArgumentParser()
    parser.add_argument("db_path", type=Path)
    parser.add_argument("ckpt_path", type=Path)
    parser.add_argument("out_path", type=Path, help="outpath without extension")
    main(parser.parse_args())
