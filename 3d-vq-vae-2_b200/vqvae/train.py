"""Drop-in for the reference's vqvae/train.py (:1-67) without PyTorch-Lightning: one process per GPU (torchrun), the
model's own flags (`VQVAE.add_model_specific_args`), `--rescale-input`, `--batch-size`, `dataset_path`.

    torchrun --nproc-per-node 8 train.py synthetic:64:512x512x128 --batch-size 1 --num-embeddings 128 256 512 ...

Per step (`vqvae.parallel.training_step`): forward in training mode (EMA codebook statistics all-reduced as one flat
buffer per level), Huber + commitment loss, backward through the library's kernels, ONE flat gradient all-reduce over
NCCL, fused Adam(amsgrad).  Checkpoints are Lightning-format dicts (`state_dict`, `hyper_parameters`) so that
`VQVAE.load_from_checkpoint`, extract_embeddings.py and decode_embeddings.py -- and the reference's own scripts -- read them."""
from __future__ import annotations

import os
import time
from argparse import ArgumentParser
from pathlib import Path

import torch
import torch.distributed as dist

from vqvae.model import VQVAE
from vqvae.parallel import training_step


def parse_arguments(argv=None):
    parser = ArgumentParser()
    parser = VQVAE.add_model_specific_args(parser)
    parser.add_argument("--rescale-input", type=int, nargs="+")
    parser.add_argument("--batch-size", type=int, default=1)
    parser.add_argument("dataset_path", type=str, help="directory of .npy volumes, or synthetic:N:HxWxD")
    parser.add_argument("--hu", action="store_true")
    parser.add_argument("--max-epochs", type=int, default=int(1e5))          # train.py:40
    parser.add_argument("--max-steps", type=int, default=-1)
    parser.add_argument("--log-every-n-steps", type=int, default=50)         # train.py:35
    parser.add_argument("--default-root-dir", type=Path, default=Path("."))
    parser.add_argument("--num-workers", type=int, default=5)
    return parser.parse_args(argv)


def save_checkpoint(model: VQVAE, args, path: Path, step: int, epoch: int) -> None:
    """The subset of a Lightning checkpoint that the callers read (`extract_embeddings.py:45`, `decode_embeddings.py:23`)."""
    path.parent.mkdir(parents=True, exist_ok=True)
    torch.save({"state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
                "hyper_parameters": {"args": args}, "global_step": step, "epoch": epoch}, str(path))


def main(args):
    from utils import open_dataset
    if args.rescale_input:
        raise NotImplementedError("--rescale-input (monai resize in the reference's CTDataModule) is not part of this build: "
                                  "store the volumes at the training resolution")
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(42)                                                     # train.py:50 seed_everything(42)
    dataset = open_dataset(args.dataset_path, hu=args.hu)
    sampler = torch.utils.data.distributed.DistributedSampler(dataset, world, rank, shuffle=True, seed=42) if world > 1 else None
    loader = torch.utils.data.DataLoader(dataset, batch_size=args.batch_size, sampler=sampler, shuffle=sampler is None,
                                         num_workers=args.num_workers, pin_memory=True, drop_last=True)
    model = VQVAE(args).to(dev).train()
    if world > 1:                                                             # identical replicas (DDP's initial broadcast)
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    optimizer = model.configure_optimizers()
    ckpt_dir = Path(args.default_root_dir) / "checkpoints"
    step, t0 = 0, time.time()
    for epoch in range(args.max_epochs):
        if sampler is not None:
            sampler.set_epoch(epoch)
        for x, num_valid in loader:
            loss = training_step(model, optimizer, (x.to(dev, non_blocking=True), num_valid))
            step += 1
            if rank == 0 and step % args.log_every_n_steps == 0:
                print(f"epoch {epoch} step {step} loss {float(loss):.6f} ({(time.time() - t0) / step:.3f} s/step)", flush=True)
            if 0 < args.max_steps <= step:
                break
        if rank == 0:
            save_checkpoint(model, args, ckpt_dir / "last.ckpt", step, epoch)
        if 0 < args.max_steps <= step:
            break
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(parse_arguments())
