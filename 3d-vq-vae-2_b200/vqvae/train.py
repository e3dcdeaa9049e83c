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
    # pl.Trainer flag the reference's job scripts use (train_vqvae_3d.job:87): continue from a Lightning-format checkpoint
    parser.add_argument("--resume-from-checkpoint", "--resume_from_checkpoint", dest="resume_from_checkpoint", type=Path, default=None)
    parser.add_argument("--val-dataset-path", type=str, default=None,
                        help="validation volumes (same syntax as dataset_path); monitored as val_recon_loss_mean like train.py:56")
    parser.add_argument("--val-check-interval", type=float, default=0.5)      # train.py:36 (fraction of an epoch)
    return parser.parse_args(argv)


def save_checkpoint(model: VQVAE, args, path: Path, step: int, epoch: int, optimizer=None, monitor=None) -> None:
    """The subset of a Lightning checkpoint that the callers read (`extract_embeddings.py:45`, `decode_embeddings.py:23`)
    plus what a resume needs: `optimizer_states` (Lightning's key; torch.optim.Adam's state layout), step and epoch."""
    path.parent.mkdir(parents=True, exist_ok=True)
    ckpt = {"state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
            "hyper_parameters": {"args": args}, "global_step": step, "epoch": epoch}
    if optimizer is not None:
        osd = optimizer.state_dict()
        osd["state"] = {k: {n: (t.detach().cpu().clone() if torch.is_tensor(t) else t) for n, t in st.items()} for k, st in osd["state"].items()}
        ckpt["optimizer_states"] = [osd]
    if monitor is not None:
        ckpt["val_recon_loss_mean"] = float(monitor)
    torch.save(ckpt, str(path))


def load_checkpoint(model: VQVAE, optimizer, path) -> dict:
    """Resume (`--resume-from-checkpoint`, train_vqvae_3d.job:87): weights + EMA codebook buffers in place (the flat
    parameter buffer of the optimizer stays the storage), then the optimizer's moments and step counter."""
    ckpt = torch.load(str(path), map_location="cpu", weights_only=False)
    model.load_state_dict({k: v for k, v in ckpt["state_dict"].items() if k.startswith(("encoder.", "decoder."))}, strict=True)
    if optimizer is not None and ckpt.get("optimizer_states"):
        optimizer.load_state_dict(ckpt["optimizer_states"][0])
    return ckpt


@torch.no_grad()
def validate(model: VQVAE, loader, dev, world: int) -> dict:
    """The validation log of the reference (model.py:143-160), averaged over the validation volumes of all ranks:
    val_recon_loss_mean (the checkpoint monitor, train.py:56), nmse, psnr -- one fused pass per volume
    (`VQVAE.validation_metrics`)."""
    was_training = model.training
    model.eval()
    keys = ("recon_loss_mean", "nmse", "psnr")
    acc = torch.zeros(len(keys) + 1, dtype=torch.float64, device=dev)
    for x, num_valid in loader:
        log = model.validation_metrics((x.to(dev, non_blocking=True), num_valid))
        for i, k in enumerate(keys):
            acc[i] += log[k].double()
        acc[-1] += 1
    if world > 1:
        dist.all_reduce(acc)
    model.train(was_training)
    n = float(acc[-1].clamp(min=1))
    return {f"val_{k}": float(acc[i]) / n for i, k in enumerate(keys)}


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
    step, first_epoch, best = 0, 0, float("inf")
    if args.resume_from_checkpoint is not None:
        ckpt = load_checkpoint(model, optimizer, args.resume_from_checkpoint)
        step, first_epoch = int(ckpt.get("global_step", 0)), int(ckpt.get("epoch", -1)) + 1
        best = float(ckpt.get("val_recon_loss_mean", best))
        if rank == 0:
            print(f"resumed from {args.resume_from_checkpoint}: step {step}, epoch {first_epoch}", flush=True)
    val_loader = None
    if args.val_dataset_path:
        val_set = open_dataset(args.val_dataset_path, hu=args.hu)
        val_sampler = torch.utils.data.distributed.DistributedSampler(val_set, world, rank, shuffle=False) if world > 1 else None
        val_loader = torch.utils.data.DataLoader(val_set, batch_size=args.batch_size, sampler=val_sampler, shuffle=False,
                                                 num_workers=args.num_workers, pin_memory=True)
    val_every = max(1, int(len(loader) * args.val_check_interval)) if val_loader is not None else 0

    def run_validation(epoch):
        """ModelCheckpoint(save_top_k=1, save_last=True, monitor='val_recon_loss_mean'), train.py:56."""
        nonlocal best
        log = validate(model, val_loader, dev, world)
        val = log["val_recon_loss_mean"]
        if rank == 0:
            print(f"epoch {epoch} step {step} " + " ".join(f"{k} {v:.6f}" for k, v in log.items()), flush=True)
            if val < best:
                best = val
                save_checkpoint(model, args, ckpt_dir / "best.ckpt", step, epoch, optimizer, monitor=val)

    start_step, t0 = step, time.time()
    for epoch in range(first_epoch, args.max_epochs):
        if sampler is not None:
            sampler.set_epoch(epoch)
        for i, (x, num_valid) in enumerate(loader):
            loss = training_step(model, optimizer, (x.to(dev, non_blocking=True), num_valid))
            step += 1
            if rank == 0 and step % args.log_every_n_steps == 0:
                print(f"epoch {epoch} step {step} loss {float(loss):.6f} ({(time.time() - t0) / (step - start_step):.3f} s/step)", flush=True)
            if val_every and (i + 1) % val_every == 0:
                run_validation(epoch)
            if 0 < args.max_steps <= step:
                break
        if rank == 0:
            save_checkpoint(model, args, ckpt_dir / "last.ckpt", step, epoch, optimizer, monitor=best if best < float("inf") else None)
        if 0 < args.max_steps <= step:
            break
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(parse_arguments())
