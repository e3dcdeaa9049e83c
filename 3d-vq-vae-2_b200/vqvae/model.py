"""Drop-in for the reference's vqvae/model.py::VQVAE (model.py:33-246) without the
PyTorch-Lightning dependency: same constructor (an argparse Namespace), attributes
(`encoder`, `decoder`, `n_bottleneck_blocks`, `num_embeddings`, `num_layers`, ...),
`forward / encode / decode`, `add_model_specific_args`, the Huber loss epilogue and a
`load_from_checkpoint` that reads Lightning-format checkpoints (`state_dict` +
`hyper_parameters`).  If pytorch_lightning is importable the class derives from
LightningModule so `pl.Trainer` accepts it; otherwise from nn.Module.
"""
from __future__ import annotations

from argparse import ArgumentParser, Namespace
from typing import Tuple

import numpy as np
import torch
from torch import nn

from . import _ops
from .layers import Decoder, Encoder2, EvonormResBlock, FixupResBlock, PreActFixupResBlock

try:  # optional
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:  # pragma: no cover - not installed in this image
    pl = None
    _Base = nn.Module


def booltype(v):
    """utils/argparse_helpers.py:1-8."""
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise ValueError("Boolean value expected.")


def center_cylinder_mask(h: int, w: int) -> torch.Tensor:
    """utils/load_nrrd_dataset.py:288-300: (h, w) bool mask, centre (h/2, w/2), radius min/2."""
    ys = np.arange(h, dtype=np.float64)[:, None]
    xs = np.arange(w, dtype=np.float64)[None, :]
    return torch.from_numpy(np.sqrt((ys - h / 2) ** 2 + (xs - w / 2) ** 2) <= min(h, w) / 2)


class VQVAE(_Base):
    supported_metrics = ("huber",)

    def __init__(self, args: Namespace):
        super().__init__()
        if pl is not None:
            self.save_hyperparameters()
        self.hparams_ns = args
        self._parse_input_args(args)
        self.encoder = Encoder2(
            in_channels=self.input_channels, base_network_channels=self.base_network_channels,
            n_enc=self.n_bottleneck_blocks, n_down_per_enc=self.n_blocks_per_bottleneck,
            n_pre_q_blocks=self.n_pre_quantization_blocks, n_post_downscale_blocks=self.n_post_downscale_blocks,
            n_post_upscale_blocks=self.n_post_upscale_blocks, num_embeddings=self.num_embeddings,
            resblock=self.resblock)
        self.decoder = Decoder(
            out_channels=self.output_channels, base_network_channels=self.base_network_channels,
            n_enc=self.n_bottleneck_blocks, n_up_per_enc=self.n_blocks_per_bottleneck,
            n_post_q_blocks=self.n_post_quantization_blocks, n_post_upscale_blocks=self.n_post_upscale_blocks,
            resblock=self.resblock)

        def init_fixup(layer):                                   # model.py:74-77
            if isinstance(layer, (FixupResBlock, PreActFixupResBlock)):
                layer.initialize_weights(num_layers=self.num_layers)
        self.apply(init_fixup)
        self._cyl_mask = None
        self._graphs = None          # shape -> captured CUDA graph (enable_cuda_graphs)

    # ---- model.py:79-89 ---------------------------------------------------------------
    def forward(self, data):
        if self._graphs is not None and not self.training and not torch.is_grad_enabled():
            return self._graphed("forward", data)
        return self._forward_eager(data)

    def _forward_eager(self, data):
        commitment_loss, quantizations, encoding_idx = zip(*self.encoder(data))
        decoded = self.decoder(quantizations)
        return decoded, (commitment_loss, quantizations, encoding_idx)

    # ---- CUDA graphs: the eval-mode forward is ~1.5k launches of small kernels; replaying a
    # captured graph removes the Python/launch overhead.  Outputs of a graphed call are static
    # buffers that the next call with the same input shape overwrites (clone to keep them).
    def enable_cuda_graphs(self, enabled: bool = True):
        self._graphs = {} if enabled else None
        return self

    def _graphed(self, what: str, data: torch.Tensor):
        key = (what, tuple(data.shape), data.device.index)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = torch.empty_like(data)
            static_in.copy_(data)
            fn = self._forward_eager if what == "forward" else (lambda d: tuple(self.encoder(d)))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # warm-up: allocator + func attributes
                fn(static_in)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = fn(static_in)
            entry = (graph, static_in, static_out)
            self._graphs[key] = entry
        graph, static_in, static_out = entry
        static_in.copy_(data, non_blocking=True)
        graph.replay()
        return static_out

    def encode(self, data):
        if self._graphs is not None and not self.training and not torch.is_grad_enabled():
            return iter(self._graphed("encode", data))
        return self.encoder(data)

    def decode(self, quantizations):
        return self.decoder(quantizations)

    # ---- CT volumes in, CT volumes out (the data module's front end + decode_embeddings' epilogue on the device) ----
    @torch.no_grad()
    def reconstruct_hu(self, hu: torch.Tensor):
        """hu: raw Hounsfield units, int16 (B, 1, H, W, D) on the device -> (reconstruction in Hounsfield units, int16, same
        shape; code indices bottom -> top).  The clip / scale / shift of utils/load_nrrd_dataset.py:73-81 and the
        ELU * 1000 - 1000 + rint of decode_embeddings.py:43-47 run on the device, so a volume crosses PCIe in 2 bytes per
        voxel each way (fp32 forward API: 4 in, 4 + indices out)."""
        o = _ops.default()
        x = o.hu_to_network(hu)
        decoded, (_, _, idx) = self(x)
        return o.elu_hu_rint(decoded, 1000.0, 1000.0, dtype=torch.int16), idx

    @torch.no_grad()
    def encode_hu(self, hu: torch.Tensor):
        """extract_embeddings.py:62-70 from raw int16 Hounsfield units: the hierarchical code indices, bottom -> top."""
        return tuple(t[2] for t in self.encode(_ops.default().hu_to_network(hu)))

    def configure_optimizers(self):
        """model.py:91-93: Adam(lr, amsgrad=True).  CUDA parameters get the library's fused step kernel."""
        if all(p.is_cuda for p in self.parameters()):
            from .optim import FusedAdamAMSGrad
            return FusedAdamAMSGrad(self.parameters(), lr=self.lr)
        return torch.optim.Adam(self.parameters(), lr=self.lr, amsgrad=True)

    # ---- loss epilogue, model.py:115-160 ----------------------------------------------
    def huber(self, batch, batch_idx=0) -> Tuple[torch.Tensor, dict]:
        """recon = mean smooth_l1(mask(ELU(decoded)), x) [over the centre cylinder if
        extract_center_cylinder]; loss = recon + sum(commitment).  One fused kernel."""
        x, num_valid = batch
        decoded, (commitment, *_) = self(x)
        nv = torch.as_tensor(num_valid, dtype=torch.int32, device=x.device).reshape(-1)
        mask = None
        if self.extract_center_cylinder:
            if self._cyl_mask is None or self._cyl_mask.numel() != x.shape[2] * x.shape[3]:
                self._cyl_mask = center_cylinder_mask(x.shape[2], x.shape[3]).to(torch.uint8).reshape(-1).to(x.device)
            mask = self._cyl_mask
        recon = _ops.default().huber_loss(decoded, x, nv, mask)
        commit = sum(commitment)
        log = {"recon_loss_mean": recon, **{f"commitment_loss_{i}": c for i, c in enumerate(commitment)}}
        return recon + commit, log

    def training_step(self, batch, batch_idx):
        return self.huber(batch, batch_idx)[0]

    def validation_step(self, batch, batch_idx):
        return self.huber(batch, batch_idx)[0]

    @torch.no_grad()
    def validation_metrics(self, batch) -> dict:
        """The reference's validation log (model.py:143-160) without materialising the masked reconstruction or the unreduced
        loss: recon_loss / loc min, max, mean, median, std (utils/logging_helpers.py:4-15), nmse, psnr (data_range 4), the
        commitment losses and the total loss -- one fused pass for the sums, a four-pass radix select for the exact medians."""
        x, num_valid = batch
        decoded, (commitment, *_) = self(x)
        nv = torch.as_tensor(num_valid, dtype=torch.int32, device=x.device).reshape(-1)
        mask = None
        if self.extract_center_cylinder:
            if self._cyl_mask is None or self._cyl_mask.numel() != x.shape[2] * x.shape[3]:
                self._cyl_mask = center_cylinder_mask(x.shape[2], x.shape[3]).to(torch.uint8).reshape(-1).to(x.device)
            mask = self._cyl_mask
        log = _ops.default().huber_metrics(decoded, x, nv, mask)
        log.update({f"commitment_loss_{i}": c for i, c in enumerate(commitment)})
        log["loss"] = log["recon_loss_mean"] + sum(commitment)
        return log

    # ---- model.py:165-210 -------------------------------------------------------------
    def _parse_input_args(self, args: Namespace):
        assert args.metric in self.supported_metrics
        self.metric = args.metric
        self.lr = args.base_lr
        self.input_channels = args.input_channels
        self.output_channels = args.input_channels
        self.base_network_channels = args.base_network_channels
        self.n_bottleneck_blocks = args.n_bottleneck_blocks
        self.n_blocks_per_bottleneck = args.n_downscales_per_bottleneck
        self.n_pre_quantization_blocks = args.n_pre_quantization_blocks
        self.n_post_quantization_blocks = args.n_post_quantization_blocks
        self.n_post_upscale_blocks = args.n_post_upscale_blocks
        self.n_post_downscale_blocks = args.n_post_downscale_blocks
        ne = list(args.num_embeddings) if isinstance(args.num_embeddings, (list, tuple)) else [args.num_embeddings]
        assert len(ne) in (1, args.n_bottleneck_blocks)
        self.num_embeddings = ne * args.n_bottleneck_blocks if len(ne) == 1 else ne
        self.resblock = {"regular": FixupResBlock, "pre-activation": PreActFixupResBlock,
                         "evonorm": EvonormResBlock}[args.block_type]
        n_down = args.n_bottleneck_blocks * args.n_downscales_per_bottleneck
        self.num_layers = (2 + 2 * n_down + args.n_pre_quantization_blocks + args.n_post_quantization_blocks
                           + args.n_post_downscale_blocks * n_down + args.n_post_upscale_blocks * n_down + 1)
        self.extract_center_cylinder = bool(getattr(args, "extract_center_cylinder", True))

    @classmethod
    def add_model_specific_args(cls, parent_parser):
        """Same flags (including the mixed spellings) as model.py:213-246."""
        p = ArgumentParser(parents=[parent_parser], add_help=False)
        p.add_argument("--input-channels", type=int, default=1)
        p.add_argument("--base-network_channels", type=int, default=4)
        p.add_argument("--n-bottleneck-blocks", type=int, default=3)
        p.add_argument("--n-downscales-per-bottleneck", type=int, default=2)
        p.add_argument("--n-pre-quantization-blocks", type=int, default=0)
        p.add_argument("--n-post-quantization-blocks", type=int, default=0)
        p.add_argument("--n-post-upscale-blocks", type=int, default=0)
        p.add_argument("--n-post-downscale-blocks", type=int, default=0)
        p.add_argument("--num-embeddings", type=int, default=256, nargs="+")
        p.add_argument("--block-type", type=str, default="pre-activation", choices=["regular", "pre-activation", "evonorm"])
        p.add_argument("--extract-center-cylinder", type=booltype, default=True)
        p.add_argument("--metric", choices=cls.supported_metrics, default=cls.supported_metrics[0])
        p.add_argument("--base_lr", default=1e-5, type=float)
        p.add_argument("--n-mix", default=2)
        return p

    @classmethod
    def default_args(cls, **overrides) -> Namespace:
        ns = cls.add_model_specific_args(ArgumentParser(add_help=False)).parse_args([])
        for k, v in overrides.items():
            if not hasattr(ns, k):
                raise AttributeError(k)
            setattr(ns, k, v)
        return ns

    @classmethod
    def load_from_checkpoint(cls, path, map_location="cpu", **kw):
        """Reads a Lightning checkpoint written by the reference's train.py (`state_dict`,
        `hyper_parameters['args']`), extract_embeddings.py:45 / decode_embeddings.py:23."""
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
        hp = ckpt.get("hyper_parameters", {})
        args = hp.get("args", hp)
        if isinstance(args, dict):
            args = Namespace(**args)
        model = cls(args)
        sd = {k: v for k, v in ckpt["state_dict"].items() if k.startswith(("encoder.", "decoder."))}
        model.load_state_dict(sd, strict=True)
        return model


# the two published configurations (slurm-jobs/train_vqvae_3d.job:77-86, train_vqvae_3d_downscaled.job:76-88)
def full_config_args() -> Namespace:
    return VQVAE.default_args(n_bottleneck_blocks=3, num_embeddings=[128, 256, 512], n_pre_quantization_blocks=50,
                              n_post_quantization_blocks=50, n_post_upscale_blocks=3, n_post_downscale_blocks=2)


def downscaled_config_args() -> Namespace:
    return VQVAE.default_args(n_bottleneck_blocks=2, num_embeddings=[128, 256], n_pre_quantization_blocks=150,
                              n_post_quantization_blocks=150, n_post_upscale_blocks=5, n_post_downscale_blocks=5)
