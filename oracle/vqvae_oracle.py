"""
oracle/vqvae_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's encode -> quantize -> decode path
(sara-nl/3D-VQ-VAE-2: vqvae/layers.py, vqvae/evonorm.py, vqvae/model.py) written as
pure functions over a flat ``state_dict`` (the reference's own key names) and a
``ModelConfig``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product
package (``3d-vq-vae-2_b200/``) never does and fails loudly without its CUDA library.

Where the arithmetic lives.  The reference contains no arithmetic of its own on this
path: every number is produced by ATen CPU kernels (conv3d, upsample_trilinear3d,
circular pad, elu, cdist, argmin, embedding ...; torch pinned 1.10.1 in
environment.yml:120, 2.11.0 in this image).  The float convolution part of this oracle
therefore calls the same ``torch.nn.functional`` ops in the same order as the reference
modules, which makes it bit-identical to the imported reference on the same host.
The integer-valued part (nearest-code search, tie-breaking, EMA counts) is restated
explicitly in C (oracle/vq_oracle.c) and numpy below, with the summation order of
ATen's cdist kernel, so that it is deterministic on any host.

Parity pinning: every function here is checked against outputs of the imported
reference modules committed under tests/golden/ (generator: tests/golden/make_golden.py);
see tests/test_oracle_golden.py.  The reference ships no golden vectors of its own
(SURVEY.md section 4), the only reference test on this path (evonorm.py:79-98) is
restated in tests/test_oracle_golden.py::test_silu_velocity_reference_selftest.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------
# configuration (restates the flag handling of vqvae/model.py:165-210)
# --------------------------------------------------------------------------------------
@dataclass
class ModelConfig:
    input_channels: int = 1
    base_network_channels: int = 4
    n_bottleneck_blocks: int = 3           # n_enc
    n_downscales_per_bottleneck: int = 2   # n_down_per_enc / n_up_per_enc
    n_pre_quantization_blocks: int = 0
    n_post_quantization_blocks: int = 0
    n_post_upscale_blocks: int = 0
    n_post_downscale_blocks: int = 0
    num_embeddings: Sequence[int] = (256,)
    block_type: str = "pre-activation"     # 'regular' | 'pre-activation' | 'evonorm'

    def embeddings(self) -> List[int]:
        ne = list(self.num_embeddings)
        if len(ne) == 1:
            ne = ne * self.n_bottleneck_blocks      # model.py:184-188
        assert len(ne) == self.n_bottleneck_blocks
        return ne

    def num_layers(self) -> int:
        """Longest path through the model, model.py:194-203."""
        n_down = self.n_bottleneck_blocks * self.n_downscales_per_bottleneck
        return (2 + 2 * n_down + self.n_pre_quantization_blocks + self.n_post_quantization_blocks
                + self.n_post_downscale_blocks * n_down + self.n_post_upscale_blocks * n_down + 1)


# the two published configurations (slurm-jobs/train_vqvae_3d.job:77-86 and
# slurm-jobs/train_vqvae_3d_downscaled.job:76-88)
FULL = ModelConfig(n_bottleneck_blocks=3, num_embeddings=(128, 256, 512),
                   n_pre_quantization_blocks=50, n_post_quantization_blocks=50,
                   n_post_upscale_blocks=3, n_post_downscale_blocks=2)
DOWNSCALED = ModelConfig(n_bottleneck_blocks=2, num_embeddings=(128, 256),
                         n_pre_quantization_blocks=150, n_post_quantization_blocks=150,
                         n_post_upscale_blocks=5, n_post_downscale_blocks=5)


# --------------------------------------------------------------------------------------
# C part of the oracle (nearest code search in the reference's exact fp32 order)
# --------------------------------------------------------------------------------------
_LIB: Optional[ctypes.CDLL] = None


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/vq_oracle.c with oracle/Makefile; returns the .so path."""
    so = os.path.join(_HERE, "_build", "libvq_oracle.so")
    src = os.path.join(_HERE, "vq_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return so


def _lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        lib.vq_oracle_assign.restype = ctypes.c_int
        lib.vq_oracle_assign.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        lib.vq_oracle_cdist.restype = None
        lib.vq_oracle_cdist.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        lib.vq_oracle_stats.restype = None
        lib.vq_oracle_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def vq_assign_c(flat: np.ndarray, embed: np.ndarray, threads: int = 0) -> Tuple[np.ndarray, int]:
    """idx[i] = first k minimising sqrt(sum_d (x_id - e_kd)^2)  (layers.py:700-702)."""
    flat = np.ascontiguousarray(flat, dtype=np.float32)
    embed = np.ascontiguousarray(embed, dtype=np.float32)
    n, d = flat.shape
    k = embed.shape[0]
    assert embed.shape[1] == d
    idx = np.empty(n, dtype=np.int64)
    used = _lib().vq_oracle_assign(flat.ctypes.data, embed.ctypes.data, n, d, k, idx.ctypes.data, threads)
    return idx, used


def cdist_c(flat: np.ndarray, embed: np.ndarray) -> np.ndarray:
    flat = np.ascontiguousarray(flat, dtype=np.float32)
    embed = np.ascontiguousarray(embed, dtype=np.float32)
    out = np.empty((flat.shape[0], embed.shape[0]), dtype=np.float32)
    _lib().vq_oracle_cdist(flat.ctypes.data, embed.ctypes.data, flat.shape[0], flat.shape[1],
                           embed.shape[0], out.ctypes.data)
    return out


def vq_stats_c(flat: np.ndarray, idx: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    flat = np.ascontiguousarray(flat, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    n = np.empty(k, dtype=np.float64)
    dw = np.empty((k, flat.shape[1]), dtype=np.float64)
    _lib().vq_oracle_stats(flat.ctypes.data, idx.ctypes.data, flat.shape[0], flat.shape[1], k,
                           n.ctypes.data, dw.ctypes.data)
    return n, dw


def cdist_numpy(flat: np.ndarray, embed: np.ndarray) -> np.ndarray:
    """numpy statement of the same order (slow; used to cross-check the C file)."""
    flat = flat.astype(np.float32)
    embed = embed.astype(np.float32)
    n, d = flat.shape
    nv = (d // 4) * 4
    agg = np.zeros((n, embed.shape[0]), np.float32)
    for j in range(d):
        diff = (flat[:, None, j] - embed[None, :, j]).astype(np.float32)
        if j < nv:
            agg = (agg + (diff * diff).astype(np.float32)).astype(np.float32)
        else:  # fused multiply-add: the product is exact in float64, one rounding to fp32
            agg = (diff.astype(np.float64) * diff.astype(np.float64) + agg.astype(np.float64)).astype(np.float32)
    return np.sqrt(agg)


# --------------------------------------------------------------------------------------
# Quantizer (layers.py:602-728)
# --------------------------------------------------------------------------------------
def quantizer_forward(sd: Dict[str, Tensor], prefix: str, x: Tensor, training: bool,
                      commitment_cost: float = 0.1, decay: float = 0.99, laplace_alpha: float = 1e-5,
                      world_size: int = 1, use_torch_cdist: bool = False
                      ) -> Tuple[Tensor, Tensor, Tensor]:
    """Returns (loss, quantized, idx); mutates the four buffers in ``sd`` in place when
    ``training`` (layers.py:685-728).  ``world_size`` only scales the first-pass
    cluster_size as layers.py:676 does (single-process oracle: the all-reduce of
    identical per-rank tensors is a multiplication)."""
    embed, embed_avg = sd[prefix + "embed"], sd[prefix + "embed_avg"]
    cluster_size, first_pass = sd[prefix + "cluster_size"], sd[prefix + "first_pass"]
    k, d = embed.shape
    x = x.float()
    chan_last = x.permute(0, 2, 3, 4, 1)                 # layers.py:690
    shape = chan_last.shape
    flat = chan_last.reshape(-1, d)                       # layers.py:693

    if training and int(first_pass) != 0:                 # _init_ema, layers.py:665-683
        mean = flat.mean(dim=0)
        std = flat.std(dim=0)                             # unbiased
        embed.mul_(std).add_(mean)
        embed_avg.copy_(embed)
        cluster_size.add_(flat.shape[0] * world_size / k)
        first_pass.mul_(0)

    if use_torch_cdist:
        idx = torch.argmin(torch.cdist(flat, embed, compute_mode="donot_use_mm_for_euclid_dist"), dim=1)
    else:
        idx = torch.from_numpy(vq_assign_c(flat.detach().numpy(), embed.detach().numpy())[0])
    quantized = embed[idx].reshape(shape)                 # layers.py:703 (codebook BEFORE the EMA update)

    if training:                                          # _update_ema, layers.py:636-663
        one_hot = F.one_hot(idx, num_classes=k).type_as(flat)
        new_cluster = one_hot.sum(dim=0) * world_size
        dw = (one_hot.T @ flat) * world_size
        cluster_size.mul_(decay).add_(new_cluster, alpha=1 - decay)
        embed_avg.mul_(decay).add_(dw, alpha=1 - decay)
        n = cluster_size.sum()
        smoothed = n * ((cluster_size + laplace_alpha) / (n + k * laplace_alpha))
        embed.copy_(embed_avg / smoothed.unsqueeze(-1))

    quantized = quantized.permute(0, 4, 1, 2, 3)          # non-contiguous view, layers.py:712
    idx = idx.reshape(shape[:-1])
    loss = commitment_cost * F.mse_loss(quantized, x)     # layers.py:716-717
    quantized = x + (quantized - x).detach()              # straight-through, layers.py:720
    return loss, quantized, idx


def embed_code(sd: Dict[str, Tensor], prefix: str, idx: Tensor) -> Tensor:
    return sd[prefix + "embed"][idx]                      # layers.py:633-634


# --------------------------------------------------------------------------------------
# blocks (layers.py:14-303, 591-597; evonorm.py:12-76)
# --------------------------------------------------------------------------------------
def _upsample2(x: Tensor) -> Tensor:
    return F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)   # layers.py:594


def _conv(x: Tensor, w: Tensor, bias: Optional[Tensor], stride: int, pad: int, circular: bool) -> Tensor:
    if pad and circular:                                  # nn.Conv3d(padding_mode='circular')
        x = F.pad(x, (pad,) * 6, mode="circular")
        pad = 0
    return F.conv3d(x, w, bias, stride=stride, padding=pad)


_MODE_GEOM = {"down": (4, 2, 1), "same": (3, 1, 1), "out": (3, 1, 1), "up": (3, 1, 1)}   # layers.py:124-132


def preact_block(sd: Dict[str, Tensor], p: str, x: Tensor, mode: str) -> Tensor:
    """PreActFixupResBlock.forward, layers.py:176-195 (circular padding, :109)."""
    k, stride, pad = _MODE_GEOM[mode]
    g = lambda name: sd[p + name]
    o = F.elu(x + g("bias1a"))
    o = F.conv3d(o + g("bias1b"), g("branch_conv1.weight"))
    o = F.elu(o + g("bias2a")) + g("bias2b")
    if mode == "up":
        o = _upsample2(o)
    o = _conv(o, g("branch_conv2.weight"), None, stride, pad, circular=True)
    o = F.elu(o + g("bias3a"))
    o = F.conv3d(o + g("bias3b"), g("branch_conv3.weight"))
    o = o * g("scale") + g("bias4")
    if (p + "skip_conv.weight") in sd:
        s = x + g("bias1c")
        if mode == "up":
            s = _upsample2(s)
        sk, ss = (2, 2) if mode == "down" else (1, 1)     # layers.py:167-168
        assert g("skip_conv.weight").shape[-1] == sk
        s = F.conv3d(s, g("skip_conv.weight"), None, stride=ss) + g("bias1d")
    else:
        s = x
    return o + s


def fixup_block(sd: Dict[str, Tensor], p: str, x: Tensor, mode: str) -> Tensor:
    """FixupResBlock.forward, layers.py:277-290 (zero padding)."""
    k, stride, pad = _MODE_GEOM[mode]
    g = lambda name: sd[p + name]
    o = x + g("bias1a")
    if mode == "up":
        o = _upsample2(o)
    o = _conv(o, g("branch_conv1.weight"), None, stride, pad, circular=False)
    o = F.elu(o + g("bias1b"))
    o = F.conv3d(o + g("bias2a"), g("branch_conv2.weight"), None, stride=1, padding=1)
    o = o * g("scale") + g("bias2b")
    s = _upsample2(x) if mode == "up" else x
    ss = 2 if mode == "down" else 1
    o = o + F.conv3d(s, g("skip_conv.weight"), g("skip_conv.bias"), stride=ss)
    return o if mode == "out" else F.elu(o)


def evonorm_s0(sd: Dict[str, Tensor], p: str, x: Tensor, eps: float = 1e-5) -> Tensor:
    """EvoNorm3DS0.forward, evonorm.py:12-26,34,70-76 (batch 1 only, like the reference)."""
    b, c = x.shape[:2]
    groups = max(c // 8, 1)                                # evonorm.py:8-9
    xg = x.reshape(b, groups, c // groups, *x.shape[2:])
    var = torch.var(xg, dim=tuple(range(2, x.dim() + 1)), keepdim=True)   # unbiased
    std = torch.sqrt(var + eps)
    std = std.expand(-1, -1, c // groups, *(-1 for _ in x.shape[2:])).reshape(1, c, *(1 for _ in x.shape[2:]))
    num = x * torch.sigmoid(x * sd[p + "v"])
    return num * sd[p + "gamma"] / std + sd[p + "beta"]


def evonorm_block(sd: Dict[str, Tensor], p: str, x: Tensor, mode: str) -> Tensor:
    """EvonormResBlock.forward, layers.py:75-83 (zero padding, convs with bias)."""
    if mode == "out":
        mode = "same"
    k, stride, pad = _MODE_GEOM[mode]
    g = lambda name: sd[p + name]
    o = F.conv3d(evonorm_s0(sd, p + "evonorm_1.", x), g("branch_conv1.weight"), g("branch_conv1.bias"))
    o = evonorm_s0(sd, p + "evonorm_2.", o)
    if mode == "up":
        o = _upsample2(o)
    o = _conv(o, g("branch_conv2.weight"), g("branch_conv2.bias"), stride, pad, circular=False)
    o = F.conv3d(evonorm_s0(sd, p + "evonorm_3.", o), g("branch_conv3.weight"), g("branch_conv3.bias"))
    if (p + "skip_conv.weight") in sd:
        s = _upsample2(x) if mode == "up" else x
        ss = 2 if mode == "down" else 1
        s = F.conv3d(s, g("skip_conv.weight"), g("skip_conv.bias"), stride=ss)
    else:
        s = x
    return o + s


_BLOCKS = {"pre-activation": preact_block, "regular": fixup_block, "evonorm": evonorm_block}


# --------------------------------------------------------------------------------------
# network assembly (layers.py:306-387, 463-588)
# --------------------------------------------------------------------------------------
def _down_block(sd, p, x, n_down, n_post, block):
    """DownBlock, layers.py:306-324: per downscale one 'down' block + n_post 'same' blocks."""
    j = 0
    for _ in range(n_down):
        x = block(sd, f"{p}layers.{j}.", x, "down"); j += 1
        for _ in range(n_post):
            x = block(sd, f"{p}layers.{j}.", x, "same"); j += 1
    return x


def _up_block(sd, p, x, n_up, n_post, block):
    """UpBlock, layers.py:327-354."""
    j = 0
    for _ in range(n_up):
        x = block(sd, f"{p}layers.{j}.", x, "up"); j += 1
        for _ in range(n_post):
            x = block(sd, f"{p}layers.{j}.", x, "same"); j += 1
    return x


def encoder2_forward(sd: Dict[str, Tensor], cfg: ModelConfig, x: Tensor, training: bool = False,
                     p: str = "encoder.", world_size: int = 1, use_torch_cdist: bool = False,
                     collect: Optional[dict] = None) -> List[Tuple[Tensor, Tensor, Tensor]]:
    """Encoder2.forward, layers.py:577-588.  Returns [(loss, quantized, idx)] ordered
    bottom (largest grid) -> top, i.e. already in the order ``reversed(quantizations)`` yields."""
    block = _BLOCKS[cfg.block_type]
    n_enc, n_down = cfg.n_bottleneck_blocks, cfg.n_downscales_per_bottleneck
    down = F.conv3d(x, sd[p + "parse_input.weight"], sd[p + "parse_input.bias"])
    downs = []
    for i in range(n_enc):
        down = _down_block(sd, f"{p}down.{i}.", down, n_down, cfg.n_post_downscale_blocks, block)
        downs.append(down)
    aux = None
    out: List = [None] * n_enc
    for i in reversed(range(n_enc)):
        h = downs[i]
        if i != n_enc - 1:       # has_aux (layers.py:367,384-385)
            up = _up_block(sd, f"{p}pre_quantize_cond.{i}.upsample.", aux, n_down, cfg.n_post_upscale_blocks, block)
            h = F.conv3d(torch.cat([h, up], dim=1), sd[f"{p}pre_quantize_cond.{i}.proj.weight"],
                         sd[f"{p}pre_quantize_cond.{i}.proj.bias"])
        h = block(sd, f"{p}pre_quantize_cond.{i}.pre_q.", h, "same")
        for j in range(cfg.n_pre_quantization_blocks):
            h = block(sd, f"{p}pre_quantize.{i}.{j}.", h, "same")
        if collect is not None:
            collect[f"latent_{i}"] = h.detach().clone()
        loss, q, idx = quantizer_forward(sd, f"{p}quantize.{i}.", h, training, world_size=world_size,
                                         use_torch_cdist=use_torch_cdist)
        aux = q
        out[i] = (loss, q, idx)
    return out


def decoder_forward(sd: Dict[str, Tensor], cfg: ModelConfig, quantizations: Sequence[Tensor],
                    p: str = "decoder.") -> Tensor:
    """Decoder.forward, layers.py:510-517; ``quantizations`` ordered bottom -> top."""
    block = _BLOCKS[cfg.block_type]
    n_enc, n_up = cfg.n_bottleneck_blocks, cfg.n_downscales_per_bottleneck
    n_proj = n_enc - 1
    out = None
    for i, level in enumerate(reversed(range(n_enc))):
        q = quantizations[level]
        if i == 0:
            out = q
        else:   # self.proj[-i]
            out = F.conv3d(torch.cat([q, out], dim=1), sd[f"{p}proj.{n_proj - i}.weight"],
                           sd[f"{p}proj.{n_proj - i}.bias"])
        for j in range(cfg.n_post_quantization_blocks):
            out = block(sd, f"{p}up.{level}.{j}.", out, "same")
        out = _up_block(sd, f"{p}up.{level}.{cfg.n_post_quantization_blocks}.", out, n_up,
                        cfg.n_post_upscale_blocks, block)
    return F.conv3d(out, sd[p + "out.weight"], sd[p + "out.bias"])


def vqvae_forward(sd: Dict[str, Tensor], cfg: ModelConfig, x: Tensor, training: bool = False,
                  use_torch_cdist: bool = False, collect: Optional[dict] = None):
    """VQVAE.forward, model.py:79-83: (decoded, (losses, quantizations, idx)) bottom -> top."""
    levels = encoder2_forward(sd, cfg, x, training, use_torch_cdist=use_torch_cdist, collect=collect)
    losses, quants, idxs = zip(*levels)
    decoded = decoder_forward(sd, cfg, quants)
    return decoded, (losses, quants, idxs)


# --------------------------------------------------------------------------------------
# loss epilogue (model.py:115-160) and the centre-cylinder crop (utils/load_nrrd_dataset.py:288-300)
# --------------------------------------------------------------------------------------
def center_cylinder_mask(h: int, w: int) -> Tensor:
    """Boolean (h, w) mask of the inscribed circle used by ExtractCenterCylinder:
    centre (h/2, w/2), radius min(h, w)/2, sqrt(dx^2+dy^2) <= radius in float64."""
    ys = np.arange(h, dtype=np.float64)[:, None]
    xs = np.arange(w, dtype=np.float64)[None, :]
    dist = np.sqrt((ys - h / 2) ** 2 + (xs - w / 2) ** 2)
    return torch.from_numpy(dist <= min(h, w) / 2)


def huber_epilogue(decoded: Tensor, x: Tensor, num_valid_slices: Sequence[int],
                   commitment: Sequence[Tensor], cylinder: bool = False) -> Tuple[Tensor, Tensor]:
    """loss = smooth_l1(ELU(decoded) masked, x).mean() + sum(commitment), model.py:120-155."""
    loc = F.elu(decoded)
    mask = torch.zeros_like(x, dtype=torch.bool)
    for n, m in zip(num_valid_slices, mask):
        m[..., n:] = True
    loc = torch.masked_fill(loc, mask, 0.0)
    if cylinder:
        keep = center_cylinder_mask(x.shape[2], x.shape[3])
        loc, x = loc[:, :, keep], x[:, :, keep]
    recon = F.smooth_l1_loss(loc, x, reduction="none").mean()
    return recon + sum(commitment), recon


def validation_log(decoded: Tensor, x: Tensor, num_valid_slices: Sequence[int], cylinder: bool = False,
                   data_range: float = 4.0) -> Dict[str, Tensor]:
    """What VQVAE.loc_metric logs beside the loss (model.py:120-149): utils/logging_helpers.py:4-15 over the unreduced loss
    and over the masked reconstruction (min / max / mean / median / std), and
    metrics/evaluate.py:18-24 (nmse = ||pred - orig||^2 / ||orig||^2, psnr = 10 log10(data_range^2 / mse), data_range 4
    as model.py:25)."""
    loc = F.elu(decoded)
    mask = torch.zeros_like(x, dtype=torch.bool)
    for n, m in zip(num_valid_slices, mask):
        m[..., n:] = True
    loc = torch.masked_fill(loc, mask, 0.0)
    if cylinder:
        keep = center_cylinder_mask(x.shape[2], x.shape[3])
        loc, x = loc[:, :, keep], x[:, :, keep]
    unreduced = F.smooth_l1_loss(loc, x, reduction="none")
    log = {}
    for name, t in (("recon_loss", unreduced), ("loc", loc)):
        log.update({f"{name}_min": t.min(), f"{name}_max": t.max(), f"{name}_mean": t.mean(), f"{name}_median": t.median(),
                    f"{name}_std": t.std()})
    log["nmse"] = torch.norm(loc - x) ** 2 / torch.norm(x) ** 2
    log["psnr"] = 10 * torch.log10((data_range ** 2) / F.mse_loss(loc, x))
    return log


# --------------------------------------------------------------------------------------
# deterministic synthetic inputs
# --------------------------------------------------------------------------------------
def synthetic_volume(shape: Sequence[int], seed: int = 42) -> Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * 4.5 - 0.5
