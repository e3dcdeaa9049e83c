"""B200-native drop-in for the reference's vqvae/layers.py.

Same public classes, constructor signatures, attribute / state_dict names and return
contracts as sara-nl/3D-VQ-VAE-2 `vqvae/layers.py` (cited per class below), so that
`vqvae/model.py`, `train.py`, `extract_embeddings.py` and `decode_embeddings.py` can import
this module instead.  The modules are parameter containers + launch logic only: every
number is produced by the sm_100a kernels of libvqvae3d_b200.so through `_ops` (C ABI,
include/vqvae3d_b200.h).  There is no PyTorch/CPU fallback; tensors must live on CUDA.

Inference (autograd off) runs each block / stack of blocks as one fused launch.  When autograd is
recording, the blocks are composed from the differentiable generic ops of `_ops` whose backward runs
the library's dgrad / wgrad / scalar-gradient kernels (csrc/backward_kernels.cu), so `loss.backward()`
works on every module here.
"""
from __future__ import annotations

from itertools import chain
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import _ops

_MODE_ID = {"same": 0, "out": 0, "down": 1, "up": 2}


def ops() -> "_ops.Ops":
    return _ops.default()


def _recording(x, *params) -> bool:
    """True when autograd is recording and something upstream wants a gradient: the fused inference kernels
    (one launch per block / stack) are bypassed and the block is composed from the differentiable generic ops."""
    if not torch.is_grad_enabled():
        return False
    return (x is not None and x.requires_grad) or any(p is not None and p.requires_grad for p in params)


class Conv3d(nn.Conv3d):
    """Parameter container with nn.Conv3d's constructor/initialisation (so that seeds
    reproduce the reference's weights); forward runs vq3d_conv3d.  Used for
    parse_input / proj / out (layers.py:377,490,508,535)."""

    def forward(self, input: torch.Tensor, input2: Optional[torch.Tensor] = None) -> torch.Tensor:
        return ops().conv3d(input, self.weight, x2=input2, bias=self.bias, stride=self.stride[0],
                            pad=self.padding[0] if not isinstance(self.padding, str) else 0,
                            circular=self.padding_mode == "circular")


class ResizeConv3D(Conv3d):
    """layers.py:591-597: trilinear x2 upsample (align_corners=False), then the convolution."""

    def forward(self, input: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        return super().forward(ops().upsample2x(input))


def _geometry(mode: str):
    """(conv class, k, stride, pad) per mode, layers.py:124-132."""
    if mode == "down":
        return Conv3d, 4, 2, 1
    if mode == "up":
        return ResizeConv3D, 3, 1, 1
    return Conv3d, 3, 1, 1


class PreActFixupResBlock(nn.Module):
    """layers.py:102-216.  1x1 -> k^3 (circular padding) -> 1x1 bottleneck with scalar Fixup
    biases/scale and ELU pre-activations; optional skip conv when the shape changes."""

    def __init__(self, in_channels, out_channels, mode, activation=nn.ELU, bottleneck_divisor=2):
        super().__init__()
        assert mode in ("down", "same", "up", "out")
        if activation is not nn.ELU:
            raise NotImplementedError("only nn.ELU is fused into the kernels (the reference never uses another)")
        self.mode = mode
        branch = max(max(in_channels, out_channels) // bottleneck_divisor, 1)
        self.activation = activation()
        for name in ("bias1a", "bias1b", "bias2a", "bias2b", "bias3a", "bias3b", "bias4"):
            setattr(self, name, nn.Parameter(torch.zeros(1)))
        self.scale = nn.Parameter(torch.ones(1))
        conv, k, stride, pad = _geometry(mode)
        # construction order = the reference's (RNG stream parity): conv1, conv2, conv3, skip
        self.branch_conv1 = Conv3d(in_channels, branch, kernel_size=1, bias=False)
        self.branch_conv2 = conv(branch, branch, kernel_size=k, stride=stride, padding=pad, bias=False,
                                 padding_mode="circular")
        self.branch_conv3 = Conv3d(branch, out_channels, kernel_size=1, bias=False)
        if not (mode in ("same", "out") and in_channels == out_channels):
            self.bias1c = nn.Parameter(torch.zeros(1))
            self.bias1d = nn.Parameter(torch.zeros(1))
            sk = 2 if mode == "down" else 1
            self.skip_conv = conv(in_channels, out_channels, kernel_size=sk, stride=sk, padding=0, bias=False)
        else:
            self.skip_conv = None

    def _params(self):
        return [p for p in self.parameters(recurse=True)]

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if _recording(input, *self._params()):
            o = ops()
            if self.mode in ("same", "out") and self.skip_conv is None and o.preact_same_backward_workspace(input, self) > 0:
                # 'same' block: ONE fused forward launch that keeps only the input + the fused 3-launch backward
                names = _ops._PreactSameFn._PARAMS
                return _ops._PreactSameFn.apply(o, self, input, *(self.get_parameter(n) for n in names))
            return self.forward_composed(input)          # training: differentiable generic ops
        o = ops()
        # wide down blocks: the fused SIMT kernel keeps all branch channels of the (2t+2)^3 input window in
        # shared memory, which leaves it a sliver of a tile at 16 channels; the composed tensor-core path wins
        wide_down = self.mode == "down" and self.branch_conv1.weight.shape[0] >= 16 and o.precision == "bf16"
        if not wide_down:
            y = o.preact_block(input, self, _MODE_ID[self.mode])      # one fused launch when covered
            if y is not None:
                return y
        return self.forward_composed(input)

    def forward_composed(self, x: torch.Tensor) -> torch.Tensor:
        """The same block as 3-5 launches of the generic kernels (any channel count)."""
        o = ops()
        t = o.conv3d(x, self.branch_conv1.weight, pre_act=True, pre_a=self.bias1a, pre_b=self.bias1b)
        if self.mode == "up":
            t = o.upsample2x(t, pre_act=True, pre_a=self.bias2a, pre_b=self.bias2b)
            t = o.conv3d(t, self.branch_conv2.weight, pad=1, circular=True)
        else:
            k = self.branch_conv2.weight.shape[2]
            t = o.conv3d(t, self.branch_conv2.weight, stride=2 if self.mode == "down" else 1, pad=1, circular=True,
                         pre_act=True, pre_a=self.bias2a, pre_b=self.bias2b)
            assert k == (4 if self.mode == "down" else 3)
        if self.skip_conv is not None:
            if self.mode == "up":
                s = o.conv3d(o.upsample2x(x, pre_b=self.bias1c), self.skip_conv.weight, post_b=self.bias1d)
            else:
                s = o.conv3d(x, self.skip_conv.weight, stride=2 if self.mode == "down" else 1, pre_b=self.bias1c,
                             post_b=self.bias1d)
        else:
            s = x
        return o.conv3d(t, self.branch_conv3.weight, pre_act=True, pre_a=self.bias3a, pre_b=self.bias3b,
                        post_scale=self.scale, post_b=self.bias4, residual=s)

    @torch.no_grad()
    def initialize_weights(self, num_layers):
        """Fixup initialisation, layers.py:197-216."""
        w = self.branch_conv1.weight
        nn.init.normal_(w, mean=0, std=float(np.sqrt(2 / (w.shape[0] * np.prod(w.shape[2:]))) * num_layers ** (-0.5)))
        nn.init.kaiming_normal_(self.branch_conv2.weight)
        nn.init.constant_(self.branch_conv3.weight, val=0)
        if self.skip_conv is not None:
            nn.init.xavier_normal_(self.skip_conv.weight)


class FixupResBlock(nn.Module):
    """layers.py:219-303 (`--block-type regular`): k^3 conv -> ELU -> 3^3 conv, zero padding,
    skip conv with bias always present, trailing ELU unless mode == 'out'."""

    def __init__(self, in_channels, out_channels, mode, activation=nn.ELU):
        super().__init__()
        assert mode in ("down", "same", "up", "out")
        if activation is not nn.ELU:
            raise NotImplementedError("only nn.ELU is fused into the kernels")
        self.mode = mode
        self.activation = activation()
        for name in ("bias1a", "bias1b", "bias2a", "bias2b"):
            setattr(self, name, nn.Parameter(torch.zeros(1)))
        self.scale = nn.Parameter(torch.ones(1))
        conv, k, stride, pad = _geometry(mode)
        self.branch_conv1 = conv(in_channels, out_channels, kernel_size=k, stride=stride, padding=pad, bias=False)
        sk = 2 if mode == "down" else 1
        self.skip_conv = conv(in_channels, out_channels, kernel_size=sk, stride=sk, padding=0, bias=True)
        self.branch_conv2 = Conv3d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        o = ops()
        stride = 2 if self.mode == "down" else 1
        if self.mode == "up":
            t = o.conv3d(o.upsample2x(input, pre_b=self.bias1a), self.branch_conv1.weight, pad=1)
            s = o.conv3d(o.upsample2x(input), self.skip_conv.weight, bias=self.skip_conv.bias)
        else:
            t = o.conv3d(input, self.branch_conv1.weight, stride=stride, pad=1, pre_b=self.bias1a)
            s = o.conv3d(input, self.skip_conv.weight, bias=self.skip_conv.bias, stride=stride)
        return o.conv3d(t, self.branch_conv2.weight, pad=1, pre_act=True, pre_a=self.bias1b, pre_b=self.bias2a,
                        post_scale=self.scale, post_b=self.bias2b, residual=s, post_act=self.mode != "out")

    def initialize_weights(self, num_layers):
        """layers.py:292-303."""
        w = self.branch_conv1.weight
        nn.init.normal_(w, mean=0, std=float(np.sqrt(2 / (w.shape[0] * np.prod(w.shape[2:]))) * num_layers ** (-0.5)))
        nn.init.constant_(self.branch_conv2.weight, val=0)
        nn.init.kaiming_normal_(self.skip_conv.weight)
        nn.init.constant_(self.skip_conv.bias, val=0)


class EvonormResBlock(nn.Module):
    """layers.py:14-98 (`--block-type evonorm`): EvoNorm3D-S0 -> conv bottleneck with biases and zero
    padding; composed from the EvoNorm kernels and the generic / tensor-core convolutions."""

    def __init__(self, in_channels, out_channels, mode, bottleneck_divisor=4):
        super().__init__()
        from .evonorm import EvoNorm3DS0
        assert mode in ("down", "same", "up", "out")
        if mode == "out":
            mode = "same"
        self.mode = mode
        branch = max(max(in_channels, out_channels) // bottleneck_divisor, 1)
        conv, k, stride, pad = _geometry(mode)
        self.evonorm_1 = EvoNorm3DS0(in_channels)
        self.branch_conv1 = Conv3d(in_channels, branch, kernel_size=1)
        self.evonorm_2 = EvoNorm3DS0(branch)
        self.branch_conv2 = conv(branch, branch, kernel_size=k, stride=stride, padding=pad)
        self.evonorm_3 = EvoNorm3DS0(branch)
        self.branch_conv3 = Conv3d(branch, out_channels, kernel_size=1)
        sk = 2 if mode == "down" else 1
        self.skip_conv = conv(in_channels, out_channels, kernel_size=sk, stride=sk, padding=0) \
            if not (mode in ("same", "out") and in_channels == out_channels) else None
        self.initialize_weights()

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        """layers.py:84-91: conv1(EN1(x)) -> conv2(EN2(.)) -> conv3(EN3(.)) + (skip(x) | x); zero padding, conv biases."""
        o = ops()
        stride = 2 if self.mode == "down" else 1
        t = o.conv3d(self.evonorm_1(input), self.branch_conv1.weight, bias=self.branch_conv1.bias)
        t = self.evonorm_2(t)
        if self.mode == "up":
            t = o.conv3d(o.upsample2x(t), self.branch_conv2.weight, bias=self.branch_conv2.bias, pad=1)
        else:
            t = o.conv3d(t, self.branch_conv2.weight, bias=self.branch_conv2.bias, stride=stride, pad=1)
        if self.skip_conv is None:
            s = input
        elif self.mode == "up":
            s = o.conv3d(o.upsample2x(input), self.skip_conv.weight, bias=self.skip_conv.bias)
        else:
            s = o.conv3d(input, self.skip_conv.weight, bias=self.skip_conv.bias, stride=stride)
        return o.conv3d(self.evonorm_3(t), self.branch_conv3.weight, bias=self.branch_conv3.bias, residual=s)

    @torch.no_grad()
    def initialize_weights(self):
        for w in (self.branch_conv1.weight, self.branch_conv2.weight, self.branch_conv3.weight):
            nn.init.kaiming_normal_(w)
        if self.skip_conv is not None:
            nn.init.xavier_normal_(self.skip_conv.weight)
            nn.init.zeros_(self.skip_conv.bias)


class BlockSequence(nn.Sequential):
    """nn.Sequential with the reference's key names whose forward hands maximal runs of
    equal-shape 'same' PreAct blocks to one vq3d_preact_stack call (the 50/150-deep stacks of
    layers.py:492-494,566-569 are launch-latency bound when run block by block)."""

    def forward(self, x: torch.Tensor, tail: Optional[nn.Module] = None, pre: Optional[nn.Module] = None) -> torch.Tensor:  # type: ignore[override]
        """tail: a 1x1 convolution applied after the last block (the decoder's `out`, layers.py:516); pre: one
        applied before the first block (the encoder's parse_input, layers.py:578).  Each is fused into the
        neighbouring block's kernel where one covers the shape, otherwise run on its own."""
        mods = list(self)
        if _recording(x, *self.parameters(), *(pre.parameters() if pre is not None else ()), *(tail.parameters() if tail is not None else ())):
            x = pre(x) if pre is not None else x
            for m in mods:
                x = m(x)
            return tail(x) if tail is not None else x
        i = 0
        done_tail = tail is None
        if pre is not None:
            y = None
            m0 = mods[0] if mods else None
            if isinstance(m0, PreActFixupResBlock) and m0.mode == "down":
                y = ops().preact_block(x, m0, _MODE_ID[m0.mode], pre=pre)
            if y is not None:
                x, i = y, 1
            else:
                x = pre(x)
        while i < len(mods):
            m = mods[i]
            j = i
            if _stackable(m):
                while j + 1 < len(mods) and _stackable(mods[j + 1]) and _same_shape(m, mods[j + 1]):
                    j += 1
                if tail is not None and j == len(mods) - 1:
                    y = ops().preact_stack(x, mods[i:j + 1], tail=tail)
                    if y is not None:
                        return y
                y = ops().preact_stack(x, mods[i:j + 1])
                if y is not None:
                    x = y
                    i = j + 1
                    continue
            if tail is not None and i == len(mods) - 1 and isinstance(m, UpBlock):
                return m(x, tail=tail)
            x = m(x)
            i += 1
        return x if done_tail else tail(x)


def _stackable(m) -> bool:
    return isinstance(m, PreActFixupResBlock) and m.mode in ("same", "out") and m.skip_conv is None


def _same_shape(a, b) -> bool:
    return a.branch_conv1.weight.shape == b.branch_conv1.weight.shape


class DownBlock(nn.Module):
    """layers.py:306-324."""

    def __init__(self, in_channels, n_down=2, resblock=FixupResBlock, n_post_downscale_blocks=0):
        super().__init__()
        self.layers = BlockSequence(*chain.from_iterable(
            (resblock(in_channels * 2 ** i, in_channels * 2 ** (i + 1), mode="down"),
             *(resblock(in_channels * 2 ** (i + 1), in_channels * 2 ** (i + 1), mode="same")
               for _ in range(n_post_downscale_blocks)))
            for i in range(n_down)))

    def forward(self, data, pre=None):
        return self.layers(data, pre=pre)


class UpBlock(nn.Module):
    """layers.py:327-354."""

    def __init__(self, in_channels, out_channels, aux_channels=0, n_up=2, mode="encoder", resblock=FixupResBlock,
                 n_post_upscale_blocks=0):
        super().__init__()
        assert mode in ("encoder", "decoder")
        self.layers = BlockSequence(*chain.from_iterable(
            (resblock(in_channels if i == n_up - 1 else out_channels * (2 ** (i + 1)), out_channels * (2 ** i), mode="up"),
             *(resblock(out_channels * (2 ** i), out_channels * (2 ** i), mode="same")
               for _ in range(n_post_upscale_blocks)))
            for i in range(n_up - 1, -1, -1)))

    def forward(self, data, tail=None):
        return self.layers(data, tail=tail)


class PreQuantizationConditioning(nn.Module):
    """layers.py:357-387.  cat([data, upsample(aux)]) + 1x1 proj runs as ONE two-source
    pointwise kernel; the concatenated tensor is never written."""

    def __init__(self, in_channels, out_channels, n_up=2, resblock=FixupResBlock, n_post_upscale_blocks=0):
        super().__init__()
        self.has_aux = in_channels - out_channels * 8 != 0
        if self.has_aux:
            self.upsample = UpBlock(out_channels * 2 ** n_up, out_channels, n_up=n_up, resblock=resblock,
                                    n_post_upscale_blocks=n_post_upscale_blocks)
            self.proj = Conv3d(in_channels, in_channels, kernel_size=1)
        self.pre_q = resblock(in_channels, out_channels, mode="same")

    def forward(self, data, auxilary=None):
        assert self.has_aux is (auxilary is not None)
        if self.has_aux:
            data = self.proj(data, self.upsample(auxilary))
        return self.pre_q(data)


class Decoder(nn.Module):
    """layers.py:463-517.  forward(quantizations ordered bottom -> top) -> (B, out, H, W, Z)."""

    def __init__(self, out_channels, base_network_channels, n_enc=3, n_up_per_enc=2, n_post_q_blocks=0,
                 n_post_upscale_blocks=0, resblock=FixupResBlock):
        super().__init__()
        self.up = nn.ModuleList()
        self.proj = nn.ModuleList()
        after = base_network_channels
        for i in range(n_enc):
            before = after * 2 ** n_up_per_enc
            assert before % 8 == 0
            embedding_dim = before // 8
            in_channels = embedding_dim + (before if i != n_enc - 1 else 0)
            if i != n_enc - 1:
                self.proj.append(Conv3d(in_channels, in_channels, kernel_size=1))
            self.up.append(BlockSequence(
                *(resblock(in_channels, in_channels, mode="same") for _ in range(n_post_q_blocks)),
                UpBlock(in_channels=in_channels, out_channels=after, n_up=n_up_per_enc, mode="decoder",
                        resblock=resblock, n_post_upscale_blocks=n_post_upscale_blocks)))
            after = before
        self.out = Conv3d(base_network_channels, out_channels, kernel_size=1)

    def forward(self, quantizations: Sequence[torch.Tensor]) -> torch.Tensor:
        out = None
        for i, (q, up) in enumerate(reversed(list(zip(quantizations, self.up)))):
            out = q if i == 0 else self.proj[-i](q, out)      # cat([q, out]) folded into the 1x1
            last = i == len(self.up) - 1
            out = up(out, tail=self.out) if last else up(out)  # the final 1x1 `out` conv rides on the last block
        return out


class Encoder2(nn.Module):
    """layers.py:519-588.  forward -> iterator of (loss, quantized, idx) ordered bottom -> top."""

    def __init__(self, in_channels, base_network_channels, num_embeddings: List[int], n_enc=3, n_down_per_enc=2,
                 n_pre_q_blocks=0, n_post_upscale_blocks=0, n_post_downscale_blocks=0, resblock=FixupResBlock):
        super().__init__()
        self.parse_input = Conv3d(in_channels, base_network_channels, kernel_size=1)
        before = base_network_channels
        self.down, self.pre_quantize, self.pre_quantize_cond, self.quantize = (nn.ModuleList() for _ in range(4))
        for i in range(n_enc):
            after = before * 2 ** n_down_per_enc
            self.down.append(DownBlock(before, n_down_per_enc, resblock=resblock,
                                       n_post_downscale_blocks=n_post_downscale_blocks))
            assert after % 8 == 0
            embedding_dim = after // 8
            self.pre_quantize_cond.append(PreQuantizationConditioning(
                in_channels=after + (embedding_dim if i != n_enc - 1 else 0), out_channels=embedding_dim,
                n_up=n_down_per_enc, resblock=resblock, n_post_upscale_blocks=n_post_upscale_blocks))
            self.pre_quantize.append(BlockSequence(*(resblock(embedding_dim, embedding_dim, mode="same")
                                                     for _ in range(n_pre_q_blocks))))
            self.quantize.append(Quantizer(num_embeddings=num_embeddings[i], embedding_dim=embedding_dim,
                                           commitment_cost=0.1))
            before = after

    def forward(self, data: torch.Tensor):
        down = data
        pyramid = []
        for i, block in enumerate(self.down):
            down = block(down, pre=self.parse_input) if i == 0 else block(down)     # parse_input rides on the first block
            pyramid.append(down)
        aux = None
        levels = []
        for feat, pre_q, cond, quant in reversed(list(zip(pyramid, self.pre_quantize, self.pre_quantize_cond, self.quantize))):
            triple = quant(pre_q(cond(feat, aux)))
            levels.append(triple)
            aux = triple[1]
        return reversed(levels)


class _StraightThroughVQ(torch.autograd.Function):
    """Autograd edge of Quantizer.forward (layers.py:716-720): identity gradient to the
    latents plus the commitment-loss term; nothing flows to the codebook."""

    @staticmethod
    def forward(ctx, inputs, quantizer):
        loss, quant, idx = quantizer._forward_impl(inputs)
        ctx.save_for_backward(inputs, quant)
        ctx.cc = quantizer.commitment_cost
        ctx.mark_non_differentiable(idx)
        return loss, quant, idx

    @staticmethod
    def backward(ctx, g_loss, g_quant, _g_idx):
        x, q = ctx.saved_tensors
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=x.device)
        gx = ops().vq_backward(g_quant, g_loss.float().reshape(()), x, q, ctx.cc)
        return gx, None


class Quantizer(nn.Module):
    """EMA vector quantizer, layers.py:602-728.  forward(x: (B, D, H, W, Z)) ->
    (loss 0-d fp32, quantized (B, D, H, W, Z) fp32 with straight-through grad, idx (B, H, W, Z) int64);
    in training mode the buffers embed / embed_avg / cluster_size / first_pass are updated
    in place exactly as the reference does (data-dependent init on the first pass, then EMA)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, commitment_cost: float, decay=0.99, laplace_alpha=1e-5):
        super().__init__()
        embed = torch.randn(num_embeddings, embedding_dim)          # CPU default generator, like layers.py:614
        self.register_buffer("embed", embed)
        self.register_buffer("embed_avg", embed.clone())
        self.register_buffer("cluster_size", torch.zeros(num_embeddings))
        self.register_buffer("first_pass", torch.as_tensor(1))
        self.commitment_cost = commitment_cost
        self.decay = decay
        self.laplace_alpha = laplace_alpha
        self.embedding_dim = embedding_dim
        self.num_embeddings = num_embeddings
        self._ema_ready = False      # host mirror of "first_pass == 0": once true, training forwards never read the device flag
                                     # again (no host sync per step; lets a whole training step be captured in a CUDA graph)

    def _load_from_state_dict(self, *args, **kwargs):
        self._ema_ready = False      # a loaded checkpoint may carry first_pass = 1
        return super()._load_from_state_dict(*args, **kwargs)

    def embed_code(self, embed_idx: torch.Tensor) -> torch.Tensor:
        return ops().embed_code(embed_idx, self.embed)

    def _world(self) -> int:
        d = torch.distributed
        return d.get_world_size() if d.is_available() and d.is_initialized() else 1

    def _forward_impl(self, inputs: torch.Tensor):
        o = ops()
        x = inputs.detach()
        if x.dtype != torch.float32:
            x = x.float()                                            # layers.py:687 (quantizer is always fp32)
        if x.dim() != 5 or x.shape[1] != self.embedding_dim:
            raise RuntimeError(f"Quantizer expects (B, {self.embedding_dim}, H, W, Z), got {tuple(x.shape)}")
        world = self._world()
        n_vectors = x.numel() // self.embedding_dim
        if self.training and not self._ema_ready:
            if bool(self.first_pass):                                # _init_ema, layers.py:665-683
                meanstd = o.vq_init_stats(x)
                if world > 1:
                    torch.distributed.all_reduce(meanstd)
                    meanstd /= world
                o.vq_init_apply(meanstd, n_vectors * world, self.embed, self.embed_avg, self.cluster_size, self.first_pass)
            self._ema_ready = True
        quant, idx, sqerr, stats = o.vq_assign(x, self.embed, want_stats=self.training)
        loss = o.vq_loss(sqerr, self.commitment_cost, x.numel())
        if self.training:                                            # _update_ema, layers.py:636-663
            if world > 1:   # ONE flat all-reduce of [counts | dw] instead of the reference's two
                torch.distributed.all_reduce(stats)
            k, d = self.num_embeddings, self.embedding_dim
            o.vq_ema_update(stats[:k], stats[k:].view(k, d), self.decay, self.laplace_alpha, self.cluster_size,
                            self.embed_avg, self.embed)
        return loss, quant, idx

    def forward(self, inputs: torch.Tensor):
        if torch.is_grad_enabled() and inputs.requires_grad:
            return _StraightThroughVQ.apply(inputs, self)
        return self._forward_impl(inputs)
