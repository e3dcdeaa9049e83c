"""Tensor-level wrappers over the C ABI (include/vqvae3d_b200.h).

PyTorch is used only for device memory and the current stream; all arithmetic happens in
libvqvae3d_b200.so.  Every wrapper insists on CUDA fp32 tensors -- there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _cabi

Tensor = torch.Tensor


class Ops:
    """Product instance: `default()`.  (tests/emu subclasses this to drive the host
    emulator build of the same kernels; the product never does.)"""

    def __init__(self, lib: Optional[C.CDLL] = None):
        self._lib = lib
        # "bf16": GEMM-shaped convolutions (C_in*k^3 >= 64, C_out >= 16) run on the tcgen05 tensor cores
        # with bf16 operands / fp32 accumulation; "fp32": every convolution on the exact fp32 SIMT kernels
        self.precision = "bf16"
        self.launches = 0            # kernels enqueued through this object (bench.py's gpu_launches)
        self.profile = None          # list -> (name, tag, algorithmic bytes, flops, ev0, ev1) per C call
        self.tc_min_voxels = int(os.environ.get("VQ3D_TC_MIN_VOXELS", "0"))   # convs with fewer output voxels stay on the fp32 SIMT kernel
        # merged-tap tensor-core kernel for the thin 512^3 / 256^3 blocks: correct, but measured SLOWER than the SIMT row kernel
        # (0.72 vs 0.52 ms per 4->2->4 block and volume: the im2col copy saturates the shared-memory pipe), so off by default
        self.thin_tc = os.environ.get("VQ3D_THIN_TC", "0") == "1"
        self.up_tc = os.environ.get("VQ3D_UP_TC", "1") == "1"      # wide 'up' blocks on the tensor-core kernel (bf16 mode)
        # 'same' blocks: fused forward that keeps only x + fused 3-launch backward (vq3d_preact_same_backward).  Exact and 2.4x fewer
        # kernels per step, but MEASURED SLOWER than the composed path inside a CUDA graph (C2 188 vs 152 ms, C3 310 vs 217 ms, and
        # never faster when restricted to small tensors: DESIGN.md 8), so it is off by default
        self.fused_block_bwd = os.environ.get("VQ3D_FUSED_BLOCK_BWD", "0") == "1"
        self.fused_block_bwd_max_voxels = int(os.environ.get("VQ3D_FUSED_BLOCK_BWD_MAXVOX", "300000"))
        self.fused_pointwise_bwd = os.environ.get("VQ3D_FUSED_PW_BWD", "1") == "1"   # k1 convolutions: one fused backward launch
        self.dgrad_as_forward = True # input gradients of stride-1 same convolutions run as forward convolutions
        # parameter gradients are added straight into an existing contiguous fp32 `.grad` (the flat buffer of FusedAdamAMSGrad):
        # the kernels accumulate with atomics anyway, so the zero-filled temporary, the per-scalar clones and autograd's
        # AccumulateGrad additions (~17 one-element kernels per residual block) disappear
        self.grad_inplace = os.environ.get("VQ3D_GRAD_INPLACE", "1") != "0"
        # scalar-gradient scratch of the convolution backwards ([4] floats each, zero on entry): slots of one pool that
        # `begin_step()` zeroes with a single launch; without begin_step (or when the pool runs out) each call allocates its own
        self._gscal_pool, self._gscal_next = {}, 0
        self.vq_tensor_cores = True  # large quantizer problems: tcgen05 candidate pass + exact re-rank (index-identical)
        self._ws = {}                # (device, stream) -> uint8 workspace of the kernels that need scratch (grown on demand)

    # -- plumbing ---------------------------------------------------------------------
    @property
    def lib(self) -> C.CDLL:
        if self._lib is None:
            self._lib = _cabi.lib()
        return self._lib

    def stream(self) -> int:
        return torch.cuda.current_stream().cuda_stream

    def _t(self, t: Optional[Tensor], dtype=torch.float32) -> Optional[Tensor]:
        if t is None:
            return None
        if not t.is_cuda:
            raise RuntimeError("3d-vq-vae-2_b200 runs on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")
        if t.dtype != dtype:
            raise RuntimeError(f"expected {dtype}, got {t.dtype}")
        return t if t.is_contiguous() else t.contiguous()

    @staticmethod
    def _p(t: Optional[Tensor]) -> Optional[int]:
        return None if t is None else t.data_ptr()

    def _call(self, name, cfunc, args, kernels=1, nbytes=0, flops=0, tag="", allow_unsupported=False) -> bool:
        """One C-ABI call.  With `profile` set, brackets it with CUDA events on the launching stream."""
        prof = self.profile
        if prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        ok = self._check(cfunc(*args), allow_unsupported)
        if ok:
            self.launches += kernels
            if prof is not None:
                e1.record()
                prof.append((name, tag, nbytes, flops, e0, e1))
        return ok

    def _workspace(self, nbytes: int, device) -> Tensor:
        """Scratch for kernels that need one.  Work on a stream is ordered and a workspace's contents never outlive
        a call, so one buffer per (device, stream) is enough; it only grows."""
        key = (device, self.stream())
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def begin_step(self, device, slots: int = 4096) -> None:
        """Called once per training step before the backward pass (vqvae.parallel.training_step): zeroes the pool of
        scalar-gradient scratch slots with ONE launch, so that the ~1 000 convolution backwards of a step do not each
        launch a fill kernel for their four floats.  Slots are handed out in call order and are valid until the next begin_step."""
        key = (device, self.stream())
        pool = self._gscal_pool.get(key)
        if pool is None or pool.shape[0] < slots:
            pool = torch.zeros(slots, 4, dtype=torch.float32, device=device)
            self._gscal_pool = {key: pool}
        else:
            pool.zero_()
        self._gscal_next = 0

    def _gscal(self, device) -> Tensor:
        pool = self._gscal_pool.get((device, self.stream()))
        if pool is not None and self._gscal_next < pool.shape[0]:
            self._gscal_next += 1
            return pool[self._gscal_next - 1]
        return torch.zeros(4, dtype=torch.float32, device=device)

    def end_step(self) -> None:
        """The pool's slots are spent: later backward calls (outside a training_step) allocate their own scratch again."""
        self._gscal_next = 1 << 60

    def _check(self, rc: int, allow_unsupported: bool = False) -> bool:
        if rc == _cabi.OK:
            return True
        if allow_unsupported and rc == _cabi.ERR_UNSUPPORTED:
            return False
        raise RuntimeError("libvqvae3d_b200: " + self.lib.vq3d_last_error().decode())

    # -- quantizer --------------------------------------------------------------------
    def vq_assign(self, x: Tensor, embed: Tensor, want_stats: bool):
        """x (B, D, *spatial) fp32 -> quant (same shape), idx (B, *spatial) int64,
        sqerr double[1], stats = flat [K | K*D] fp32 buffer (counts then dw) | None."""
        x = self._t(x.detach())
        embed = self._t(embed)
        B, D = x.shape[0], x.shape[1]
        S = x[0, 0].numel()
        K = embed.shape[0]
        quant = torch.empty_like(x)
        idx = torch.empty((B,) + tuple(x.shape[2:]), dtype=torch.int64, device=x.device)
        sqerr = torch.empty(1, dtype=torch.float64, device=x.device)        # zeroed by the C call
        counts = dw = stats = None
        if want_stats:
            stats = torch.zeros(K * (D + 1), dtype=torch.float32, device=x.device)   # one flat all-reduce buffer
            counts, dw = stats[:K], stats[K:]
        meta = dict(nbytes=B * S * (8 * D + 8), flops=2 * B * S * K * D, tag=f"N{B * S}xD{D}xK{K}")
        if self.vq_tensor_cores and D in (32, 64, 128) and B * S >= 32768 and K >= 64:
            need = self.lib.vq3d_vq_assign_tc_workspace(D, K)
            if need:       # tensor-core candidate pass + exact fp32 re-rank: same bits as the scan below
                ws = self._workspace(need, x.device)
                if self._call("vq_assign_tc", self.lib.vq3d_vq_assign_tc,
                              (self._p(x), self._p(embed), B, D, S, K, self._p(quant), self._p(idx), self._p(sqerr), self._p(counts),
                               self._p(dw), self._p(ws), ws.numel(), self.stream()), kernels=4, allow_unsupported=True, **meta):
                    return quant, idx, sqerr, stats
        self._call("vq_assign", self.lib.vq3d_vq_assign,
                   (self._p(x), self._p(embed), B, D, S, K, self._p(quant), self._p(idx), self._p(sqerr), self._p(counts),
                    self._p(dw), self.stream()), **meta)
        return quant, idx, sqerr, stats

    def vq_loss(self, sqerr: Tensor, commitment_cost: float, numel: int) -> Tensor:
        loss = torch.empty((), dtype=torch.float32, device=sqerr.device)
        self._call("vq_loss", self.lib.vq3d_vq_loss, (self._p(sqerr), float(commitment_cost), numel, self._p(loss), self.stream()))
        return loss

    def vq_ema_update(self, counts, dw, decay, alpha, cluster_size, embed_avg, embed) -> None:
        K, D = embed.shape
        for t in (cluster_size, embed_avg, embed):
            assert t.is_contiguous()
        self._call("vq_ema_update", self.lib.vq3d_vq_ema_update,
                   (self._p(self._t(counts)), self._p(self._t(dw)), K, D, float(decay), float(alpha),
                    self._p(self._t(cluster_size)), self._p(self._t(embed_avg)), self._p(self._t(embed)), self.stream()))

    def vq_init_stats(self, x: Tensor) -> Tensor:
        x = self._t(x.detach())
        B, D = x.shape[0], x.shape[1]
        S = x[0, 0].numel()
        scratch = torch.empty(2 * D, dtype=torch.float64, device=x.device)
        meanstd = torch.empty((2, D), dtype=torch.float32, device=x.device)
        self._call("vq_init_stats", self.lib.vq3d_vq_init_stats,
                   (self._p(x), B, D, S, self._p(scratch), self._p(meanstd), self.stream()), kernels=2)
        return meanstd

    def vq_init_apply(self, meanstd, total_vectors, embed, embed_avg, cluster_size, first_pass) -> None:
        K, D = embed.shape
        self._call("vq_init_apply", self.lib.vq3d_vq_init_apply,
                   (self._p(self._t(meanstd)), K, D, float(total_vectors), self._p(self._t(embed)), self._p(self._t(embed_avg)),
                    self._p(self._t(cluster_size)), self._p(self._t(first_pass, torch.int64)), self.stream()))

    def embed_code(self, idx: Tensor, embed: Tensor) -> Tensor:
        idx = self._t(idx, torch.int64)
        embed = self._t(embed)
        K, D = embed.shape
        out = torch.empty(tuple(idx.shape) + (D,), dtype=torch.float32, device=embed.device)
        self._call("vq_embed_code", self.lib.vq3d_vq_embed_code,
                   (self._p(idx), self._p(embed), idx.numel(), D, K, self._p(out), self.stream()), nbytes=idx.numel() * (8 + 4 * D))
        return out

    def vq_backward(self, grad_quant, grad_loss, x, quant, commitment_cost) -> Tensor:
        x, quant = self._t(x), self._t(quant)
        gx = torch.empty_like(x)
        self._call("vq_backward", self.lib.vq3d_vq_backward,
                   (self._p(self._t(grad_quant)), self._p(self._t(grad_loss)), self._p(x), self._p(quant), x.numel(),
                    float(commitment_cost), self._p(gx), self.stream()), nbytes=16 * x.numel())
        return gx

    # -- convolution ------------------------------------------------------------------
    def conv3d(self, x1: Tensor, w: Tensor, *, x2: Optional[Tensor] = None, bias: Optional[Tensor] = None, stride: int = 1,
               pad: int = 0, circular: bool = False, pre_act: bool = False, pre_a: Optional[Tensor] = None,
               pre_b: Optional[Tensor] = None, post_scale: Optional[Tensor] = None, post_b: Optional[Tensor] = None,
               residual: Optional[Tensor] = None, post_act: bool = False) -> Tensor:
        """The fused convolution of include/vqvae3d_b200.h (vq3d_conv3d / vq3d_conv3d_tc).  Differentiable: when
        autograd is recording and an input requires grad, the backward is vq3d_conv3d_backward."""
        tensors = (x1, x2, w, bias, pre_a, pre_b, post_scale, post_b, residual)
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
            # post_act = FixupResBlock's trailing ELU (layers.py:288-289): fused in the forward kernel, undone first in backward
            return _ConvFn.apply(self, dict(stride=stride, pad=pad, circular=circular, pre_act=pre_act, post_act=post_act), *tensors)
        return self._conv3d_fwd(x1, w, x2=x2, bias=bias, stride=stride, pad=pad, circular=circular, pre_act=pre_act, pre_a=pre_a,
                                pre_b=pre_b, post_scale=post_scale, post_b=post_b, residual=residual, post_act=post_act)

    def conv_desc(self, x1, x2, w, stride, pad, circular, pre_act, pre_a, pre_b, post_scale=None, post_b=None, bias=None,
                  residual=None, y=None, post_act=False):
        B, C1, H, W, Z = x1.shape
        return _cabi.ConvDesc(B=B, H=H, W=W, Z=Z, C1=C1, C2=0 if x2 is None else x2.shape[1], Cout=w.shape[0], k=w.shape[2],
                              stride=stride, pad=pad, pad_circular=int(circular), pre_act=int(pre_act), post_act=int(post_act),
                              x1=self._p(x1), x2=self._p(x2), w=self._p(w), bias=self._p(bias), pre_a=self._p(pre_a), pre_b=self._p(pre_b),
                              post_scale=self._p(post_scale), post_b=self._p(post_b), residual=self._p(residual), y=self._p(y))

    def conv3d_backward(self, cfg, x1, x2, w, pre_a, pre_b, post_scale, post_b, gy, need, acc_w=None, acc_bias=None):
        """need: dict of booleans (x1, x2, w, bias, pre_a, pre_b, post_scale, post_b).  Returns the gradients (or None).
        acc_w / acc_bias: existing gradient buffers to ADD the weight / bias gradient to (then None is returned for them)."""
        D = lambda t: None if t is None else self._t(t.detach())
        x1, x2, w, pre_a, pre_b, post_scale, post_b, gy = D(x1), D(x2), D(w), D(pre_a), D(pre_b), D(post_scale), D(post_b), D(gy)
        k = w.shape[2]
        tag = f"{w.shape[1]}->{w.shape[0]} k{k}s{cfg['stride']} @{x1.shape[2]}x{x1.shape[3]}x{x1.shape[4]}"
        if self.fused_pointwise_bwd and k == 1 and cfg["stride"] == 1 and cfg["pad"] == 0 and w.shape[0] * (w.shape[1] + 1) <= 3072:
            # pointwise convolution: every gradient from ONE fused launch (the kernel recomputes what d scale needs)
            dev = x1.device
            gx1 = torch.empty_like(x1) if need["x1"] else None
            gx2 = torch.empty_like(x2) if (x2 is not None and need["x2"]) else None
            gw = (acc_w if acc_w is not None else torch.zeros_like(w)) if need["w"] else None
            gbias = (acc_bias if acc_bias is not None else torch.zeros(w.shape[0], dtype=torch.float32, device=dev)) if need["bias"] else None
            want_scal = need["pre_a"] or need["pre_b"] or need["post_scale"] or need["post_b"]
            gscal = self._gscal(dev) if want_scal else None
            d = self.conv_desc(x1, x2, w, 1, 0, False, cfg["pre_act"], pre_a, pre_b, post_scale, post_b)
            g = _cabi.ConvBwd(gy=self._p(gy), raw=None, gx1=self._p(gx1), gx2=self._p(gx2), gw=self._p(gw), gbias=self._p(gbias),
                              gscalars=self._p(gscal), skip_input_grads=0)
            if self._call("conv1x1_backward", self.lib.vq3d_conv1x1_backward, (C.byref(d), C.byref(g), self.stream()), tag=tag,
                          allow_unsupported=True):
                return gx1, gx2, None if acc_w is not None else gw, None if acc_bias is not None else gbias, gscal
        raw = None
        if need["post_scale"]:        # d scale = sum(gy * conv output before the post transform): recompute it
            raw = self._conv3d_fwd(x1, w, x2=x2, stride=cfg["stride"], pad=cfg["pad"], circular=cfg["circular"], pre_act=cfg["pre_act"],
                                   pre_a=pre_a, pre_b=pre_b, force_fp32=True)
        dev = x1.device
        gx1 = torch.empty_like(x1) if need["x1"] else None
        gx2 = torch.empty_like(x2) if (x2 is not None and need["x2"]) else None
        gw = (acc_w if acc_w is not None else torch.zeros_like(w)) if need["w"] else None
        gbias = (acc_bias if acc_bias is not None else torch.zeros(w.shape[0], dtype=torch.float32, device=dev)) if need["bias"] else None
        want_scal = need["pre_a"] or need["pre_b"] or need["post_scale"] or need["post_b"]
        gscal = self._gscal(dev) if want_scal else None
        d = self.conv_desc(x1, x2, w, cfg["stride"], cfg["pad"], cfg["circular"], cfg["pre_act"], pre_a, pre_b, post_scale, post_b)
        k = w.shape[2]
        tag = f"{w.shape[1]}->{w.shape[0]} k{k}s{cfg['stride']} @{x1.shape[2]}x{x1.shape[3]}x{x1.shape[4]}"
        # stride-1 "same" convolutions: the input gradient IS a forward convolution of gy with the mirrored, transposed weight,
        # so it runs on the forward kernels (tcgen05 in bf16 mode) and only the pre-transform chain is finished separately
        fwd_dgrad = (self.dgrad_as_forward and cfg["stride"] == 1 and k % 2 == 1 and cfg["pad"] == (k - 1) // 2
                     and (need["x1"] or need["x2"] or need["pre_a"] or need["pre_b"]))
        if fwd_dgrad:
            wt = (w.flip(2, 3, 4) if k > 1 else w).transpose(0, 1).contiguous()
            gu = self._conv3d_fwd(gy, wt, pad=cfg["pad"], circular=cfg["circular"])
            self._call("conv3d_dgrad_finish", self.lib.vq3d_conv3d_dgrad_finish,
                       (C.byref(d), self._p(gu), self._p(gx1), self._p(gx2), self._p(gscal), self.stream()), tag=tag)
        if not fwd_dgrad or need["w"] or need["bias"] or need["post_scale"] or need["post_b"]:
            g = _cabi.ConvBwd(gy=self._p(gy), raw=self._p(raw), gx1=self._p(gx1), gx2=self._p(gx2), gw=self._p(gw), gbias=self._p(gbias),
                              gscalars=self._p(gscal), skip_input_grads=int(fwd_dgrad))
            self._call("conv3d_backward", self.lib.vq3d_conv3d_backward, (C.byref(d), C.byref(g), self.stream()), kernels=3, tag=tag)
        return gx1, gx2, None if acc_w is not None else gw, None if acc_bias is not None else gbias, gscal

    def _conv3d_fwd(self, x1: Tensor, w: Tensor, *, x2: Optional[Tensor] = None, bias: Optional[Tensor] = None, stride: int = 1,
                    pad: int = 0, circular: bool = False, pre_act: bool = False, pre_a: Optional[Tensor] = None,
                    pre_b: Optional[Tensor] = None, post_scale: Optional[Tensor] = None, post_b: Optional[Tensor] = None,
                    residual: Optional[Tensor] = None, post_act: bool = False, force_fp32: bool = False) -> Tensor:
        detach = lambda t: None if t is None else t.detach()
        x1, x2, w, bias, pre_a, pre_b, post_scale, post_b, residual = map(detach, (x1, x2, w, bias, pre_a, pre_b, post_scale, post_b, residual))
        x1, x2, w = self._t(x1), self._t(x2), self._t(w)
        B, C1, H, W, Z = x1.shape
        C2 = 0 if x2 is None else x2.shape[1]
        Cout, Cin, k = w.shape[0], w.shape[1], w.shape[2]
        if Cin != C1 + C2:
            raise RuntimeError(f"conv3d: weight expects {Cin} input channels, got {C1}+{C2}")
        if x2 is not None and tuple(x2.shape[2:]) != (H, W, Z):
            raise RuntimeError("conv3d: x1/x2 spatial mismatch")
        out_sp = tuple((n + 2 * pad - k) // stride + 1 for n in (H, W, Z))
        y = torch.empty((B, Cout) + out_sp, dtype=torch.float32, device=x1.device)
        residual = self._t(residual)
        if residual is not None and residual.shape != y.shape:
            raise RuntimeError("conv3d: residual shape mismatch")
        d = _cabi.ConvDesc(B=B, H=H, W=W, Z=Z, C1=C1, C2=C2, Cout=Cout, k=k, stride=stride, pad=pad,
                           pad_circular=int(circular), pre_act=int(pre_act), post_act=int(post_act),
                           x1=self._p(x1), x2=self._p(x2), w=self._p(w), bias=self._p(self._t(bias)),
                           pre_a=self._p(self._t(pre_a)), pre_b=self._p(self._t(pre_b)),
                           post_scale=self._p(self._t(post_scale)), post_b=self._p(self._t(post_b)),
                           residual=self._p(residual), y=self._p(y))
        so = y[0, 0].numel()
        meta = dict(nbytes=4 * B * (Cin * H * W * Z + Cout * so * (2 if residual is not None else 1)),
                    flops=2 * B * Cin * Cout * k ** 3 * so, tag=f"{Cin}->{Cout} k{k}s{stride} @{H}x{W}x{Z}")
        # k3 "same" convolutions with few input channels on big tensors: the shared-memory tiled fp32 kernel beats the gather-based
        # tensor-core kernel (measured 9->9 @64x64x32: 35 vs 131 us), and is exact
        tiled_simt = k == 3 and stride == 1 and pad == 1 and Cin <= 16 and B * so >= 16384
        if (self.precision == "bf16" and not force_fp32 and not tiled_simt and Cout >= 8 and Cin * k ** 3 >= 32
                and B * so >= self.tc_min_voxels):
            need = self.lib.vq3d_conv3d_tc_workspace(C.byref(d))       # > 0: few voxels, long K -> split-K through a workspace
            ws = self._workspace(need, x1.device) if need else None
            if self._call("conv3d_tc", self.lib.vq3d_conv3d_tc, (C.byref(d), self._p(ws), need, self.stream()),
                          allow_unsupported=True, **meta):
                return y
        self._call("conv3d", self.lib.vq3d_conv3d, (C.byref(d), self.stream()), **meta)
        return y

    def upsample2x(self, x: Tensor, *, pre_act: bool = False, pre_a: Optional[Tensor] = None,
                   pre_b: Optional[Tensor] = None) -> Tensor:
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, pre_a, pre_b)):
            return _UpsampleFn.apply(self, pre_act, x, pre_a, pre_b)
        return self._upsample2x_fwd(x, pre_act=pre_act, pre_a=pre_a, pre_b=pre_b)

    def upsample2x_backward(self, gy, x, pre_act, pre_a, pre_b, want_scal):
        gy, x = self._t(gy.detach()), self._t(x.detach())
        B, Cc, H, W, Z = x.shape
        gx = torch.empty_like(x)
        gscal = torch.zeros(4, dtype=torch.float32, device=x.device) if want_scal else None
        D = lambda t: None if t is None else self._t(t.detach())
        self._call("upsample2x_backward", self.lib.vq3d_upsample2x_backward,
                   (self._p(gy), self._p(x), B, Cc, H, W, Z, int(pre_act), self._p(D(pre_a)), self._p(D(pre_b)), self._p(gx), self._p(gscal),
                    self.stream()), tag=f"{Cc}ch @{H}x{W}x{Z}")
        return gx, gscal

    def _upsample2x_fwd(self, x: Tensor, *, pre_act: bool = False, pre_a: Optional[Tensor] = None,
                        pre_b: Optional[Tensor] = None) -> Tensor:
        detach = lambda t: None if t is None else t.detach()
        x, pre_a, pre_b = detach(x), detach(pre_a), detach(pre_b)
        x = self._t(x)
        B, Cc, H, W, Z = x.shape
        y = torch.empty((B, Cc, 2 * H, 2 * W, 2 * Z), dtype=torch.float32, device=x.device)
        self._call("upsample2x", self.lib.vq3d_upsample2x,
                   (self._p(x), B, Cc, H, W, Z, int(pre_act), self._p(self._t(pre_a)), self._p(self._t(pre_b)), self._p(y),
                    self.stream()), nbytes=4 * B * Cc * H * W * Z * 9, tag=f"{Cc}ch @{H}x{W}x{Z}")
        return y

    def preact_desc(self, x: Tensor, y: Optional[Tensor], blk, mode: int, tail=None, out_y: Optional[Tensor] = None,
                    pre=None) -> "_cabi.PreactDesc":
        B, Cin, H, W, Z = x.shape
        if pre is not None:          # fused leading 1x1 conv: x is its 1-channel input, the block sees pre's output channels
            Cin = pre.weight.shape[0]
        w1, w2, w3 = blk.branch_conv1.weight, blk.branch_conv2.weight, blk.branch_conv3.weight
        skip = blk.skip_conv.weight if blk.skip_conv is not None else None
        g = lambda n: self._p(self._t(getattr(blk, n).data)) if hasattr(blk, n) else None
        return _cabi.PreactDesc(B=B, H=H, W=W, Z=Z, Cin=Cin, Cb=w1.shape[0], Cout=w3.shape[0], mode=mode,
                                x=self._p(x), w1=self._p(self._t(w1.data)), w2=self._p(self._t(w2.data)),
                                w3=self._p(self._t(w3.data)), wskip=None if skip is None else self._p(self._t(skip.data)),
                                b1a=g("bias1a"), b1b=g("bias1b"), b2a=g("bias2a"), b2b=g("bias2b"), b3a=g("bias3a"),
                                b3b=g("bias3b"), b4=g("bias4"), scale=g("scale"), b1c=g("bias1c"), b1d=g("bias1d"),
                                y=None if y is None else self._p(y),
                                out_w=None if tail is None else self._p(self._t(tail.weight.data)),
                                out_b=None if tail is None or tail.bias is None else self._p(self._t(tail.bias.data)),
                                out_y=None if out_y is None else self._p(out_y),
                                pre_w=None if pre is None else self._p(self._t(pre.weight.data)),
                                pre_b=None if pre is None or pre.bias is None else self._p(self._t(pre.bias.data)))

    def preact_block(self, x: Tensor, blk, mode: int, pre=None) -> Optional[Tensor]:
        """Whole PreActFixupResBlock in one launch; None if no fused kernel covers the shape.  pre: a 1x1 Conv3d
        (1 -> Cin channels) applied to x on the fly (the encoder's parse_input)."""
        x = self._t(x)
        B, Cin, H, W, Z = x.shape
        if pre is not None:
            if Cin != 1 or pre.weight.shape[1] != 1:
                return None
            Cin = pre.weight.shape[0]
        Cout = blk.branch_conv3.weight.shape[0]
        sp = {0: (H, W, Z), 1: (H // 2, W // 2, Z // 2), 2: (2 * H, 2 * W, 2 * Z)}[mode]
        y = torch.empty((B, Cout) + sp, dtype=torch.float32, device=x.device)
        d = self.preact_desc(x, y, blk, mode, pre=pre)
        si, so = H * W * Z, sp[0] * sp[1] * sp[2]
        Cb = blk.branch_conv1.weight.shape[0]
        k3 = blk.branch_conv2.weight.shape[2] ** 3
        meta = dict(nbytes=4 * B * ((1 if pre is not None else Cin) * si + Cout * so),
                    flops=2 * B * (Cin * Cb * si + Cb * Cb * k3 * so + Cb * Cout * so + (Cin * Cout * so * (8 if mode == 1 else 1) if blk.skip_conv is not None else 0)),
                    tag=f"{'in+' if pre is not None else ''}{('same', 'down', 'up')[mode]} {Cin}->{Cb}->{Cout} @{H}x{W}x{Z}")
        # wide 'up' blocks: low-resolution pointwise stage + trilinear expansion (fp32 SIMT) feeding the tcgen05 stack kernel
        if mode == 2 and self.precision == "bf16" and self.up_tc and pre is None and blk.skip_conv is not None and (Cout, Cb) in self.TC_UP_SHAPES:
            need = self.lib.vq3d_preact_up_tc_workspace(C.byref(d))
            if need:
                ws = self._workspace(need, x.device)
                if self._call("preact_up_tc", self.lib.vq3d_preact_up_tc, (C.byref(d), self._p(ws), ws.numel(), self.stream()),
                              allow_unsupported=True, kernels=4, **meta):
                    return y
        ok = self._call("preact_block", self.lib.vq3d_preact_block, (C.byref(d), self.stream()), allow_unsupported=True, **meta)
        return y if ok else None

    def preact_same_backward_workspace(self, x: Tensor, blk) -> int:
        """> 0 when the fused backward covers this 'same' block (no skip convolution, C <= 32, Cb <= 16)."""
        if not self.fused_block_bwd or blk.skip_conv is not None or blk.branch_conv2.weight.shape[2] != 3:
            return 0
        if x.shape[0] * x[0, 0].numel() > self.fused_block_bwd_max_voxels:      # big thin tensors: the composed kernels win clearly
            return 0
        d = self.preact_desc(x, None, blk, 0)
        return int(self.lib.vq3d_preact_same_backward_workspace(C.byref(d)))

    def preact_same_backward(self, x: Tensor, blk, gy: Tensor, want_x: bool):
        """Every gradient of a 'same' block from its input x and gy (vq3d_preact_same_backward: 3 launches).
        -> gx | None, gw1, gw2, gw3, gscal[8] (d bias1a, bias1b, bias2a, bias2b, bias3a, bias3b, bias4, scale)."""
        x, gy = self._t(x.detach()), self._t(gy.detach())
        d = self.preact_desc(x, None, blk, 0)
        need = int(self.lib.vq3d_preact_same_backward_workspace(C.byref(d)))
        ws = self._workspace(need, x.device)
        gx = torch.empty_like(x) if want_x else None
        w1, w2, w3 = blk.branch_conv1.weight, blk.branch_conv2.weight, blk.branch_conv3.weight
        gw1, gw2, gw3 = torch.zeros_like(w1), torch.zeros_like(w2), torch.zeros_like(w3)
        gscal = torch.zeros(8, dtype=torch.float32, device=x.device)
        B, Cc, H, W, Z = x.shape
        self._call("preact_same_backward", self.lib.vq3d_preact_same_backward,
                   (C.byref(d), self._p(gy), self._p(ws), ws.numel(), self._p(gx), self._p(gw1), self._p(gw2), self._p(gw3), self._p(gscal),
                    self.stream()), kernels=3, tag=f"same {Cc}->{w1.shape[0]}->{Cc} @{H}x{W}x{Z}", nbytes=16 * x.numel())
        return gx, gw1, gw2, gw3, gscal

    # (Cout, Cb) pairs whose 'up' blocks run on the tensor-core kernel in bf16 mode (vq3d_preact_up_tc)
    TC_UP_SHAPES = {(8, 9), (16, 16), (8, 8), (32, 36)}

    # (C, Cb) pairs whose 'same' blocks run on the tensor-core stack kernel in bf16 mode
    TC_STACK_SHAPES = {(8, 4), (16, 8), (18, 9), (32, 16), (64, 32), (72, 36)}
    TC_STACK_MAX_BLOCKS = 24     # blocks per launch (kTcsMaxBlocks)

    def preact_stack(self, x: Tensor, blocks, tail=None) -> Optional[Tensor]:
        """n >= 1 consecutive equal-shape 'same' blocks; None if unsupported.  tail: a 1x1 Conv3d (C -> 1)
        to fuse into the last block's epilogue (the decoder's `out` conv); the result is then its output."""
        x = self._t(x)
        n = len(blocks)
        B, Cc, H, W, Z = x.shape
        # the fused tail is a C -> 1 pointwise convolution (out_w is read as one C-vector); anything else runs on its own
        if tail is not None and (tuple(tail.weight.shape[:2]) != (1, Cc) or tuple(tail.weight.shape[2:]) != (1, 1, 1)):
            return None
        y = torch.empty_like(x)
        tmp = torch.empty_like(x) if n > 1 else None
        out_y = torch.empty((B, 1, H, W, Z), dtype=torch.float32, device=x.device) if tail is not None else None
        arr = (_cabi.PreactDesc * n)()
        for i, blk in enumerate(blocks):
            last = i == n - 1
            arr[i] = self.preact_desc(x, y, blk, 0, tail if last else None, out_y if last else None)
        Cb = blocks[0].branch_conv1.weight.shape[0]
        meta = dict(nbytes=n * 8 * x.numel() - (3 * x.numel() if tail is not None else 0),
                    flops=n * 2 * B * H * W * Z * (2 * Cc * Cb + 27 * Cb * Cb),
                    tag=f"stack{n}{'+out' if tail is not None else ''} {Cc}->{Cb}->{Cc} @{H}x{W}x{Z}")
        res = out_y if tail is not None else y
        # Cb <= 4 pads the MMA K extent 4x: the tensor pipe's smem operand fetch then costs more than the SIMT
        # kernel on big tensors; keep the tensor-core path for them only where latency (small tensors) dominates
        tc_ok = tail is None and (Cc, Cb) in self.TC_STACK_SHAPES and (Cb > 4 or H * W * Z <= 65536)
        if self.precision == "bf16" and tc_ok:
            need = self.lib.vq3d_preact_stack_tc_workspace(C.byref(arr[0]))
            if need:
                ws = self._workspace(need, x.device)
                launches = -(-n // self.TC_STACK_MAX_BLOCKS)
                if self._call("preact_stack_tc", self.lib.vq3d_preact_stack_tc,
                              (arr, n, self._p(ws), ws.numel(), self.stream()), allow_unsupported=True, kernels=launches, **meta):
                    return y
        # thin blocks on big tensors: conv2 as one merged-tap GEMM per 128 voxels (the SIMT row kernel is FMA-issue bound there)
        if (self.precision == "bf16" and self.thin_tc and (Cc, Cb) in ((4, 2), (8, 4)) and H * W * Z > 65536 and Z in (32, 64, 128)
                and H % 8 == 0 and W % 8 == 0 and not (n == 1 and tail is not None)):
            if self._call("preact_stack_thin_tc", self.lib.vq3d_preact_stack_thin_tc, (arr, n, self._p(tmp), self.stream()),
                          allow_unsupported=True, kernels=n, **meta):
                return res
        if n == 1:
            d = arr[0]
            return res if self._call("preact_block", self.lib.vq3d_preact_block, (C.byref(d), self.stream()),
                                     allow_unsupported=True, **meta) else None
        ok = self._call("preact_stack", self.lib.vq3d_preact_stack, (arr, n, self._p(tmp), self.stream()), allow_unsupported=True,
                        kernels=n, **meta)
        return res if ok else None

    def elu_backward(self, gy: Tensor, y: Tensor) -> Tensor:
        gy, y = self._t(gy), self._t(y)
        gx = torch.empty_like(gy)
        self._call("elu_backward", self.lib.vq3d_elu_backward, (self._p(gy), self._p(y), self._p(gx), gy.numel(), self.stream()),
                   nbytes=12 * gy.numel())
        return gx

    def evonorm_s0(self, x: Tensor, v: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5) -> Tensor:
        """EvoNorm3D-S0 (evonorm.py:70-76), batch 1 like the reference; differentiable (x, v, gamma, beta)."""
        if torch.is_grad_enabled() and any(t.requires_grad for t in (x, v, gamma, beta)):
            return _EvoNormFn.apply(self, float(eps), x, v, gamma, beta)
        return self._evonorm_s0_fwd(x, v, gamma, beta, eps)[0]

    def _evonorm_s0_fwd(self, x: Tensor, v: Tensor, gamma: Tensor, beta: Tensor, eps: float = 1e-5):
        """-> (y, std [C], group sums scratch [2 * groups] doubles)."""
        x = self._t(x.detach())
        B, Cc = x.shape[0], x.shape[1]
        if B != 1:
            raise RuntimeError("EvoNorm3DS0 supports batch size 1 only (as the reference, evonorm.py:24)")
        S = x[0, 0].numel()
        groups = max(Cc // 8, 1)
        scratch = torch.empty(2 * groups, dtype=torch.float64, device=x.device)
        std = torch.empty(Cc, dtype=torch.float32, device=x.device)
        y = torch.empty_like(x)
        self._call("evonorm_s0_stats", self.lib.vq3d_evonorm_s0_stats,
                   (self._p(x), Cc, S, groups, float(eps), self._p(scratch), self._p(std), self.stream()), kernels=2, nbytes=4 * x.numel())
        flat = lambda t: self._p(self._t(t.detach().reshape(-1)))
        self._call("evonorm_s0_apply", self.lib.vq3d_evonorm_s0_apply,
                   (self._p(x), flat(v), flat(gamma), flat(beta), self._p(std), Cc, S, self._p(y), self.stream()), nbytes=8 * x.numel())
        return y, std, scratch

    def evonorm_s0_backward(self, gy: Tensor, x: Tensor, v: Tensor, gamma: Tensor, std: Tensor, scratch: Tensor):
        """Gradients of EvoNorm3D-S0 wrt (x, v, gamma, beta): two kernels (per-channel sums, elementwise dx) around the
        C-sized algebra (evonorm.py:12-26,36-47 differentiated by hand; batch 1)."""
        x, gy = self._t(x.detach()), self._t(gy)
        Cc = x.shape[1]
        S = x[0, 0].numel()
        groups = max(Cc // 8, 1)
        cpg = Cc // groups
        n = float(cpg * S)
        vf, gf = self._t(v.detach().reshape(-1)), self._t(gamma.detach().reshape(-1))
        sums = torch.empty(3 * Cc, dtype=torch.float64, device=x.device)
        self._call("evonorm_s0_backward_sums", self.lib.vq3d_evonorm_s0_backward_sums,
                   (self._p(x), self._p(gy), self._p(vf), Cc, S, self._p(sums), self.stream()), kernels=2, nbytes=8 * x.numel())
        sums = sums.view(Cc, 3)
        inv = 1.0 / std.double()                                              # per channel (equal within a group)
        g_beta = sums[:, 0]
        g_gamma = sums[:, 1] * inv
        g_v = sums[:, 2] * gf.double() * inv
        a_grp = (sums[:, 1] * gf.double()).view(groups, cpg).sum(1).repeat_interleave(cpg)   # d loss / d (1 / std_g)
        coef_a = (gf.double() * inv).float().contiguous()
        coef_b = (-a_grp * inv ** 3 / (n - 1.0)).float().contiguous()
        mean = (scratch.view(groups, 2)[:, 0] / n).repeat_interleave(cpg).float().contiguous()
        gx = torch.empty_like(x)
        self._call("evonorm_s0_backward_apply", self.lib.vq3d_evonorm_s0_backward_apply,
                   (self._p(x), self._p(gy), self._p(vf), self._p(coef_a), self._p(coef_b), self._p(mean), Cc, S, self._p(gx),
                    self.stream()), nbytes=12 * x.numel())
        return gx, g_v.float().view_as(v), g_gamma.float().view_as(gamma), g_beta.float().view_as(gamma)

    def elu_hu_rint(self, decoded: Tensor, scale: float = 1000.0, offset: float = 1000.0, dtype=torch.int64) -> Tensor:
        """rint(ELU(decoded) * scale - offset): the Hounsfield-unit output of decode_embeddings.py:43-47, as int64 (the
        reference's `astype(int)`) or saturating int16 (2 bytes per voxel over PCIe instead of 8)."""
        decoded = self._t(decoded.detach())
        out = torch.empty(decoded.shape, dtype=dtype, device=decoded.device)
        if dtype == torch.int16:
            self._call("elu_hu_rint_i16", self.lib.vq3d_elu_hu_rint_i16,
                       (self._p(decoded), decoded.numel(), float(scale), float(offset), self._p(out), self.stream()), nbytes=6 * decoded.numel())
        elif dtype == torch.int64:
            self._call("elu_hu_rint", self.lib.vq3d_elu_hu_rint,
                       (self._p(decoded), decoded.numel(), float(scale), float(offset), self._p(out), self.stream()), nbytes=12 * decoded.numel())
        else:
            raise RuntimeError("elu_hu_rint: dtype must be torch.int64 or torch.int16")
        return out

    def hu_to_network(self, hu: Tensor, min_hu: float = -1500.0, max_hu: float = 3000.0, mul: float = 0.001, add: float = 1.0) -> Tensor:
        """Raw int16 Hounsfield units -> the network's fp32 input range on the device: clip, * 0.001f, + 1
        (utils/load_nrrd_dataset.py:73-81)."""
        hu = self._t(hu, torch.int16)
        out = torch.empty(hu.shape, dtype=torch.float32, device=hu.device)
        self._call("hu_to_network", self.lib.vq3d_hu_to_network,
                   (self._p(hu), hu.numel(), float(min_hu), float(max_hu), float(mul), float(add), self._p(out), self.stream()), nbytes=6 * hu.numel())
        return out

    def huber_loss(self, decoded: Tensor, x: Tensor, num_valid: Optional[Tensor], mask_hw: Optional[Tensor]) -> Tensor:
        """mean smooth_l1(mask(ELU(decoded)), x) as a 0-d fp32 tensor (model.py:120-152); differentiable wrt decoded."""
        if torch.is_grad_enabled() and decoded.requires_grad:
            return _HuberFn.apply(self, decoded, x, num_valid, mask_hw)
        acc = self.huber_elu_mask(decoded, x, num_valid, mask_hw)
        return (acc[0] / acc[1]).float()

    def huber_backward(self, decoded, x, num_valid, mask_hw, count, gloss) -> Tensor:
        decoded, x = self._t(decoded.detach()), self._t(x.detach())
        B, H, W, Z, num_valid = self._huber_extent(decoded, x, num_valid)
        gdec = torch.empty_like(decoded)
        self._call("huber_elu_mask_backward", self.lib.vq3d_huber_elu_mask_backward,
                   (self._p(decoded), self._p(x), self._p(self._t(num_valid, torch.int32)), self._p(self._t(mask_hw, torch.uint8)), B, H, W, Z,
                    self._p(count), self._p(self._t(gloss.detach().reshape(1).float())), self._p(gdec), self.stream()), nbytes=12 * x.numel())
        return gdec

    def adam_amsgrad_step(self, p: Tensor, g: Tensor, m: Tensor, v: Tensor, vmax: Tensor, lr, b1, b2, eps, step) -> None:
        for t in (p, g, m, v, vmax):
            assert t.is_contiguous() and t.dtype == torch.float32
        self._call("adam_amsgrad_step", self.lib.vq3d_adam_amsgrad_step,
                   (self._p(p), self._p(g), self._p(m), self._p(v), self._p(vmax), p.numel(), float(lr), float(b1), float(b2), float(eps),
                    int(step), self.stream()), nbytes=28 * p.numel())

    def adam_amsgrad_step_dev(self, p: Tensor, g: Tensor, m: Tensor, v: Tensor, vmax: Tensor, lr, b1, b2, eps, step_state: Tensor) -> None:
        """Same update with the step counter / bias corrections in `step_state` (3 doubles on the device): graph-capturable."""
        for t in (p, g, m, v, vmax):
            assert t.is_contiguous() and t.dtype == torch.float32
        assert step_state.dtype == torch.float64 and step_state.numel() == 3
        self._call("adam_amsgrad_step", self.lib.vq3d_adam_amsgrad_step_dev,
                   (self._p(p), self._p(g), self._p(m), self._p(v), self._p(vmax), p.numel(), float(lr), float(b1), float(b2), float(eps),
                    self._p(step_state), self.stream()), kernels=2, nbytes=28 * p.numel())

    @staticmethod
    def _huber_extent(decoded: Tensor, x: Tensor, num_valid: Optional[Tensor]):
        """smooth_l1 runs over every channel (model.py:163): the kernels index (b, hw, z), so channels fold into the batch
        extent and each volume's num_valid_slices entry is repeated per channel."""
        if decoded.shape != x.shape:
            raise RuntimeError(f"huber: decoded {tuple(decoded.shape)} vs target {tuple(x.shape)}")
        B, Cc, H, W, Z = x.shape
        if num_valid is not None:
            if num_valid.numel() != B:
                raise RuntimeError(f"huber: num_valid_slices has {num_valid.numel()} entries for a batch of {B}")
            if Cc > 1:
                num_valid = num_valid.repeat_interleave(Cc)
        return B * Cc, H, W, Z, num_valid

    def huber_elu_mask(self, decoded: Tensor, x: Tensor, num_valid: Optional[Tensor], mask_hw: Optional[Tensor]):
        decoded, x = self._t(decoded.detach()), self._t(x.detach())
        B, H, W, Z, num_valid = self._huber_extent(decoded, x, num_valid)
        acc = torch.zeros(2, dtype=torch.float64, device=x.device)
        self._call("huber_elu_mask", self.lib.vq3d_huber_elu_mask,
                   (self._p(decoded), self._p(x), self._p(self._t(num_valid, torch.int32)), self._p(self._t(mask_hw, torch.uint8)),
                    B, H, W, Z, acc[0:1].data_ptr(), acc[1:2].data_ptr(), self.stream()), nbytes=8 * x.numel())
        return acc


    def huber_metrics(self, decoded: Tensor, x: Tensor, num_valid: Optional[Tensor], mask_hw: Optional[Tensor], data_range: float = 4.0,
                      medians: bool = True) -> dict:
        """The reference's validation log (model.py:143-149) -- recon_loss_{min,max,mean,median,std}, loc_{min,max,mean,median,std}
        (utils/logging_helpers.py:4-15), nmse, psnr (metrics/evaluate.py:18-24, data_range 4 as model.py:25) -- as 0-d tensors:
        one fused pass over (decoded, x) for the sums, four more for the exact medians (radix select; `medians=False` skips them)."""
        decoded, x = self._t(decoded.detach()), self._t(x.detach())
        B, H, W, Z, num_valid = self._huber_extent(decoded, x, num_valid)
        sums = torch.zeros(8, dtype=torch.float64, device=x.device)
        inf = float("inf")
        minmax = torch.tensor([inf, -inf, inf, -inf], dtype=torch.float32, device=x.device)
        self._call("huber_elu_mask_stats", self.lib.vq3d_huber_elu_mask_stats,
                   (self._p(decoded), self._p(x), self._p(self._t(num_valid, torch.int32)), self._p(self._t(mask_hw, torch.uint8)),
                    B, H, W, Z, self._p(sums), self._p(minmax), self.stream()), nbytes=8 * x.numel())
        s_l, n, s_d2, s_x2, s_loc, s_loc2, s_l2 = (sums[i] for i in range(7))
        std = lambda s1, s2: torch.sqrt(torch.clamp((s2 - s1 * s1 / n) / (n - 1), min=0.0)).float()          # torch.std: unbiased
        log = {"recon_loss_min": minmax[2], "recon_loss_max": minmax[3], "recon_loss_mean": (s_l / n).float(), "recon_loss_std": std(s_l, s_l2),
               "loc_min": minmax[0], "loc_max": minmax[1], "loc_mean": (s_loc / n).float(), "loc_std": std(s_loc, s_loc2),
               "nmse": (s_d2 / s_x2).float(), "psnr": (10.0 * torch.log10(data_range ** 2 / (s_d2 / n))).float()}
        if medians:
            med = torch.empty(2, dtype=torch.float32, device=x.device)
            nws = int(self.lib.vq3d_huber_elu_mask_medians_workspace())
            ws = self._workspace(nws, x.device)
            self._call("huber_elu_mask_medians", self.lib.vq3d_huber_elu_mask_medians,
                       (self._p(decoded), self._p(x), self._p(self._t(num_valid, torch.int32)), self._p(self._t(mask_hw, torch.uint8)),
                        B, H, W, Z, self._p(med), self._p(ws), ws.numel(), self.stream()), nbytes=5 * 8 * x.numel())
            log["loc_median"], log["recon_loss_median"] = med[0], med[1]
        return log


class _ConvFn(torch.autograd.Function):
    """Autograd edge of Ops.conv3d; residual's gradient is gy itself."""

    @staticmethod
    def forward(ctx, ops, cfg, x1, x2, w, bias, pre_a, pre_b, post_scale, post_b, residual):
        y = ops._conv3d_fwd(x1, w, x2=x2, bias=bias, pre_a=pre_a, pre_b=pre_b, post_scale=post_scale, post_b=post_b, residual=residual, **cfg)
        ctx.ops, ctx.cfg = ops, cfg
        ctx.save_for_backward(x1, x2, w, pre_a, pre_b, post_scale, post_b, y if cfg.get("post_act") else None)
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.bias_ref = bias
        return y

    @staticmethod
    def _acc_target(ops, t, wanted):
        """The existing gradient buffer of parameter `t` when the kernels may add to it directly (Ops.grad_inplace), else None."""
        if not (wanted and ops.grad_inplace and t is not None and t.is_leaf and t.requires_grad):
            return None
        g = t.grad
        if g is None or g.dtype != torch.float32 or g.device != t.device or not g.is_contiguous() or g.shape != t.shape or g.requires_grad:
            return None
        return g

    @staticmethod
    def backward(ctx, gy):
        x1, x2, w, pre_a, pre_b, post_scale, post_b, y_act = ctx.saved_tensors
        n = ctx.needs_input_grad      # (ops, cfg, x1, x2, w, bias, pre_a, pre_b, post_scale, post_b, residual)
        if y_act is not None:         # trailing ELU: everything below sees the gradient wrt its input
            gy = ctx.ops.elu_backward(gy.contiguous(), y_act)
        need = dict(x1=n[2], x2=n[3] and x2 is not None, w=n[4], bias=n[5] and ctx.has_bias, pre_a=n[6] and pre_a is not None,
                    pre_b=n[7] and pre_b is not None, post_scale=n[8] and post_scale is not None, post_b=n[9] and post_b is not None)
        gy = gy.contiguous()
        acc_w = _ConvFn._acc_target(ctx.ops, w, need["w"])
        acc_bias = _ConvFn._acc_target(ctx.ops, ctx.bias_ref, need["bias"])
        gx1, gx2, gw, gbias, gs = ctx.ops.conv3d_backward(ctx.cfg, x1, x2, w, pre_a, pre_b, post_scale, post_b, gy, need,
                                                          acc_w=acc_w, acc_bias=acc_bias)
        # the four Fixup scalars: ONE multi-tensor addition into their .grad instead of a clone + an AccumulateGrad kernel each
        outs, tgt, src = [], [], []
        for t, flag, i in ((pre_a, need["pre_a"], 0), (pre_b, need["pre_b"], 1), (post_scale, need["post_scale"], 2), (post_b, need["post_b"], 3)):
            a = _ConvFn._acc_target(ctx.ops, t, flag) if (flag and t.numel() == 1) else None
            if a is not None:
                tgt.append(a.view(1)); src.append(gs[i:i + 1]); outs.append(None)
            else:
                outs.append(gs[i:i + 1].clone() if flag else None)
        if tgt:
            torch._foreach_add_(tgt, src)
        return (None, None, gx1, gx2, gw, gbias, outs[0], outs[1], outs[2], outs[3], gy if (ctx.has_res and n[10]) else None)


class _PreactSameFn(torch.autograd.Function):
    """Autograd edge of a whole 'same' PreActFixupResBlock (layers.py:176-195): one fused forward launch that saves only the
    block input, and the fused 3-launch backward (vq3d_preact_same_backward).  Parameter order = _PARAMS below."""

    _PARAMS = ("branch_conv1.weight", "branch_conv2.weight", "branch_conv3.weight", "bias1a", "bias1b", "bias2a", "bias2b",
               "bias3a", "bias3b", "bias4", "scale")

    @staticmethod
    def forward(ctx, ops, blk, x, *params):
        y = ops.preact_block(x.detach(), blk, 0)
        if y is None:                                   # no fused forward kernel for the shape: the composed kernels (still no graph)
            y = blk.forward_composed(x.detach())
        ctx.ops, ctx.blk = ops, blk
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gx, gw1, gw2, gw3, gs = ctx.ops.preact_same_backward(x, ctx.blk, gy.contiguous(), ctx.needs_input_grad[2])
        scal = [gs[i:i + 1].clone() for i in range(8)]
        grads = [gw1, gw2, gw3] + scal
        need = ctx.needs_input_grad[3:]
        return (None, None, gx) + tuple(g if n else None for g, n in zip(grads, need))


class _EvoNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ops, eps, x, v, gamma, beta):
        y, std, scratch = ops._evonorm_s0_fwd(x, v, gamma, beta, eps)
        ctx.ops = ops
        ctx.save_for_backward(x, v, gamma, std, scratch)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, v, gamma, std, scratch = ctx.saved_tensors
        gx, gv, gg, gb = ctx.ops.evonorm_s0_backward(gy.contiguous(), x, v, gamma, std, scratch)
        n = ctx.needs_input_grad
        return None, None, gx if n[2] else None, gv if n[3] else None, gg if n[4] else None, gb if n[5] else None


class _UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ops, pre_act, x, pre_a, pre_b):
        ctx.ops, ctx.pre_act = ops, pre_act
        ctx.save_for_backward(x, pre_a, pre_b)
        return ops._upsample2x_fwd(x, pre_act=pre_act, pre_a=pre_a, pre_b=pre_b)

    @staticmethod
    def backward(ctx, gy):
        x, pre_a, pre_b = ctx.saved_tensors
        n = ctx.needs_input_grad
        want = (n[3] and pre_a is not None) or (n[4] and pre_b is not None)
        gx, gs = ctx.ops.upsample2x_backward(gy.contiguous(), x, ctx.pre_act, pre_a, pre_b, want)
        return (None, None, gx if n[2] else None, gs[0:1].clone() if (n[3] and pre_a is not None) else None,
                gs[1:2].clone() if (n[4] and pre_b is not None) else None)


class _HuberFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ops, decoded, x, num_valid, mask_hw):
        acc = ops.huber_elu_mask(decoded, x, num_valid, mask_hw)
        ctx.ops, ctx.nv, ctx.mask = ops, num_valid, mask_hw
        ctx.save_for_backward(decoded, x, acc)
        return (acc[0] / acc[1]).float()

    @staticmethod
    def backward(ctx, gloss):
        decoded, x, acc = ctx.saved_tensors
        return None, ctx.ops.huber_backward(decoded, x, ctx.nv, ctx.mask, acc[1:2], gloss), None, None, None


_DEFAULT: Optional[Ops] = None


def default() -> Ops:
    global _DEFAULT
    if _DEFAULT is None:
        _DEFAULT = Ops()
    return _DEFAULT
