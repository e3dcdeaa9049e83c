"""TEST INFRASTRUCTURE ONLY: drives the host-emulator build of the kernel sources
(tests/emu/build_emu.py) through the product's own op wrappers and modules, on CPU tensors,
so kernel/launch logic is checked against the oracle before any GPU time is spent."""
import ctypes
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_emu  # noqa: E402
from vqvae import _cabi, _ops  # noqa: E402


class EmuOps(_ops.Ops):
    def __init__(self):
        lib = _cabi.declare(ctypes.CDLL(build_emu.build()))
        assert lib.vq3d_is_cuda_build() == 0
        super().__init__(lib)

    def stream(self):
        return 0

    def _t(self, t, dtype=torch.float32):
        if t is None:
            return None
        assert not t.is_cuda and t.dtype == dtype, (t.device, t.dtype)
        return t if t.is_contiguous() else t.contiguous()


class use_emulator:
    """Context manager: route the product modules to the emulator for the duration of a test."""

    def __enter__(self):
        self.prev = _ops._DEFAULT
        _ops._DEFAULT = EmuOps()
        return _ops._DEFAULT

    def __exit__(self, *a):
        _ops._DEFAULT = self.prev
