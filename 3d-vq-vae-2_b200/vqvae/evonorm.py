"""Drop-in for the reference's vqvae/evonorm.py (EvoNorm3D-S0, evonorm.py:59-76).

Parameter names/shapes (v, gamma, beta: (C, 1, 1, 1); v=1, gamma=0, beta=0) match the
reference so checkpoints load.  Forward = two kernels of libvqvae3d_b200 (group statistics in double,
then the elementwise SiLU-velocity / normalise / affine) and its hand-derived backward, csrc/evonorm_kernels.cu."""
import torch
from torch import nn


def determine_num_groups(in_channels, preferred_channels_per_group=8):
    return max(in_channels // preferred_channels_per_group, 1)     # evonorm.py:8-9


class EvoNorm3DS0(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.v = nn.Parameter(torch.ones((in_channels, 1, 1, 1)))
        self.gamma = nn.Parameter(torch.zeros((in_channels, 1, 1, 1)))
        self.beta = nn.Parameter(torch.zeros((in_channels, 1, 1, 1)))

    def forward(self, x):
        assert x.dim() == 5
        from . import _ops
        return _ops.default().evonorm_s0(x, self.v, self.gamma, self.beta)
