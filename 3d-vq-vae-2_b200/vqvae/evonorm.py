"""Drop-in for the reference's vqvae/evonorm.py (EvoNorm3D-S0, evonorm.py:59-76).

Parameter names/shapes (v, gamma, beta: (C, 1, 1, 1); v=1, gamma=0, beta=0) match the
reference so checkpoints load.  The sm_100a kernels for it (group variance + SiLU-velocity,
SURVEY.md 8a row N1) are scheduled after the pre-activation path; forward raises until then."""
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
        raise NotImplementedError("EvoNorm3DS0: sm_100a kernels not built yet")
