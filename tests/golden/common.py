"""Shared by tests/golden/make_golden.py (runs where /root/reference exists) and the
tests (run anywhere): deterministic, portable weights so fixtures only need to hold the
reference's OUTPUTS.  numpy's legacy RandomState stream is frozen by numpy policy."""
from __future__ import annotations

import zlib
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

GOLDEN_DIR = __import__("os").path.dirname(__import__("os").path.abspath(__file__))


def golden_state_dict(spec: Sequence[Tuple[str, Sequence[int], str]], seed: int) -> Dict[str, torch.Tensor]:
    """spec: (key, shape, dtype-name) in the reference's state_dict order.  Values:
    conv weights ~ N(0, 1/fan_in), scalar Fixup biases ~ N(0, .1), scale ~ 1 + N(0, .1),
    conv biases ~ N(0, .1), codebooks ~ N(0, 1), cluster_size ~ U(1, 2), first_pass = 0,
    EvoNorm v ~ 1 + N(0,.1), gamma ~ N(1,.1), beta ~ N(0,.1)."""
    out = {}
    for key, shape, dtype in spec:
        rs = np.random.RandomState((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))
        shape = tuple(shape)
        leaf = key.split(".")[-1]
        parent = key.split(".")[-2] if "." in key else ""
        if leaf == "first_pass":
            val = np.zeros(shape, np.int64)
        elif leaf == "cluster_size":
            val = rs.uniform(1.0, 2.0, size=shape)
        elif leaf in ("embed", "embed_avg"):
            val = rs.standard_normal(shape)
        elif leaf == "weight":
            fan_in = int(np.prod(shape[1:]))
            val = rs.standard_normal(shape) / np.sqrt(fan_in)
            if parent == "branch_conv3" or (parent == "branch_conv2" and len(shape) == 5 and "evonorm" not in key):
                val = val * 0.7
        elif leaf == "scale" or leaf == "v":
            val = 1.0 + 0.1 * rs.standard_normal(shape)
        elif leaf == "gamma":
            val = 1.0 + 0.1 * rs.standard_normal(shape)
        else:  # every bias flavour, beta
            val = 0.1 * rs.standard_normal(shape)
        t = torch.from_numpy(np.asarray(val)).to(getattr(torch, dtype))
        out[key] = t
    return out


def spec_of(sd: Dict[str, torch.Tensor]):
    return [(k, list(v.shape), str(v.dtype).replace("torch.", "")) for k, v in sd.items()]


def portable_randn(shape, seed):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(tuple(shape)).astype(np.float32))


def portable_volume(shape, seed):
    """Synthetic CT-range volume: U(-0.5, 4.0), numpy stream."""
    return torch.from_numpy((np.random.RandomState(seed).random_sample(tuple(shape)) * 4.5 - 0.5).astype(np.float32))
