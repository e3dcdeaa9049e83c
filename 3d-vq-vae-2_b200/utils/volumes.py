"""Volume sources for train.py / extract_embeddings.py.

The reference reads NRRD CT scans through utils/load_nrrd_dataset.py (clip to [-1500, 3000] HU, / 1000, + 1, pad the depth,
optional rescale; :73-86) and hands batches `(volume (B, 1, H, W, D) fp32, num_valid_slices)`.  That reader needs
pynrrd / monai, which this image does not have; the scripts therefore accept
  * a directory of `.npy` volumes that already went through that preprocessing, or raw HU volumes with `hu=True`
    (the same clip / scale / shift is applied here), and
  * `synthetic:N:HxWxD` -- N seeded volumes in the reference's value range,
both yielding exactly the reference's batch structure."""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np
import torch

MIN_HU, MAX_HU, SCALE = -1500.0, 3000.0, 1000.0          # utils/load_nrrd_dataset.py:73-81


def preprocess_hu(vol: np.ndarray) -> np.ndarray:
    """HU -> network range: clip, * (1 + (-1 + 1/1000)) in fp32, + 1 (air -> 0); utils/load_nrrd_dataset.py:73-81 (monai
    ScaleIntensity multiplies by 1 + factor; same arithmetic as the device kernel vq3d_hu_to_network)."""
    return (np.clip(vol.astype(np.float32), MIN_HU, MAX_HU) * np.float32(1 + (-1 + 1 / SCALE)) + np.float32(1.0)).astype(np.float32)


class SyntheticVolumeDataset(torch.utils.data.Dataset):
    def __init__(self, n: int, shape: Tuple[int, int, int], seed: int = 42):
        self.n, self.shape, self.seed = n, tuple(shape), seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed + i)
        return torch.rand(1, *self.shape, generator=g) * 4.5 - 0.5, self.shape[2]


class NpyVolumeDataset(torch.utils.data.Dataset):
    def __init__(self, path: str, hu: bool = False):
        self.files: List[str] = sorted(os.path.join(path, f) for f in os.listdir(path) if f.endswith(".npy"))
        if not self.files:
            raise FileNotFoundError(f"no .npy volumes under {path}")
        self.hu = hu

    def __len__(self):
        return len(self.files)

    def __getitem__(self, i):
        v = np.load(self.files[i])
        if v.ndim != 3:
            raise ValueError(f"{self.files[i]}: expected an (H, W, D) volume, got shape {v.shape}")
        v = preprocess_hu(v) if self.hu else v.astype(np.float32)
        return torch.from_numpy(v)[None], v.shape[2]


def open_dataset(spec: str, hu: bool = False) -> torch.utils.data.Dataset:
    spec = str(spec)
    if spec.startswith("synthetic:"):
        _, n, shp = spec.split(":")
        return SyntheticVolumeDataset(int(n), tuple(int(a) for a in shp.lower().split("x")))
    return NpyVolumeDataset(spec, hu=hu)
