"""Synthetic stand-in for the reference's Lightning datamodules (``mld/data/EgoBody.py:111-163``,
``mld/data/Gimo.py``): carries the float64 ``mean``/``std`` statistics, ``numdims`` and ``renorm`` the
model reads, and yields batches in the tuple layout ``ego_eval`` unpacks.  Real datasets are not
available offline."""
from __future__ import annotations

import numpy as np
import torch
from typing import Optional

from . import synthetic


class SyntheticDataModule:
    is_mm = False

    def __init__(self, cfg=None, name: str = "egobody", batch_size: int = 8, n_batches: int = 1, n_points: int = 20000,
                 T: int = 60, seed: int = 1234, ragged: bool = False, with_images: Optional[bool] = None):
        self.name = name
        self.njoints = 24 if name == "egobody" else 21
        self.numdims = 75 if name == "egobody" else 69
        mean, std = synthetic.norm_stats()
        self.mean, self.std = mean.numpy(), std.numpy()           # float64 [1,78] like the npy stats
        self.batch_size, self.n_batches, self.n_points, self.T, self.seed, self.ragged = batch_size, n_batches, n_points, T, seed, ragged
        if with_images is None:       # the image-conditioned configs' datasets yield (..., scene, images, length)
            try:
                with_images = "image" in list(cfg.model.condition)
            except Exception:          # noqa: BLE001
                with_images = False
        self.with_images = bool(with_images)

    def renorm(self, features):
        """EgoBody.py:151-157 -- promotes to float64 because the statistics are float64."""
        return features * torch.tensor(self.std[0, : self.numdims]).to(features.device) + torch.tensor(
            self.mean[0, : self.numdims]).to(features.device)

    def batch(self, i: int = 0, device=None):
        b = synthetic.make_batch(self.batch_size, seed=self.seed + i, n_points=self.n_points, T=self.T, ragged=self.ragged,
                                 dataset=self.name, with_images=self.with_images)
        if device is not None:
            b = tuple(x.to(device) if torch.is_tensor(x) else x for x in b)
        return b

    def test_dataloader(self, device=None):
        for i in range(self.n_batches):
            yield self.batch(i, device)
