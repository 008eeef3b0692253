"""The training input pipeline of the reference on device.

Reference: `FugroDataset.sample_points` (data.py:107-143) draws `num_point` points of a tile without replacement with
class-balancing weights in Python loops over every point, `__getitem__` (data.py:146-164) indexes the tile,
`collate_pcs` (data.py:169-182) re-draws and copies into a float64 batch, and the training loop converts it with
`permute(0, 2, 1).float()` (train.py:176).  With a millisecond-scale step that host path is the bottleneck, so here the
tiles stay resident in HBM and a batch is produced by two kernels per block: `dgcnn_sample_points` (weighted sampling
without replacement as an exponential race + radix select) and `dgcnn_gather_points` (collate + layout + cast).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops


class DeviceTiles:
    """Tiles ("rooms", data.py:55-75) resident on one device: points [N_t, F] (float64 as loaded, or float32) and labels."""

    def __init__(self, tiles: Sequence[np.ndarray], num_classes: int, device: Optional[torch.device] = None,
                 dtype: torch.dtype = torch.float64):
        """tiles: arrays [N_t, F + 1] whose last column is the label (the .npy layout of data.py:62-64)."""
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.num_classes = num_classes
        self.points: List[torch.Tensor] = []
        self.labels: List[torch.Tensor] = []
        for t in tiles:
            t = np.asarray(t)
            self.points.append(torch.as_tensor(t[:, :-1]).to(self.device, dtype).contiguous())
            self.labels.append(torch.as_tensor(t[:, -1]).to(self.device).to(torch.int32).contiguous())
        self.num_features = self.points[0].shape[1] if self.points else 0

    def __len__(self) -> int:
        return len(self.points)

    def sample_indices(self, tile: int, num_point: int, seed: int) -> torch.Tensor:
        """data.py:107-143: int32 [num_point] point indices of `tile`, class-weighted, without replacement
        (index order; the point whose key is the n-th smallest comes last)."""
        lab = self.labels[tile]
        nt = lab.numel()
        lib = _lib.load()
        out = torch.empty(num_point, dtype=torch.int32, device=self.device)
        nbytes = lib.dgcnn_sample_points_workspace_bytes(nt)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        ops._call("dgcnn_sample_points", self.device, ops._ptr(lab), nt, self.num_classes, num_point, int(seed) & (2 ** 64 - 1),
                  ops._ptr(out), ops._ptr(ws), nbytes)
        return out

    def batch(self, tile_ids: Sequence[int], num_point: int, seed: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One training batch: x float32 [B, F, num_point] (the layout `model(x)` takes, train.py:176-179), labels int64
        [B, num_point] and the tile ids (data.py:164 returns them for the mask construction)."""
        B, F = len(tile_ids), self.num_features
        x = torch.empty(B, F, num_point, dtype=torch.float32, device=self.device)
        y = torch.empty(B, num_point, dtype=torch.int64, device=self.device)
        for b, t in enumerate(tile_ids):
            idx = self.sample_indices(t, num_point, seed * 1000003 + b)
            pts = self.points[t]
            ops._call("dgcnn_gather_points", self.device, ops._ptr(pts), 1 if pts.dtype == torch.float64 else 0, pts.stride(0),
                      ops._ptr(self.labels[t]), ops._ptr(idx), num_point, F, ops._ptr(x[b]), x.stride(1), x.stride(2), ops._ptr(y[b]))
        return x, y, torch.as_tensor(list(tile_ids), dtype=torch.long)
