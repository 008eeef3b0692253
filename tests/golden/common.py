"""Deterministic inputs/weights shared by make_golden.py (reference side) and the tests.

Everything is derived from numpy's legacy MT19937 `RandomState`, whose stream is frozen across
numpy versions, so the fixtures only need to store the reference's OUTPUTS.
"""
from __future__ import annotations

import numpy as np
import torch

BN_CHANNELS = (64, 64, 64, 64, 64, None, 512, 256)   # bn6 = emb_dims


def points(seed: int, B: int, C: int, N: int, kind: str = "unit") -> torch.Tensor:
    """[B,C,N] fp32.  kinds: unit U[0,1); real (xy in [0,30), z in [100,130), agl in [0,20), C==4);
    dup (unit with ~30% of points duplicated, as produced by with-replacement sampling,
    prepare_data/pointcloud_util.py:73-78); randn."""
    rs = np.random.RandomState(seed)
    if kind == "unit":
        a = rs.rand(B, C, N)
    elif kind == "randn":
        a = rs.randn(B, C, N)
    elif kind == "real":
        assert C == 4
        a = np.concatenate([rs.rand(B, 2, N) * 30, 100 + rs.rand(B, 1, N) * 30, rs.rand(B, 1, N) * 20], 1)
    elif kind == "dup":
        a = rs.rand(B, C, N)
        src = rs.randint(0, N, size=(B, N))
        take = rs.rand(B, N) < 0.3
        for b in range(B):
            a[b][:, take[b]] = a[b][:, src[b][take[b]]]
    else:
        raise ValueError(kind)
    return torch.from_numpy(a.astype(np.float32))


def weights(seed: int, num_classes: int, num_features: int, emb_dims: int) -> dict:
    """Canonical parameter dict (reference state_dict names, no aliases), non-trivial BN params."""
    rs = np.random.RandomState(seed)
    shapes = {
        "conv1.0.weight": (64, 2 * num_features, 1, 1), "conv2.0.weight": (64, 64, 1, 1),
        "conv3.0.weight": (64, 128, 1, 1), "conv4.0.weight": (64, 64, 1, 1),
        "conv5.0.weight": (64, 128, 1, 1), "conv6.0.weight": (emb_dims, 192, 1),
        "conv7.0.weight": (512, 1216, 1),   # in-channels hard-coded at model.py:134
        "conv8.0.weight": (256, 512, 1),
        "conv9.weight": (num_classes, 256, 1),
    }
    p = {}
    for name, shp in shapes.items():
        bound = 1.0 / np.sqrt(shp[1])
        p[name] = torch.from_numpy(rs.uniform(-bound, bound, size=shp).astype(np.float32))
    for i, c in enumerate(BN_CHANNELS, start=1):
        c = emb_dims if c is None else c
        # gamma of mixed sign exercises the max/min selection of the fused kernels
        p[f"bn{i}.weight"] = torch.from_numpy(rs.uniform(-1.0, 1.5, size=c).astype(np.float32))
        p[f"bn{i}.bias"] = torch.from_numpy(rs.uniform(-0.3, 0.3, size=c).astype(np.float32))
        p[f"bn{i}.running_mean"] = torch.from_numpy(rs.uniform(-0.2, 0.2, size=c).astype(np.float32))
        p[f"bn{i}.running_var"] = torch.from_numpy(rs.uniform(0.5, 1.5, size=c).astype(np.float32))
        p[f"bn{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return p


def cotangent(seed: int, shape) -> torch.Tensor:
    return torch.from_numpy(np.random.RandomState(seed).randn(*shape).astype(np.float32))


# (name, B, C, N, k, kind, seed)
KNN_CASES = (
    ("unit4", 2, 4, 96, 8, "unit", 11),
    ("real4", 1, 4, 257, 20, "real", 12),
    ("dup4", 2, 4, 128, 10, "dup", 13),
    ("feat64", 2, 64, 160, 16, "randn", 14),
    ("feat128", 1, 128, 100, 30, "randn", 15),
    ("xyz3", 1, 3, 50, 5, "unit", 16),
    ("kEqN", 1, 4, 24, 24, "unit", 17),
)
GRAPH_CASES = (
    ("g4", 2, 4, 40, 5, "real", 21),
    ("g64", 1, 64, 48, 7, "randn", 22),
)
# (name, B, F, N, k, classes, emb, kind, seed, train)
MODEL_CASES = (
    ("eval_small", 2, 4, 96, 8, 5, 1024, "real", 31, False),
    ("eval_readme", 1, 4, 200, 20, 5, 1024, "unit", 32, False),
    ("train_small", 2, 4, 80, 6, 5, 1024, "real", 33, True),
    # BASELINE.json configs[0] exactly: README example, B=1 F=4 N=1000 k=20 emb 1024 (README.md:71-86)
    ("eval_cfg1", 1, 4, 1000, 20, 5, 1024, "unit", 34, False),
)


# (name, B, classes, N, masked fraction, seed) — the drivers' per-point passes (train.py:180-197, test.py:119-127)
DRIVER_CASES = (
    ("loss_small", 2, 5, 64, 0.5, 41),
    ("loss_cfg", 3, 5, 700, 0.3, 42),
    ("loss_allmasked", 1, 5, 33, 1.0, 43),
    ("loss_nomask", 2, 7, 40, 0.0, 44),
)


def driver_inputs(seed: int, B: int, classes: int, N: int, frac: float):
    """logits [B,classes,N] fp32, labels [B,N] fp64 (the loaders deliver float64, data.py:181), mask [B,N] fp64 0/1."""
    rs = np.random.RandomState(seed)
    logits = torch.from_numpy((rs.randn(B, classes, N) * 2.5).astype(np.float32))
    labels = torch.from_numpy(rs.randint(0, classes, size=(B, N)).astype(np.float64))
    mask = torch.from_numpy((rs.rand(B, N) < frac).astype(np.float64))
    return logits, labels, mask
