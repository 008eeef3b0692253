"""Device input pipeline against the reference's sampler restated in oracle/drivers_oracle.py
(data.py:107-143, 169-182).  Random sampling cannot be compared draw by draw (different generators), so parity is
(1) exact structural properties of every draw and (2) the sampling DISTRIBUTION: per-class selection counts averaged over
many draws must agree with numpy's `choice(replace=False, p)` within the Monte-Carlo error."""
import numpy as np
import pytest
import torch

import __graft_entry__ as entry
from oracle import drivers_oracle as D

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pipe():
    entry.load_package()
    return __import__(entry.PKG_NAME + ".input_pipeline", fromlist=["x"])


def _tile(seed, n, probs, F=4):
    rs = np.random.RandomState(seed)
    labels = rs.choice(len(probs), size=n, p=probs).astype(np.float64)
    pts = np.concatenate([rs.rand(n, 2) * 30 + 120000.0, 100 + rs.rand(n, 1) * 30, rs.rand(n, 1) * 20], 1)[:, :F]
    return np.concatenate([pts, labels[:, None]], 1)


def test_sample_structure_and_determinism(pipe):
    tile = _tile(1, 50000, [0.55, 0.25, 0.12, 0.06, 0.02])
    tiles = pipe.DeviceTiles([tile], num_classes=5, device=DEV)
    idx = tiles.sample_indices(0, 7000, seed=11).cpu().numpy()
    assert idx.shape == (7000,) and idx.min() >= 0 and idx.max() < 50000
    assert np.unique(idx).size == 7000                               # without replacement
    assert np.all(np.diff(idx[:-1]) > 0)                             # index order (the key at the threshold comes last)
    assert np.array_equal(idx, tiles.sample_indices(0, 7000, seed=11).cpu().numpy())       # pure function of (labels, seed)
    assert not np.array_equal(idx, tiles.sample_indices(0, 7000, seed=12).cpu().numpy())
    # n == N: every point exactly once
    assert np.array_equal(np.sort(tiles.sample_indices(0, 50000, seed=3).cpu().numpy()), np.arange(50000))
    # a class that is absent never appears; a single-class tile falls back to uniform (data.py:137-138)
    one = _tile(2, 5000, [1.0, 0.0, 0.0])
    t1 = pipe.DeviceTiles([one], num_classes=3, device=DEV)
    assert t1.sample_indices(0, 1000, seed=1).numel() == 1000


def test_sampling_distribution_matches_numpy_choice(pipe):
    probs = [0.6, 0.25, 0.1, 0.05]
    tile = _tile(5, 6000, probs)
    labels = tile[:, -1]
    tiles = pipe.DeviceTiles([tile], num_classes=4, device=DEV)
    n, draws = 1500, 60
    ours = np.zeros(4)
    ref = np.zeros(4)
    rs = np.random.RandomState(9)
    for d in range(draws):
        idx = tiles.sample_indices(0, n, seed=1000 + d).cpu().numpy()
        ours += np.bincount(labels[idx].astype(int), minlength=4)
        ridx = D.sample_points(labels, [0, 1, 2, 3], n, rs)
        ref += np.bincount(labels[ridx].astype(int), minlength=4)
    ours /= draws
    ref /= draws
    # rare classes are over-sampled exactly like the reference does: means agree within 3 % (Monte-Carlo error ~1 %)
    assert np.all(np.abs(ours - ref) <= 0.03 * ref + 3.0), (ours, ref)
    # and it is NOT uniform sampling: the rarest class is drawn well above its population share
    assert ours[3] / n > 1.4 * probs[3]


def test_batch_layout_and_values(pipe):
    tiles_np = [_tile(20 + i, 9000 + 500 * i, [0.5, 0.3, 0.2]) for i in range(3)]
    tiles = pipe.DeviceTiles(tiles_np, num_classes=3, device=DEV)
    x, y, ids = tiles.batch([2, 0, 1, 2], 4096, seed=5)
    assert x.dtype == torch.float32 and tuple(x.shape) == (4, 4, 4096) and x.is_contiguous()
    assert y.dtype == torch.int64 and tuple(y.shape) == (4, 4096) and ids.tolist() == [2, 0, 1, 2]
    # every column is a row of its tile, cast like `.float()` does (train.py:176) — the absolute map coordinates survive
    for b, t in enumerate([2, 0, 1, 2]):
        idx = tiles.sample_indices(t, 4096, seed=5 * 1000003 + b).cpu().numpy()
        ref = torch.tensor(tiles_np[t][idx, :4]).t().float()
        assert torch.equal(x[b].cpu(), ref)
        assert np.array_equal(y[b].cpu().numpy(), tiles_np[t][idx, 4].astype(np.int64))
    # the same tile drawn twice in one batch gives two different samples
    assert not torch.equal(x[0], x[3])
    # the batch feeds the model as is
    pkg = entry.load_package()
    net = pkg.DGCNN(3, 4, 10, cuda=True).eval()
    assert tuple(net(x[:, :, :600].contiguous()).shape) == (4, 3, 600)


def test_collate_oracle_shapes():
    rs = np.random.RandomState(0)
    data = [(rs.rand(50, 4), rs.randint(0, 3, 50).astype(np.float64), 7), (rs.rand(40, 4), rs.randint(0, 3, 40).astype(np.float64), 8)]
    p, l, i = D.collate_pcs(data, rs)
    assert tuple(p.shape) == (2, 40, 4) and p.dtype == torch.float64 and tuple(l.shape) == (2, 40) and i.tolist() == [7, 8]
