"""Pin the CPU oracle (oracle/dgcnn_oracle.py) to the reference's own outputs (tests/golden/*.npz).

The fixtures were produced by tests/golden/make_golden.py from /root/reference/model.py.
"""
import os

import numpy as np
import pytest
import torch

import common
from oracle import dgcnn_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("case", common.KNN_CASES, ids=[c[0] for c in common.KNN_CASES])
def test_knn_matches_reference_bitwise(golden_dir, case):
    name, B, C, N, k, kind, seed = case
    g = _load(golden_dir, "knn.npz")
    x = common.points(seed, B, C, N, kind)
    idx = O.knn(x, k)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (B, N, k)
    assert np.array_equal(idx.numpy(), g[f"{name}/idx"].astype(np.int64))
    # row-chunked evaluation is the same function
    assert torch.equal(O.knn_chunked(x, k, chunk=37), idx)
    # set comparison helper reports a clean bill on identical inputs
    rep = O.compare_neighbour_sets(x, idx, idx.flip(-1))
    assert rep["identical"] == rep["rows"] and rep["hard"] == 0


@pytest.mark.parametrize("case", common.GRAPH_CASES, ids=[c[0] for c in common.GRAPH_CASES])
def test_graph_feature_matches_reference(golden_dir, case):
    name, B, C, N, k, kind, seed = case
    g = _load(golden_dir, "graph_feature.npz")
    x = common.points(seed, B, C, N, kind)
    f = O.graph_feature(x, k)
    assert tuple(f.shape) == (B, 2 * C, N, k)
    assert tuple(f.stride()) == tuple(g[f"{name}/strides"])      # permuted view, channels-last
    assert np.array_equal(f.contiguous().numpy(), g[f"{name}/feature"])


def _sub(y):
    y = y.detach().reshape(-1)
    return (y if y.numel() <= 20000 else y[::37]).numpy()


@pytest.mark.parametrize("case", common.MODEL_CASES, ids=[c[0] for c in common.MODEL_CASES])
def test_model_matches_reference(golden_dir, case):
    name, B, Fd, N, k, classes, emb, kind, seed, train = case
    g = _load(golden_dir, "model.npz")
    x = common.points(seed, B, Fd, N, kind)
    p = common.weights(seed + 100, classes, Fd, emb)
    if train:
        for key, v in p.items():
            if v.is_floating_point() and "running" not in key:
                v.requires_grad_(True)
    net = O.OracleDGCNN(p, k, dropout=0.0 if train else 0.5, training=train)
    logits = net(x)
    np.testing.assert_array_equal(logits.detach().numpy(), g[f"{name}/logits"])
    for li in (1, 2, 3):
        assert np.array_equal(net.trace[f"idx{li}"].numpy(), g[f"{name}/idx{li}"].astype(np.int64))
    if train:
        cot = common.cotangent(seed + 200, tuple(logits.shape))
        (logits * cot).sum().backward()
        order = [str(s) for s in g[f"{name}/param_order"]]
        assert len(order) == 25
        for pname in order:
            got = p[pname].grad
            ref = g[f"grad/{name}/{pname}"]
            np.testing.assert_allclose(_sub(got), ref, rtol=5e-4, atol=2e-5 * float(np.abs(ref).max() + 1e-30))
        for i in range(1, 9):
            for stat in ("running_mean", "running_var"):
                np.testing.assert_allclose(p[f"bn{i}.{stat}"].numpy(), g[f"{name}/bn{i}.{stat}"], rtol=1e-6, atol=1e-7)
            assert int(p[f"bn{i}.num_batches_tracked"]) == int(g[f"{name}/bn{i}.num_batches_tracked"]) == 1
    else:
        with torch.no_grad():
            for d in range(1, 10):
                y = O.OracleDGCNN(p, k, training=False)(x, depth=d)
                assert tuple(y.shape) == tuple(g[f"{name}/depth{d}.shape"])
                np.testing.assert_array_equal(_sub(y), g[f"{name}/depth{d}"])


def test_reference_state_dict_surface(golden_dir):
    g = _load(golden_dir, "model.npz")
    keys = [str(s) for s in g["eval_small/state_keys"]]
    assert len(keys) == 89
    canon = O.canonical_params({k: None for k in keys})
    assert sorted(canon) == sorted(common.weights(0, 5, 4, 1024))


def test_near_tie_classifier_flags_real_mismatch():
    x = common.points(5, 1, 4, 64, "unit")
    idx = O.knn(x, 8)
    bad = idx.clone()
    far = O.pairwise_score(x).topk(64, dim=-1)[1][..., -1]       # the farthest point of each row
    bad[0, 3, 0] = far[0, 3]
    rep = O.compare_neighbour_sets(x, idx, bad)
    assert rep["hard"] == 1 and rep["identical"] == rep["rows"] - 1


def test_near_tie_classifier_accepts_exact_ties():
    x = common.points(6, 1, 4, 40, "unit")
    x[0, :, 7] = x[0, :, 9]                                        # duplicate point -> equal score columns
    idx = O.knn(x, 6)
    swapped = idx.clone()
    has7 = (idx == 7).any(-1)
    has9 = (idx == 9).any(-1)
    only7 = has7 & ~has9
    if only7.any():
        swapped[only7] = torch.where(swapped[only7] == 7, torch.full_like(swapped[only7], 9), swapped[only7])
        rep = O.compare_neighbour_sets(x, idx, swapped)
        assert rep["hard"] == 0 and rep["near_tie"] == int(only7.sum())
