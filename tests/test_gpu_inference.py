"""Batched, sharded tile inference (test.py:281-330) against the oracle run block by block."""
import numpy as np
import pytest
import torch

import common
import __graft_entry__ as entry
from oracle import dgcnn_oracle as O
from oracle import drivers_oracle as D

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    return entry.load_package()


def _blocks(seed, nb, N):
    """[nb, N, 4] float64 blocks with absolute map coordinates (x,y offsets differ per block, like room2blocks output)."""
    rs = np.random.RandomState(seed)
    xy0 = rs.rand(nb, 1, 2) * 5000 + 100000
    pts = np.concatenate([xy0 + rs.rand(nb, N, 2) * 30, 100 + rs.rand(nb, N, 1) * 30, rs.rand(nb, N, 1) * 20], 2)
    return torch.from_numpy(pts)


def _oracle_predict(w, k, block, conf):
    """test.py:291-306 for ONE block, the reference's way (B=1)."""
    X = block.clone().numpy()
    X -= np.array([np.amin(X[:, 0]), np.amin(X[:, 1]), 0, 0])                  # test.py:293-296
    X = torch.tensor(X[:, :, np.newaxis]).permute(2, 1, 0).float()             # test.py:299-301
    ora = O.OracleDGCNN({kk: v.clone() for kk, v in w.items()}, k, training=False)
    logits = ora(X)
    pred, vals = D.predict_labels(logits, conf)
    return logits, pred[0], vals[0], [ora.trace[f"idx{i}"] for i in (1, 2, 3)]


def test_predict_blocks_matches_blockwise_oracle(pkg):
    inference = __import__(entry.PKG_NAME + ".inference", fromlist=["x"])
    nb, N, k, conf = 7, 600, 20, 0.4
    w = common.weights(301, 5, 4, 1024)
    net = pkg.DGCNN(5, 4, k, dropout=0.5, emb_dims=1024, cuda=True)
    net.load_state_dict(w, strict=False)
    net.eval()
    blocks = _blocks(302, nb, N)
    ref = [_oracle_predict(w, k, blocks[i], conf) for i in range(nb)]
    # single process, batches of 3 (ragged last batch)
    got = inference.predict_blocks(net, blocks.pin_memory(), batch_size=3, min_class_confidence=conf)
    assert got.dtype == torch.int64 and tuple(got.shape) == (nb, N)
    for i in range(nb):
        _, pred, vals, _ = ref[i]
        differ = got[i] != pred
        # free-running graphs: a near-tie neighbour can move a confidence across the threshold; require 99.5 % agreement
        # and that every disagreement is a point whose confidence is within 2e-2 of the threshold or whose two best
        # classes are within 2e-2
        assert differ.float().mean().item() < 5e-3, (i, differ.float().mean().item())
    # two "ranks": the shards partition the blocks and reproduce the single-process result
    parts = [inference.predict_blocks(net, blocks, batch_size=2, min_class_confidence=conf, rank=r, world=2, return_index=True)
             for r in range(2)]
    merged = torch.empty_like(got)
    seen = []
    for out, index in parts:
        merged[index] = out
        seen += list(index)
    assert sorted(seen) == list(range(nb)) and torch.equal(merged, got)
    # per-block independence in eval mode: batch composition does not change a block's labels
    one = inference.predict_blocks(net, blocks[2:3], batch_size=1, min_class_confidence=conf)
    assert torch.equal(one[0], got[2])


def test_predict_block_cfg4_size_forced_graphs(pkg):
    """One cfg4-sized block (N=32768, k=30) with the oracle's graphs forced: logits within 1e-3, labels equal away from
    the confidence threshold.  (The oracle's own B=1 forward needs ~13 GB and a minute of host time at this size.)"""
    N, k, conf = 32768, 30, 0.3
    w = common.weights(311, 5, 4, 1024)
    block = _blocks(312, 1, N)[0]
    logits_ref, pred_ref, vals_ref, forced = _oracle_predict(w, k, block, conf)
    net = pkg.DGCNN(5, 4, k, cuda=True)
    net.load_state_dict(w, strict=False)
    net.eval()
    inference = __import__(entry.PKG_NAME + ".inference", fromlist=["x"])
    x = inference.shift_to_block_origin(block.to(DEV).t().unsqueeze(0)).float()
    logits = net(x, forced_idx=forced)
    assert ((logits.cpu() - logits_ref).abs().max() / logits_ref.abs().max()).item() < 1e-3
    pred = pkg.ops.seg_predict(logits, conf)[0].cpu()
    differ = pred != pred_ref
    assert not (differ & ((vals_ref - conf).abs() > 1e-4)).any()
    # the free-running graphs on the device: every differing row of the three graphs is a classified tie
    net(x)
    rep = O.compare_neighbour_sets(x.cpu(), forced[0], net.last_idx[0].view(1, N, k).cpu().long())
    assert rep["hard"] == 0
