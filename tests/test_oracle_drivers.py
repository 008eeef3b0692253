"""Pin oracle/drivers_oracle.py to the reference: tests/golden/drivers.npz was produced by exec'ing the reference's own
source lines (util.py:17-34 cal_loss, train.py:180-197 / 201 / 268-270, test.py:123-127) in make_golden.py."""
import os

import numpy as np
import pytest
import torch

import common
from oracle import drivers_oracle as D


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "drivers.npz"))


@pytest.mark.parametrize("case", common.DRIVER_CASES, ids=[c[0] for c in common.DRIVER_CASES])
def test_train_criterion_matches_reference_lines(g, case):
    name, B, C, N, frac, seed = case
    logits, labels, mask = common.driver_inputs(seed, B, C, N, frac)
    for fn in (D.train_criterion, D.train_criterion_loops):
        z = logits.clone().requires_grad_(True)
        loss, pred = fn(z, labels.clone(), mask.numpy(), C)
        loss.backward()
        assert float(loss.detach()) == float(g[f"{name}/train_loss"])                # bit-identical on this machine
        assert np.array_equal(pred.numpy(), g[f"{name}/train_pred"])
        np.testing.assert_array_equal(z.grad.numpy(), g[f"{name}/train_dlogits"])


@pytest.mark.parametrize("case", common.DRIVER_CASES, ids=[c[0] for c in common.DRIVER_CASES])
def test_validation_and_predict_match_reference_lines(g, case):
    name, B, C, N, frac, seed = case
    logits, labels, _ = common.driver_inputs(seed, B, C, N, frac)
    z = logits.clone().requires_grad_(True)
    loss, _ = D.validation_criterion(z, labels, C)
    loss.backward()
    assert float(loss.detach()) == float(g[f"{name}/val_loss"])
    np.testing.assert_array_equal(z.grad.numpy(), g[f"{name}/val_dlogits"])
    for conf in (0.8, 0.5):
        pred, vals = D.predict_labels(logits, conf)
        assert pred.dtype == torch.int64
        assert np.array_equal(pred.numpy(), g[f"{name}/pred_conf{conf}"])
    np.testing.assert_array_equal(vals.numpy(), g[f"{name}/vals"])


def test_create_train_mask_is_class_balanced():
    rs = np.random.RandomState(3)
    labels = rs.choice(5, size=700, p=[0.5, 0.2, 0.15, 0.1, 0.05]).astype(np.float64)
    m = D.create_train_mask(labels, [0, 1, 2, 3, 4], 700, np.random.RandomState(4))
    n = int(min(min(np.sum(labels == c) for c in range(5)), 700 // 5))
    assert m.shape == labels.shape and set(np.unique(m)) <= {0.0, 1.0}
    for c in range(5):
        assert int(m[labels == c].sum()) == n                                   # data.py:93-103
    m2 = D.create_train_mask(labels, [0, 1, 2, 3, 4], 700, np.random.RandomState(4), exclude_classes=[4])
    assert m2[labels == 4].sum() == 0 and int(m2[labels == 0].sum()) == int(min(min(np.sum(labels == c) for c in range(4)), 140))
