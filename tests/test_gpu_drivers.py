"""GPU parity of the drivers' per-point passes: dgcnn_seg_loss_f32 / dgcnn_seg_predict_f32 through
the C-ABI against the reference-made fixtures (tests/golden/drivers.npz) and the CPU oracle (oracle/drivers_oracle.py).
Tolerances: loss 2e-6 relative, gradients 1e-6 of their largest entry (FP32 exp/log on both sides), predictions bit-exact
except where the two softmax evaluations straddle the confidence threshold by < 1e-6."""
import os

import numpy as np
import pytest
import torch

import common
import __graft_entry__ as entry
from oracle import drivers_oracle as D

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    return entry.load_package()


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "drivers.npz"))


@pytest.mark.parametrize("case", common.DRIVER_CASES, ids=[c[0] for c in common.DRIVER_CASES])
def test_seg_loss_golden(pkg, g, case):
    name, B, C, N, frac, seed = case
    logits, labels, mask = common.driver_inputs(seed, B, C, N, frac)
    # ---- training form (mask), labels/mask as the loaders deliver them: float64 host tensors moved to the device
    z = logits.to(DEV).requires_grad_(True)
    loss, pred = pkg.ops.seg_loss(z, labels.to(DEV), mask.to(DEV), return_pred=True)
    loss.backward()
    ref = float(g[f"{name}/train_loss"])
    assert abs(float(loss) - ref) <= 2e-6 * abs(ref)
    assert np.array_equal(pred.cpu().numpy(), g[f"{name}/train_pred"])
    gref = torch.from_numpy(g[f"{name}/train_dlogits"])
    assert (z.grad.cpu() - gref).abs().max().item() <= 1e-6 * max(gref.abs().max().item(), 1e-12) + 1e-12
    if frac < 1.0:                                    # unmasked points carry no gradient
        off = (mask == 0).unsqueeze(1).expand(-1, C, -1)
        assert z.grad.cpu()[off].abs().max().item() == 0.0
    # ---- validation form (no mask)
    z2 = logits.to(DEV).requires_grad_(True)
    loss2 = pkg.ops.seg_loss(z2, labels.to(DEV))
    (3.0 * loss2).backward()                          # upstream gradient is honoured
    ref2 = float(g[f"{name}/val_loss"])
    assert abs(float(loss2) - ref2) <= 2e-6 * abs(ref2)
    gref2 = 3.0 * torch.from_numpy(g[f"{name}/val_dlogits"])
    assert (z2.grad.cpu() - gref2).abs().max().item() <= 1e-6 * gref2.abs().max().item()


@pytest.mark.parametrize("case", common.DRIVER_CASES, ids=[c[0] for c in common.DRIVER_CASES])
def test_seg_predict_golden(pkg, g, case):
    name, B, C, N, frac, seed = case
    logits, _, _ = common.driver_inputs(seed, B, C, N, frac)
    vals = torch.from_numpy(g[f"{name}/vals"])
    for conf in (0.8, 0.5):
        pred, c = pkg.ops.seg_predict(logits.to(DEV), conf, return_conf=True)
        assert pred.dtype == torch.int64 and tuple(pred.shape) == (B, N)
        ref = torch.from_numpy(g[f"{name}/pred_conf{conf}"].astype(np.int64))
        differ = pred.cpu() != ref
        assert not (differ & ((vals - conf).abs() > 1e-6)).any()
        assert (c.cpu() - vals).abs().max().item() < 1e-6
    # point-major view of the same logits (the layout the head produces before the final transpose)
    rows = logits.permute(0, 2, 1).contiguous().to(DEV).permute(0, 2, 1)
    assert torch.equal(pkg.ops.seg_predict(rows, 0.8), pkg.ops.seg_predict(logits.to(DEV), 0.8))


def test_seg_loss_vs_oracle_at_training_shape(pkg):
    """cfg2 shape: B=8, N=7000, 5 classes, class-balanced masks from the reference's create_train_mask."""
    B, C, N = 8, 5, 7000
    rs = np.random.RandomState(7)
    logits = torch.from_numpy((rs.randn(B, C, N) * 2).astype(np.float32))
    labels = torch.from_numpy(rs.choice(C, size=(B, N), p=[0.4, 0.3, 0.15, 0.1, 0.05]).astype(np.float64))
    mask = torch.from_numpy(np.stack([D.create_train_mask(labels[b].numpy(), list(range(C)), N, rs) for b in range(B)]))
    zr = logits.clone().requires_grad_(True)
    lref, pref = D.train_criterion(zr, labels, mask.numpy(), C)
    lref.backward()
    z = logits.to(DEV).requires_grad_(True)
    loss, pred = pkg.ops.seg_loss(z, labels.to(DEV), mask.to(DEV), return_pred=True)
    loss.backward()
    assert abs(float(loss) - float(lref)) <= 2e-6 * abs(float(lref))
    assert torch.equal(pred.cpu().long(), pref)
    assert (z.grad.cpu() - zr.grad).abs().max().item() <= 1e-6 * zr.grad.abs().max().item()
    # deterministic: fixed-order reduction
    loss_b = pkg.ops.seg_loss(logits.to(DEV), labels.to(DEV), mask.to(DEV))
    assert float(loss_b) == float(loss)


def test_seg_loss_errors(pkg):
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.ops.seg_loss(torch.zeros(1, 5, 8), torch.zeros(1, 8))
    with pytest.raises(RuntimeError):
        pkg.ops.seg_loss(torch.zeros(1, 40, 8, device=DEV), torch.zeros(1, 8, device=DEV))     # > 31 classes
