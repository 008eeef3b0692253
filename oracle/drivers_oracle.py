"""CPU oracle for the drivers' per-point passes around the model.  TEST INFRASTRUCTURE ONLY (see dgcnn_oracle.py).

Restates, loop for loop, what the reference's drivers do between the model's logits and their bookkeeping:
  cal_loss                util.py:17-34      label-smoothed cross entropy (eps 0.2) — applied by the drivers to softmax
                                             OUTPUTS, so log_softmax sees probabilities (quirk kept)
  train_criterion_loops   train.py:180-197   softmax over classes, then the per-sample / per-point Python loops that build
                                             `focus_pred` [B, classes+1, N] and `focus_labels` from the mask
  train_criterion         (same)             the vectorised equivalent used for larger cases; tests check both agree
  validation_criterion    train.py:265-270
  predict_labels          test.py:119-127    permute, softmax(dim=2), max, `pred[vals < min_conf] = UNCLASSIFIED`
  create_train_mask       data.py:80-105     class-balanced point mask of one block (numpy RandomState passed in)

Deviations needed to run the reference's lines at all (train.py does not run as shipped): `mask` is taken as
a [B, N] array (upstream np.concatenate's the per-sample masks into 1-D and then indexes it as 2-D), `focus_pred` is created
on the logits' device, and `create_train_mask` receives the block's own labels instead of the whole tile's.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

UNCLASSIFIED = 31        # test.py:15


def cal_loss(pred: torch.Tensor, gold: torch.Tensor, smoothing: bool = True) -> torch.Tensor:
    gold = gold.contiguous().view(-1)                                        # util.py:20
    if smoothing:
        eps = 0.2                                                            # util.py:23
        n_class = pred.size(1)
        one_hot = torch.zeros_like(pred).scatter(1, gold.view(-1, 1), 1)     # util.py:26
        one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / (n_class - 1)  # util.py:27
        log_prb = F.log_softmax(pred, dim=1)                                 # util.py:28
        return -(one_hot * log_prb).sum(dim=1).mean()                        # util.py:30
    return F.cross_entropy(pred, gold, reduction="mean")


def train_criterion_loops(logits: torch.Tensor, labels: torch.Tensor, mask: torch.Tensor, num_classes: int):
    """train.py:180-197 with its Python loops; returns (loss, pred [B,N])."""
    labels_pred = F.softmax(logits, dim=1)                                   # :180
    focus_labels = num_classes * torch.ones_like(labels)                     # :183
    focus_pred = torch.zeros((labels_pred.shape[0], num_classes + 1, labels_pred.shape[2]), dtype=labels_pred.dtype)
    for i in range(mask.shape[0]):                                           # :185
        masked_idxs = np.where(mask[i, :])[0]
        focus_labels[i, masked_idxs] = labels[i, masked_idxs]
        focus_pred[i, :, masked_idxs] = torch.cat((labels_pred[i, :, masked_idxs], torch.zeros(1, len(masked_idxs), dtype=labels_pred.dtype)), dim=0)
        for j in range(len(mask[i, :])):                                     # :189
            if mask[i, j] == 0:
                focus_pred[i, num_classes, j] = 1
    labels_pred = focus_pred.permute(0, 2, 1).contiguous()                   # :195
    loss = cal_loss(labels_pred.view(-1, num_classes + 1), focus_labels.view(-1, 1).squeeze().long())      # :197
    return loss, labels_pred.max(dim=2)[1]                                   # :201


def train_criterion(logits: torch.Tensor, labels: torch.Tensor, mask: torch.Tensor, num_classes: int):
    """Vectorised restatement of train_criterion_loops (same tensors, no Python loops)."""
    p = F.softmax(logits, dim=1)
    on = torch.as_tensor(np.asarray(mask) != 0)                              # [B,N]
    B, C, N = p.shape
    focus_pred = torch.zeros(B, num_classes + 1, N, dtype=p.dtype)
    focus_pred[:, :num_classes, :] = p * on.unsqueeze(1)
    focus_pred[:, num_classes, :] = (~on).to(p.dtype)
    focus_labels = torch.where(on, labels.long(), torch.full_like(labels.long(), num_classes))
    lp = focus_pred.permute(0, 2, 1).contiguous()
    return cal_loss(lp.view(-1, num_classes + 1), focus_labels.view(-1)), lp.max(dim=2)[1]


def validation_criterion(logits: torch.Tensor, labels: torch.Tensor, num_classes: int):
    """train.py:265-270."""
    labels_pred = F.softmax(logits, dim=1).permute(0, 2, 1).contiguous()
    loss = cal_loss(labels_pred.view(-1, num_classes), labels.view(-1, 1).squeeze().long())
    return loss, labels_pred.max(dim=2)[1]


def predict_labels(logits: torch.Tensor, min_class_confidence: float = 0.8):
    """test.py:119-127 (and the per-block form test.py:304-308): returns (pred int64 [B,N], vals)."""
    seg_pred = logits.permute(0, 2, 1).contiguous()
    seg_pred = F.softmax(seg_pred, dim=2)
    vals, pred = seg_pred.max(dim=2)
    pred[torch.where(vals < min_class_confidence)] = UNCLASSIFIED
    return pred, vals


def create_train_mask(labels: np.ndarray, classes, tot_samples: int, rs: np.random.RandomState, exclude_classes=()):
    """data.py:80-105 for ONE block: n = min(smallest class count, tot_samples // #classes) random points of every
    (present, non-excluded) class."""
    label_counts = np.zeros(len(classes))
    for i in range(len(classes)):
        label_counts[i] = np.sum(labels == classes[i])
    min_label_count = min([label_counts[i] for i in range(len(label_counts)) if i not in exclude_classes])
    n_samples = int(min(min_label_count, np.floor(tot_samples / len(classes))))
    train_mask = np.zeros(labels.shape)
    for label in classes:
        this_label_idxs = np.where(labels == label)[0]
        if len(this_label_idxs) > 0:
            if label in exclude_classes:
                continue
            training_idxs = rs.choice(this_label_idxs, n_samples, replace=False)
            train_mask[training_idxs] = 1
    return train_mask


def sample_points(labels: np.ndarray, classes, tot_samples: int, rs: np.random.RandomState) -> np.ndarray:
    """data.py:107-143 for one tile: class-weighted sampling of `tot_samples` point indices without replacement."""
    label_counts = np.zeros(len(classes))
    for i in range(len(classes)):
        label_counts[i] = np.sum(labels == classes[i])
    tot = sum(label_counts)
    shuf_nums = np.array([tot - i for i in label_counts])
    tot = sum(shuf_nums)
    weights = shuf_nums / tot
    weights /= weights.sum()
    point_weights = weights[labels.astype(np.int64)]            # the reference fills this with a Python loop over the points
    point_weights = point_weights / point_weights.sum()
    if np.isnan(sum(point_weights)):
        point_weights = np.full(len(labels), 1.0 / len(labels))
    return rs.choice(np.arange(len(labels)), tot_samples, p=point_weights, replace=False)


def collate_pcs(data, rs: np.random.RandomState):
    """data.py:169-182: every sample is re-drawn to the smallest sample size of the batch; float64 tensors out."""
    points, labels, idxs = zip(*data)
    min_num_points = min(p.shape[0] for p in points)
    batch_points = np.zeros((len(points), min_num_points, points[0].shape[1]))
    batch_labels = np.zeros((len(points), min_num_points))
    for i in range(len(points)):
        sub = rs.choice(range(points[i].shape[0]), min_num_points, replace=False)
        batch_points[i, :, :] = points[i][sub, :]
        batch_labels[i, :] = labels[i][sub]
    return torch.tensor(batch_points), torch.tensor(batch_labels), torch.tensor(idxs)
