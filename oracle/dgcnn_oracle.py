"""CPU oracle for the DGCNN EdgeConv hot path.  TEST INFRASTRUCTURE ONLY.

This file is a clean-room CPU restatement (torch-on-CPU, FP32) of the arithmetic of
BenCurran98/AHN3-dgcnn.pytorch `model.py`.  It exists to *check* the CUDA path:
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product package never imports anything from `oracle/` and has no
CPU fallback.

Where the arithmetic lives: the reference is pure Python on top of PyTorch/ATen (third-party,
not vendored; reference pin pytorch=1.9.1 cpu, `requirements.txt:97`; this image has torch
2.11.0).  The restatement therefore calls the *same ATen entry points in the same order*
(`matmul`, `sum`, `topk`, `conv2d`, `batch_norm`, `leaky_relu`, `max`) so that on one machine it
is bit-identical to importing the reference.

Pinning: the reference ships no tests, fixtures or golden vectors.  The oracle is
pinned by `tests/golden/*.npz`, produced by `tests/golden/make_golden.py`, which imports the
reference's own `model.py` from /root/reference in the build container (with the one-kwarg shim
for the stray `dim9=True` at model.py:147) and records its inputs/outputs.  `tests/test_oracle.py`
checks this file against those vectors bit-for-bit (indices) / to 0 ulp-1e-6 (floats).

Reference lines followed:
  pairwise_score / knn          model.py:35-39
  graph_feature                 model.py:54-83
  layer hyper-parameters        model.py:107-141  (bias=False, LeakyReLU 0.2, BN eps 1e-5 mom 0.1)
  wiring / depth early returns  model.py:144-193
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.2      # model.py:118
BN_EPS = 1e-5          # nn.BatchNorm default used at model.py:107-114
BN_MOMENTUM = 0.1

# (conv key, bn key, is_2d) in execution order; shapes follow model.py:116-141
_EDGE_LAYERS = (("conv1", "bn1"), ("conv2", "bn2"), ("conv3", "bn3"), ("conv4", "bn4"), ("conv5", "bn5"))
_HEAD_LAYERS = (("conv6", "bn6"), ("conv7", "bn7"), ("conv8", "bn8"))


# --------------------------------------------------------------------------------------------
# kNN                                                                          model.py:25-40
# --------------------------------------------------------------------------------------------
def pairwise_score(x: torch.Tensor) -> torch.Tensor:
    """pd[b,i,j] = (-|x_j|^2 - (-2 x_i.x_j)) - |x_i|^2, rounded after every step (model.py:35-37)."""
    gram = torch.matmul(x.transpose(2, 1), x)          # [B,N,N]
    inner = -2 * gram
    sq = torch.sum(x ** 2, dim=1, keepdim=True)       # [B,1,N]
    return -sq - inner - sq.transpose(2, 1)


def knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """Indices (int64 [B,N,k]) of the k largest pd per row, self included (model.py:39)."""
    return pairwise_score(x).topk(k=k, dim=-1)[1]


def knn_chunked(x: torch.Tensor, k: int, chunk: int = 1024) -> torch.Tensor:
    """Row-chunked evaluation of the identical expression for N too large for a B*N*N matrix.

    SURVEY.md §8c measured this bit-identical to `knn()` on CPU; tests re-check it.
    """
    B, C, N = x.shape
    sq = torch.sum(x ** 2, dim=1, keepdim=True)       # [B,1,N]
    xt = x.transpose(2, 1)
    out = torch.empty(B, N, k, dtype=torch.int64)
    for s in range(0, N, chunk):
        e = min(N, s + chunk)
        inner = -2 * torch.matmul(xt[:, s:e, :], x)   # [B,c,N]
        pd = -sq - inner - sq.transpose(2, 1)[:, s:e, :]
        out[:, s:e, :] = pd.topk(k=k, dim=-1)[1]
    return out


def score_rows(x: torch.Tensor, b: int, rows: Sequence[int]) -> torch.Tensor:
    """pd rows for a few query points of sample b (used by the near-tie classifier)."""
    rows_t = torch.as_tensor(list(rows), dtype=torch.long)
    xb = x[b:b + 1]
    sq = torch.sum(xb ** 2, dim=1, keepdim=True)
    inner = -2 * torch.matmul(xb.transpose(2, 1)[:, rows_t, :], xb)
    return (-sq - inner - sq.transpose(2, 1)[:, rows_t, :])[0]


# --------------------------------------------------------------------------------------------
# Edge features                                                                model.py:43-83
# --------------------------------------------------------------------------------------------
def graph_feature(x: torch.Tensor, k: int, idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B,C,N] -> [B,2C,N,k] = cat(x_j - x_i, x_i) as a permuted view of a [B,N,k,2C] buffer.

    `idx` (int64 [B,N,k], per-sample indices) may be forced; otherwise knn(x,k) (model.py:57).
    """
    B, N = x.size(0), x.size(2)
    x = x.view(B, -1, N)
    if idx is None:
        idx = knn(x, k)
    C = x.size(1)
    flat = (idx + torch.arange(B, device=idx.device).view(-1, 1, 1) * N).view(-1)   # model.py:64-69
    rows = x.transpose(2, 1).contiguous()                              # [B,N,C]   model.py:73
    nbr = rows.view(B * N, -1)[flat, :].view(B, N, k, C)               # model.py:74-77
    ctr = rows.view(B, N, 1, C).repeat(1, 1, k, 1)                     # model.py:78
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2)      # model.py:81


# --------------------------------------------------------------------------------------------
# Model                                                                        model.py:85-193
# --------------------------------------------------------------------------------------------
def init_params(num_classes: int, num_features: int, emb_dims: int = 1024, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Canonical (alias-free) parameter/buffer dict with PyTorch's default initialisers.

    Shapes from model.py:107-141.  Keys use the reference's state_dict names (`convN.0.weight`,
    `bnN.weight` ...).  Initialisation: kaiming_uniform(a=sqrt(5)) for convs == U(-1/sqrt(fan_in), +),
    BN weight 1 / bias 0 / running_mean 0 / running_var 1.
    """
    g = torch.Generator().manual_seed(seed)
    shapes = {
        "conv1.0.weight": (64, 2 * num_features, 1, 1), "conv2.0.weight": (64, 64, 1, 1),
        "conv3.0.weight": (64, 128, 1, 1), "conv4.0.weight": (64, 64, 1, 1),
        "conv5.0.weight": (64, 128, 1, 1), "conv6.0.weight": (emb_dims, 192, 1),
        "conv7.0.weight": (512, 1216, 1), "conv8.0.weight": (256, 512, 1),
        "conv9.weight": (num_classes, 256, 1),
    }
    p: Dict[str, torch.Tensor] = {}
    for name, shp in shapes.items():
        bound = 1.0 / math.sqrt(shp[1])
        p[name] = (torch.rand(shp, generator=g) * 2 - 1) * bound
    for i, c in enumerate((64, 64, 64, 64, 64, emb_dims, 512, 256), start=1):
        p[f"bn{i}.weight"] = torch.ones(c)
        p[f"bn{i}.bias"] = torch.zeros(c)
        p[f"bn{i}.running_mean"] = torch.zeros(c)
        p[f"bn{i}.running_var"] = torch.ones(c)
        p[f"bn{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
    return p


def canonical_params(state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Drop the `convN.1.*` BN aliases of a reference state_dict (same storage as `bnN.*`)."""
    return {k: v for k, v in state_dict.items() if not (k.startswith("conv") and ".1." in k)}


class OracleDGCNN:
    """Functional re-statement of `DGCNN.forward` (model.py:144-193).

    * `params`: canonical dict (see `init_params`); tensors may require grad.
    * `training`: train-mode BN (batch statistics, running-stat update in place) and dropout.
    * `forced_idx`: optional list of up to three int64 [B,N,k] neighbour tables (layer 1..3);
      a `None` entry means "run knn" (teacher forcing).
    * after a call, `self.trace` holds idx1..3 and x1..3 for per-layer parity checks.
    """

    def __init__(self, params: Dict[str, torch.Tensor], k: int, dropout: float = 0.5, training: bool = False):
        self.p = params
        self.k = k
        self.dropout = dropout
        self.training = training
        self.trace: Dict[str, torch.Tensor] = {}

    # conv(1x1, no bias) -> BN -> LeakyReLU(0.2)                        model.py:116-139
    def _block(self, x: torch.Tensor, conv: str, bn: str) -> torch.Tensor:
        w = self.p[f"{conv}.0.weight"]
        y = F.conv2d(x, w) if w.dim() == 4 else F.conv1d(x, w)
        y = F.batch_norm(y, self.p[f"{bn}.running_mean"], self.p[f"{bn}.running_var"],
                         self.p[f"{bn}.weight"], self.p[f"{bn}.bias"],
                         self.training, BN_MOMENTUM, BN_EPS)
        if self.training:
            self.p[f"{bn}.num_batches_tracked"] += 1
        return F.leaky_relu(y, LRELU_SLOPE)

    def _edge(self, x: torch.Tensor, layer: int, forced: Optional[List[Optional[torch.Tensor]]]) -> torch.Tensor:
        idx = None
        if forced is not None and layer - 1 < len(forced):
            idx = forced[layer - 1]
        if idx is None:
            idx = knn(x.view(x.size(0), -1, x.size(2)), self.k)
        self.trace[f"idx{layer}"] = idx
        return graph_feature(x, self.k, idx)

    def forward(self, x: torch.Tensor, depth: int = 10,
                forced_idx: Optional[List[Optional[torch.Tensor]]] = None) -> torch.Tensor:
        n = x.size(2)
        e = self._edge(x, 1, forced_idx)                                   # model.py:147
        e = self._block(self._block(e, "conv1", "bn1"), "conv2", "bn2")   # :148-149
        x1 = e.max(dim=-1)[0]                                              # :150
        self.trace["x1"] = x1
        if depth == 1:
            return x1
        e = self._edge(x1, 2, forced_idx)                                  # :154
        e = self._block(self._block(e, "conv3", "bn3"), "conv4", "bn4")   # :155-156
        x2 = e.max(dim=-1)[0]                                              # :157
        self.trace["x2"] = x2
        if depth == 2:
            return x2
        e = self._edge(x2, 3, forced_idx)                                  # :161
        e = self._block(e, "conv5", "bn5")                                 # :162
        x3 = e.max(dim=-1)[0]                                              # :163
        self.trace["x3"] = x3
        if depth == 3:
            return x3
        cat = torch.cat((x1, x2, x3), dim=1)                               # :167
        if depth == 4:
            return cat
        h = self._block(cat, "conv6", "bn6")                               # :172
        if depth == 5:
            return h
        g = h.max(dim=-1, keepdim=True)[0]                                 # :175
        if depth == 6:
            return g
        h = torch.cat((g.repeat(1, 1, n), x1, x2, x3), dim=1)              # :179-180
        if depth == 7:
            return h
        h = self._block(h, "conv7", "bn7")                                 # :184
        if depth == 8:
            return h
        h = self._block(h, "conv8", "bn8")                                 # :187
        if depth == 9:
            return h
        h = F.dropout(h, self.dropout, self.training)                      # :190
        return F.conv1d(h, self.p["conv9.weight"])                         # :191

    __call__ = forward


# --------------------------------------------------------------------------------------------
# Parity helpers: neighbour *sets* and the near-tie classifier
# --------------------------------------------------------------------------------------------
def compare_neighbour_sets(x: torch.Tensor, idx_ref: torch.Tensor, idx_test: torch.Tensor,
                           tol_ulps: float = 0.0) -> Dict[str, float]:
    """Compare per-row neighbour sets; classify every differing row.

    A differing row is a *near tie* iff every index in the symmetric difference has a reference
    score pd[i,j] within `tol_ulps` ulps (ulp taken at |x_i|^2+|x_j|^2, the magnitude at which the
    reference rounds) of the reference's k-th score for that row.  tol_ulps=0 accepts exact ties only.
    Returns counts: rows, identical, near_tie, hard (hard must be 0 for parity).
    """
    B, N, k = idx_ref.shape
    a = torch.sort(idx_ref.to(torch.int64), dim=-1)[0]
    b = torch.sort(idx_test.to(torch.int64), dim=-1)[0]
    diff_rows = (a != b).any(dim=-1).nonzero()
    near = hard = 0
    worst = 0.0
    sq = torch.sum(x ** 2, dim=1)                      # [B,N]
    for bi in diff_rows[:, 0].unique().tolist():
        rows = diff_rows[diff_rows[:, 0] == bi][:, 1].tolist()
        pd = score_rows(x, bi, rows)                   # [r,N]
        for r, i in enumerate(rows):
            sa, sb = set(a[bi, i].tolist()), set(b[bi, i].tolist())
            if len(sb) != k or min(sb) < 0 or max(sb) >= N:
                hard += 1
                continue
            kth = pd[r, list(sa)].min()
            sym = list(sa ^ sb)
            mag = sq[bi, i] + sq[bi, sym]
            ulp = torch.maximum(mag, torch.tensor(1e-30)).log2().floor().sub(23).exp2()
            dev = ((pd[r, sym] - kth).abs() / ulp).max().item()
            worst = max(worst, dev)
            if dev <= tol_ulps:
                near += 1
            else:
                hard += 1
    rows_total = B * N
    return {"rows": rows_total, "identical": rows_total - diff_rows.size(0), "near_tie": near,
            "hard": hard, "worst_ulps": worst}
