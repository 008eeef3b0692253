"""Drop-in replacement for the reference `model.py` (BenCurran98/AHN3-dgcnn.pytorch).

Same public surface, so train.py / test.py / visualisation.py keep their call sites:
  * `knn(x[B,C,N], k) -> LongTensor[B,N,k]`                                    (reference model.py:25)
  * `get_graph_feature(x, k=20, use_cuda=True) -> Tensor[B,2C,N,k]`            (reference model.py:43)
  * `DGCNN(num_classes, num_features, k, dropout=0.5, emb_dims=1024, cuda=False)`, `model(x, depth=10)`
    with the reference's 89-key state_dict and parameter order                 (reference model.py:85-193)

Everything on the EdgeConv hot path (kNN, gather, edge features, conv1-5 + BN + LeakyReLU + max over k,
and their backward) runs in hand-written sm_100a kernels via libdgcnn_b200.so; the B x N x N distance matrix
and the B x 2C x N x k edge tensor are never materialised.  There is no CPU path: inputs must be CUDA tensors.

Deliberate compatibility extensions:
  * `cuda=True` moves the module to the current CUDA device (the reference only stores the flag).
  * the mistyped keyword `ropout=` used at reference train.py:103 is accepted as an alias of `dropout=`.
  * the stray `dim9=` keyword of reference model.py:147 is accepted and ignored by `get_graph_feature`.
  * `forward(..., forced_idx=[idx1, idx2, idx3])` (keyword-only) teacher-forces neighbour tables for parity tests;
    the tables actually used are kept in `self.last_idx`.
"""
from __future__ import annotations

import os
import sys
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

_PKG = "ahn3_dgcnn_pytorch_b200"
if not __package__:
    # Imported the way the reference's drivers import it: this directory is on sys.path and they say
    # `from model import DGCNN` (train.py:12, test.py:11, visualisation.py).  Load the real package (its directory name is
    # not an identifier, so it is registered under _PKG) and make `model` an alias of its `model` submodule: one set of
    # classes, one loaded library.  The rest of this file then runs as a discarded duplicate with relative imports bound
    # to the package.
    if _PKG not in sys.modules:
        import importlib.util as _ilu
        _dir = os.path.dirname(os.path.abspath(__file__))
        _spec = _ilu.spec_from_file_location(_PKG, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
        _mod = _ilu.module_from_spec(_spec)
        sys.modules[_PKG] = _mod
        try:
            _spec.loader.exec_module(_mod)
        except BaseException:
            sys.modules.pop(_PKG, None)
            raise
    sys.modules[__name__] = sys.modules[_PKG + ".model"]
    __package__ = _PKG
    __spec__ = sys.modules[_PKG + ".model"].__spec__

from . import ops

LRELU_SLOPE = 0.2
_EDGE_WIDTH = 64


def knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """k nearest neighbours (self included) of every point in feature space; int64 [B,N,k]."""
    B, Cn, N = x.shape
    idx = ops.knn_rows(ops.pack_rows(x), B, Cn, N, k)
    return idx.view(B, N, k).long()


def get_graph_feature(x: torch.Tensor, k: int = 20, use_cuda: bool = True, **_compat) -> torch.Tensor:
    """cat(x_j - x_i, x_i) over the kNN graph as a [B,2C,N,k] permuted view (channels-last storage)."""
    B, N = x.size(0), x.size(2)
    x = x.reshape(B, -1, N)
    Cn = x.size(1)
    rows = ops.pack_rows(x)
    idx = ops.knn_rows(rows, B, Cn, N, k)
    return ops.graph_feature_rows(rows, idx, B, Cn, N, k).permute(0, 3, 1, 2)


def _bn_state(bn: nn.Module) -> "ops.BNState":
    """Running statistics + (eps, momentum) of a BatchNorm module; momentum=None is nn.BatchNorm's cumulative moving
    average, i.e. an update factor of 1 / (batches seen including this one)."""
    mom = bn.momentum
    if mom is None:
        mom = 1.0 / float(int(bn.num_batches_tracked) + 1) if bn.track_running_stats else 0.0
    return ops.BNState(bn.running_mean, bn.running_var, bn.eps, mom)


def _conv_bn_act(conv: nn.Module, bn: nn.Module) -> nn.Sequential:
    return nn.Sequential(conv, bn, nn.LeakyReLU(negative_slope=LRELU_SLOPE))


class DGCNN(nn.Module):
    """Dynamic Graph CNN for per-point segmentation; B200-native EdgeConv path."""

    # (in-channel multiplier of num_features or absolute, out) for conv1..conv5; conv6-9 below
    def __init__(self, num_classes, num_features, k, dropout=0.5, emb_dims=1024, cuda=False, **compat):
        super().__init__()
        if "ropout" in compat:                      # reference train.py:103 typo
            dropout = compat.pop("ropout")
        if compat:
            raise TypeError(f"DGCNN.__init__() got unexpected keyword arguments {sorted(compat)}")
        self.k = k
        self.num_classes = num_classes
        self.use_cuda = cuda
        self.num_features = num_features
        self.emb_dims = emb_dims
        self.last_idx: List[torch.Tensor] = []

        # Registration order fixes parameters() order and state_dict keys: bn1..bn8 first, then the
        # Sequentials (which alias the same BN modules as convN.1), then dp1, conv9.
        widths = (64, 64, 64, 64, 64, emb_dims, 512, 256)
        for i, w in enumerate(widths, start=1):
            setattr(self, f"bn{i}", nn.BatchNorm2d(w) if i <= 5 else nn.BatchNorm1d(w))
        edge_in = (2 * num_features, 64, 128, 64, 128)
        for i, cin in enumerate(edge_in, start=1):
            setattr(self, f"conv{i}", _conv_bn_act(nn.Conv2d(cin, _EDGE_WIDTH, kernel_size=1, bias=False),
                                                   getattr(self, f"bn{i}")))
        head = ((192, emb_dims), (1216, 512), (512, 256))      # conv7 in-channels are fixed at 1216 upstream
        for i, (cin, cout) in enumerate(head, start=6):
            setattr(self, f"conv{i}", _conv_bn_act(nn.Conv1d(cin, cout, kernel_size=1, bias=False),
                                                   getattr(self, f"bn{i}")))
        self.dp1 = nn.Dropout(p=dropout)
        self.conv9 = nn.Conv1d(256, num_classes, kernel_size=1, bias=False)
        if cuda:
            self.to(torch.device("cuda"))

    # ------------------------------------------------------------------------------------------
    def _edge_block(self, rows: torch.Tensor, idx: torch.Tensor, B: int, Cn: int, N: int,
                    first: int, second: Optional[int]) -> torch.Tensor:
        conv_a, bn_a = getattr(self, f"conv{first}")[0], getattr(self, f"bn{first}")
        conv_b = bn_b = None
        if second is not None:
            conv_b, bn_b = getattr(self, f"conv{second}")[0], getattr(self, f"bn{second}")
        st_a = _bn_state(bn_a)
        st_b = None if bn_b is None else _bn_state(bn_b)
        if self.training:
            out = ops.EdgeConvTrain.apply(rows, idx, (B, Cn, N, self.k), conv_a.weight, bn_a.weight, bn_a.bias,
                                          None if conv_b is None else conv_b.weight,
                                          None if bn_b is None else bn_b.weight,
                                          None if bn_b is None else bn_b.bias, st_a, st_b)
            ops.bump_batches(bn_a)
            if bn_b is not None:
                ops.bump_batches(bn_b)
            return out
        aff_a = ops.eval_affine(bn_a.weight.detach(), bn_a.bias.detach(), st_a)
        aff_b = None if bn_b is None else ops.eval_affine(bn_b.weight.detach(), bn_b.bias.detach(), st_b)
        return ops.edgeconv_eval(rows, idx, B, Cn, N, self.k, conv_a.weight,
                                 aff_a, None if conv_b is None else conv_b.weight, aff_b)

    def _graph(self, rows: torch.Tensor, B: int, Cn: int, N: int, layer: int,
               forced: Optional[Sequence[Optional[torch.Tensor]]]) -> torch.Tensor:
        idx = None
        if forced is not None and layer - 1 < len(forced) and forced[layer - 1] is not None:
            idx = forced[layer - 1].to(device=rows.device, dtype=torch.int32).reshape(B * N, self.k).contiguous()
        if idx is None:
            idx = ops.knn_rows(rows.detach(), B, Cn, N, self.k, ordered=ops.DETERMINISTIC)     # the blocks only need the SET
        self.last_idx.append(idx)
        return idx

    @staticmethod
    def _bcn(rows: torch.Tensor, B: int, N: int) -> torch.Tensor:
        """point-major [B*N, C] -> the reference's [B, C, N] (contiguous)."""
        return rows.view(B, N, -1).permute(0, 2, 1).contiguous()

    def _dense(self, h: torch.Tensor, i: int) -> torch.Tensor:
        conv, bn = getattr(self, f"conv{i}")[0], getattr(self, f"bn{i}")
        return F.leaky_relu(bn(ops.linear(h, conv.weight.squeeze(-1))), LRELU_SLOPE)

    def forward(self, x, depth=10, *, forced_idx=None):
        if not x.is_cuda:
            raise RuntimeError("DGCNN (B200 build) needs CUDA tensors: there is no CPU fallback on this path")
        if x.dim() != 3 or x.size(1) != self.num_features:
            raise RuntimeError(f"expected input [B, {self.num_features}, N], got {tuple(x.shape)}")
        if self.training:
            ops.defer_batches()
            try:
                return self._forward(x, depth, forced_idx)
            finally:
                ops.flush_batches()          # num_batches_tracked of every BatchNorm passed: one launch
        with torch.no_grad():            # the inference kernels are not differentiable
            return self._forward(x, depth, forced_idx)

    def _forward(self, x, depth, forced_idx):
        B, Fdim, N = x.shape
        self.last_idx = []
        ops.clear_shared_planes()
        rows0 = ops.pack_rows(x.float() if x.dtype != torch.float32 else x)
        x1 = self._edge_block(rows0, self._graph(rows0, B, Fdim, N, 1, forced_idx), B, Fdim, N, 1, 2)
        if depth == 1:
            return self._bcn(x1, B, N)
        x2 = self._edge_block(x1, self._graph(x1, B, 64, N, 2, forced_idx), B, 64, N, 3, 4)
        if depth == 2:
            return self._bcn(x2, B, N)
        x3 = self._edge_block(x2, self._graph(x2, B, 64, N, 3, forced_idx), B, 64, N, 5, None)
        if depth == 3:
            return self._bcn(x3, B, N)

        # ---- shared MLP + segmentation head, point-major dense GEMMs (reference model.py:167-193) ----
        feat = torch.cat((x1, x2, x3), dim=1)                              # [M,192]
        if depth == 4:
            return self._bcn(feat, B, N)
        if depth == 5:
            return self._bcn(self._dense(feat, 6), B, N)                   # [M,emb] only materialised on request
        # conv6 + bn6 + LeakyReLU + max over the points fused: the [B,emb,N] activation is never formed
        bn6 = self.bn6
        if N >= 32:
            g = ops.GlobalFeature.apply(feat, self.conv6[0].weight.squeeze(-1), bn6.weight, bn6.bias, _bn_state(bn6),
                                        self.training, B, N)               # [B,emb] global feature
            if self.training:
                ops.bump_batches(bn6)
        else:                            # blocks shorter than one epilogue row group: materialise [M,emb] (tiny)
            g = self._dense(feat, 6).view(B, N, -1).max(dim=1)[0]
        if depth == 6:
            return g.unsqueeze(-1)
        if depth == 7:
            return torch.cat((g.unsqueeze(-1).expand(-1, -1, N), self._bcn(feat, B, N)), dim=1)
        w7 = self.conv7[0].weight.squeeze(-1)                              # [512, 1216]
        E = g.size(1)
        if E + feat.size(1) != w7.size(1):
            raise RuntimeError(f"conv7 expects {w7.size(1)} input channels but emb_dims+192 = {E + feat.size(1)}")
        # conv7([g ; x1,x2,x3]) = W7[:, :E] g (once per block) + W7[:, E:] feat  — no repeat/cat of the global feature
        # the per-block term is added inside the fused BatchNorm + LeakyReLU passes
        if self.training and depth > 8 and ops.FUSED_HEAD and self.bn7.track_running_stats and self.bn8.track_running_stats:
            h8 = ops.head_mlp(feat, ops.linear(g, w7[:, :E]), w7[:, E:], self.bn7, self.conv8[0].weight.squeeze(-1), self.bn8, N)
        else:
            h7 = ops.bn_lrelu_rows(ops.linear(feat, w7[:, E:]), self.bn7, self.training, N, bias=ops.linear(g, w7[:, :E]))
            if depth == 8:
                return self._bcn(h7, B, N)
            h8 = ops.bn_lrelu_rows(ops.linear(h7, self.conv8[0].weight.squeeze(-1)), self.bn8, self.training, N)
        if depth == 9:
            return self._bcn(h8, B, N)
        logits = ops.linear(self.dp1(h8), self.conv9.weight.squeeze(-1))   # [M,classes]
        return self._bcn(logits, B, N)
