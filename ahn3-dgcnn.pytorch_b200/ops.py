"""Host-side operators over the C-ABI (include/dgcnn_b200.h).

PyTorch is used for device memory, streams and autograd bookkeeping only; every per-edge computation
runs in libdgcnn_b200.so.  All tensors must live on a CUDA device — there is no CPU path.

Internal layout: features are point-major "rows" [B*N, C] (row p = b*N + n), neighbour tables are
int32 [B*N, k] with block-local indices.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib

COUT = 64
TC_KNN_MIN_N = 256       # below this every key is a candidate anyway and the exact tile kernel is as fast
EDGECONV_IMPL = 0        # dgcnn_edgeconv_args.impl: 0 = default (pipelined tcgen05), 1 = FP32 FMA, 2 = serial tcgen05 (A/B checks)
DETERMINISTIC = os.environ.get("DGCNN_DETERMINISTIC", "0") == "1"   # bit-reproducible gradients: no float atomics anywhere (slower, see set_deterministic)
THIN_LINEAR = True       # inputs <= 8 channels (the input block's point-level GEMM): streaming FP32-FMA kernels
NARROW_LINEAR = True     # outputs <= 8 (classifier): streaming FP32-FMA kernel instead of padded tensor-core tiles
ATOMIC_WGRAD = True      # row-contraction (weight-gradient) GEMMs add their row splits into D atomically; False = partial tiles + fixed-order reduce
FUSED_HEAD = True        # training: conv7 -> bn7 -> conv8 -> bn8 as one autograd node, activations between BN and GEMM only as planes
FUSED_BWD = True         # two-conv blocks: conv_b backward with the scatter fused in (no dz_a[E,64]); False = the two-pass kernels
# bf16 planes per operand in the GRADIENT GEMMs (forward GEMMs always use 3 = FP32-accurate).  2 = 3 MMAs per product,
# relative error ~2^-17; 3 = 6 MMAs, ~2^-24 like the forward.  DGCNN_GRAD_NSPLIT overrides; bench.py reports both.
GRAD_NSPLIT = int(os.environ.get("DGCNN_GRAD_NSPLIT", "2"))
BN_EPS_DEFAULT = 1e-5


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _dev_index(dev: torch.device) -> int:
    return torch.cuda.current_device() if dev.index is None else dev.index


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (the B200 path has no CPU fallback), got {t.device}")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32, got {t.dtype}")


class KernelLog:
    """Counts kernel launches made through the C-ABI and (optionally) brackets each call with CUDA events on
    the launching stream, so bench.py can report per-kernel device time measured inside the timed region."""

    def __init__(self):
        self.launches = 0
        self.timing = False
        self.events = []          # (name, start_event, end_event)

    def reset(self, timing: bool = False):
        self.launches = 0
        self.timing = timing
        self.events = []

    def summary(self):
        """name -> (calls, total_ms); call after torch.cuda.synchronize()."""
        out = {}
        for name, a, b in self.events:
            c, t = out.get(name, (0, 0.0))
            out[name] = (c + 1, t + a.elapsed_time(b))
        return out


LOG = KernelLog()

# kernels launched by each entry point
_LAUNCHES = {"dgcnn_knn_f32": 2, "dgcnn_knn_tc_f32": 6, "dgcnn_knn_small_f32": 3, "dgcnn_edgeconv_bwd_b": 3, "dgcnn_edgeconv_bwd_b_fused": 3,
             "dgcnn_tc_gemm_f32": 3, "dgcnn_tc_gemm_bn_lrelu_segmax_f32": 5, "dgcnn_rows_bn_lrelu_fwd": 3, "dgcnn_rows_bn_lrelu_bwd": 3,
             "dgcnn_rows_colsum_f32": 2, "dgcnn_seg_loss_f32": 2}


def _call(name: str, dev: torch.device, *args, tag: str = "") -> None:
    """Enqueue one C-ABI entry point on `dev`'s current stream (appended as the last argument).

    The library launches on the CURRENT device, so a tensor living on another device than the caller's current one
    (`model.to('cuda:1')` without `set_device`) needs the guard; PyTorch's own ops behave the same way."""
    fn = getattr(_lib.load(), name)
    LOG.launches += _LAUNCHES.get(name, 1)
    index = _dev_index(dev)
    guard = torch.cuda.device(index) if index != torch.cuda.current_device() else contextlib.nullcontext()
    with guard:
        st = C.c_void_p(torch.cuda.current_stream(index).cuda_stream)
        if LOG.timing:
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args, st)
            b.record()
            LOG.events.append((name + tag, a, b))
        else:
            rc = fn(*args, st)
    _lib.check(rc)


# per-device scratch for the per-CTA partial reductions
_scratch = {}


def _partials(dev: torch.device) -> Tuple[torch.Tensor, torch.Tensor]:
    # one buffer per (device, stream): kernels of different streams never share scratch
    key = (_dev_index(dev), torch.cuda.current_stream(_dev_index(dev)).cuda_stream)
    if key not in _scratch:
        lib = _lib.load()
        _scratch[key] = (
            torch.empty(lib.dgcnn_edgeconv_partials_bytes() // 8, dtype=torch.float64, device=dev),
            torch.empty(lib.dgcnn_edgeconv_dw_partials_bytes() // 4, dtype=torch.float32, device=dev),
        )
    return _scratch[key]


def _rows_view(rows: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """Accept [M, C] with unit inner stride, 16-byte aligned rows; return (tensor, ld)."""
    if rows.dim() != 2 or rows.stride(1) != 1 or rows.stride(0) % 4 or rows.data_ptr() % 16:
        rows = rows.contiguous()
        if rows.size(1) % 4:
            pad = 4 - rows.size(1) % 4
            rows = torch.nn.functional.pad(rows, (0, pad))
    return rows, rows.stride(0)


# ------------------------------------------------------------------------------------------------
# layout
# ------------------------------------------------------------------------------------------------
def _pack_rows(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x, "pack_rows")
    B, Cn, N = x.shape
    c4 = (Cn + 3) // 4 * 4
    rows = torch.empty(B * N, c4, dtype=torch.float32, device=x.device)
    _call("dgcnn_pack_rows_f32", x.device, _ptr(x), x.stride(0), x.stride(1), x.stride(2), B, Cn, N, _ptr(rows), c4)
    return rows


class PackRows(torch.autograd.Function):
    """pack_rows with its adjoint, so that gradients reach the input point cloud like in the reference."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return _pack_rows(x)

    @staticmethod
    def backward(ctx, drows):
        B, Cn, N = ctx.shape
        return unpack_rows(drows, B, Cn, N)


def pack_rows(x: torch.Tensor) -> torch.Tensor:
    """x [B,C,N] (any strides) -> rows [B*N, round4(C)], padding columns zero (model.py:73)."""
    if x.requires_grad and torch.is_grad_enabled():
        return PackRows.apply(x)
    return _pack_rows(x)


def unpack_rows(rows: torch.Tensor, B: int, Cn: int, N: int) -> torch.Tensor:
    """rows [B*N, >=C] -> contiguous [B,C,N]."""
    rows, ld = _rows_view(rows)
    y = torch.empty(B, Cn, N, dtype=torch.float32, device=rows.device)
    _call("dgcnn_unpack_rows_f32", rows.device, _ptr(rows), ld, B, Cn, N, _ptr(y))
    return y


# ------------------------------------------------------------------------------------------------
# kNN / graph feature
# ------------------------------------------------------------------------------------------------
def knn_rows(rows: torch.Tensor, B: int, Cn: int, N: int, k: int, method: str = "auto",
             stats: Optional[list] = None, ordered: bool = True) -> torch.Tensor:
    """Neighbour table int32 [B*N, k] of point-major rows (model.py:25-40).

    method: "exact" = FP32-FMA tile kernel (any C <= 128, any k <= 128); "tc" = tcgen05 filter + exact re-score
    (C <= 128, k <= 48, N >= 128); "small" = thread-per-query kernel (C <= 16, k <= 48, N >= 512); all three give the
    identical table; "auto" picks the fastest supported one.  `stats` (tc only): receives the 5 fallback counters.
    ordered=False (tc only): the same neighbour sets in unspecified order — all the EdgeConv blocks need, and faster."""
    _need_cuda(rows, "knn")
    rows, ld = _rows_view(rows)
    lib = _lib.load()
    idx = torch.empty(B * N, k, dtype=torch.int32, device=rows.device)
    if method == "auto":
        method = "exact"
        if k <= N:
            if N >= TC_KNN_MIN_N and lib.dgcnn_knn_tc_supported(Cn, N, k):
                method = "tc"
            elif lib.dgcnn_knn_small_supported(Cn, N, k):
                method = "small"
    if method in ("tc", "small"):
        nbytes = getattr(lib, f"dgcnn_knn_{method}_workspace_bytes")(B, N)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=rows.device)
        if method == "tc":
            _call("dgcnn_knn_tc_f32", rows.device, _ptr(rows), ld, B, Cn, N, k, 1 if ordered else 0, _ptr(idx), _ptr(ws), nbytes,
                  tag=f"[C={Cn}]")
        else:
            _call("dgcnn_knn_small_f32", rows.device, _ptr(rows), ld, B, Cn, N, k, _ptr(idx), _ptr(ws), nbytes)
        if stats is not None and method == "tc":
            out5 = (C.c_int * 5)()
            _lib.check(lib.dgcnn_knn_tc_fallback_stats(_ptr(ws), B, N, Cn, out5,
                                                       C.c_void_p(torch.cuda.current_stream(_dev_index(rows.device)).cuda_stream)))
            stats[:] = list(out5)
    else:
        norms = torch.empty(B * N, dtype=torch.float32, device=rows.device)
        _call("dgcnn_knn_f32", rows.device, _ptr(rows), ld, B, Cn, N, k, _ptr(idx), _ptr(norms))
    return idx


def graph_feature_rows(rows: torch.Tensor, idx: torch.Tensor, B: int, Cn: int, N: int, k: int) -> torch.Tensor:
    """[B,N,k,2C] edge-feature buffer (model.py:74-81); API-parity path, not used by DGCNN.forward."""
    rows, ld = _rows_view(rows)
    out = torch.empty(B, N, k, 2 * Cn, dtype=torch.float32, device=rows.device)
    _call("dgcnn_graph_feature_f32", rows.device, _ptr(rows), ld, _ptr(idx), B, Cn, N, k, _ptr(out))
    return out


# ------------------------------------------------------------------------------------------------
# dense GEMMs of the shared MLP / head on the tensor cores (tcgen05, FP32-accurate bf16x3 split)
# ------------------------------------------------------------------------------------------------
_gemm_ws = {}


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """Grow-only GEMM workspace per (device, stream)."""
    key = (_dev_index(dev), torch.cuda.current_stream(_dev_index(dev)).cuda_stream)
    ws = _gemm_ws.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25), dtype=torch.uint8, device=dev)
        _gemm_ws[key] = ws
    return ws


def tc_gemm(a: torch.Tensor, b: torch.Tensor, nsplit: int = 3) -> torch.Tensor:
    """D[M,N] = a[M,K] @ b[N,K]^T in FP32 via dgcnn_tc_gemm_f32; a and b may be arbitrary 2-D strided views."""
    M, K = a.shape
    N = b.shape[0]
    lib = _lib.load()
    ws = _workspace(a.device, lib.dgcnn_tc_gemm_workspace_bytes(M, N, K, nsplit))
    d = torch.empty(M, N, dtype=torch.float32, device=a.device)
    _call("dgcnn_tc_gemm_f32", a.device, _ptr(a), a.stride(0), a.stride(1), _ptr(b), b.stride(0), b.stride(1), _ptr(d), N, M, N, K,
          nsplit, _ptr(ws), ws.numel())
    return d


class Planes:
    """An FP32 matrix [R, K] in the tensor-core operand format (nsplit bf16 planes, see include/dgcnn_b200.h)."""

    __slots__ = ("buf", "R", "K", "nsplit")

    def __init__(self, x: Optional[torch.Tensor], nsplit: int, shape: Optional[Tuple[int, int]] = None, device=None):
        """Planes(x, nsplit) splits x; Planes(None, nsplit, shape=(R, K), device=dev) only allocates (a kernel fills them)."""
        lib = _lib.load()
        self.R, self.K = x.shape if x is not None else shape
        self.nsplit = nsplit
        nbytes = lib.dgcnn_tc_planes_bytes(self.R, self.K, nsplit)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=x.device if x is not None else device)
        if x is not None:
            _call("dgcnn_tc_split_planes_f32", x.device, _ptr(x), x.stride(0), x.stride(1), self.R, self.K, nsplit, _ptr(self.buf), nbytes)


def gemm_planes(a: Planes, b: Planes, nsplit: int, transposed: bool = False) -> torch.Tensor:
    """transposed=False: D[a.R, b.R] = A B^T (contraction over the K columns).
    transposed=True:  D[a.K, b.K] = A^T B  (contraction over the rows: weight gradients, no re-split)."""
    lib = _lib.load()
    dev = a.buf.device
    if transposed:
        assert a.R == b.R
        M, N, K, mode = a.K, b.K, a.R, (2 if ATOMIC_WGRAD else 1)
    else:
        assert a.K == b.K
        M, N, K, mode = a.R, b.R, a.K, 0
    ws = _workspace(dev, lib.dgcnn_tc_gemm_planes_workspace_bytes(mode, M, N, K))
    d = torch.empty(M, N, dtype=torch.float32, device=dev)
    _call("dgcnn_tc_gemm_planes_f32", dev, mode, _ptr(a.buf), a.K, a.nsplit, _ptr(b.buf), b.K, b.nsplit, nsplit, _ptr(d), N, M, N, K,
          _ptr(ws), ws.numel())
    return d


# Planes of activations that several autograd nodes of ONE step consume (feat: the head node's forward and the global
# feature's backward).  Keyed by the tensor's memory and version; the model clears it at the start of every forward, and the
# producers stay alive in the autograd graph until then, so an address cannot be reused under a live entry.
_shared_planes = {}


def _planes_key(x: torch.Tensor):
    return (x.data_ptr(), tuple(x.shape), tuple(x.stride()), x._version, str(x.device))


def share_planes(x: torch.Tensor, planes: Planes) -> None:
    _shared_planes[_planes_key(x)] = planes


def shared_planes(x: torch.Tensor, nsplit: int) -> Planes:
    """The planes registered for x if they hold at least nsplit planes, else a fresh split."""
    pl = _shared_planes.get(_planes_key(x))
    return pl if pl is not None and pl.nsplit >= nsplit else Planes(x, nsplit)


def clear_shared_planes() -> None:
    _shared_planes.clear()


# `num_batches_tracked += 1` of every BatchNorm a training forward went through (nn.BatchNorm's bookkeeping): collected and
# applied as ONE multi-tensor launch at the end of the forward instead of eight scalar kernels.
_pending_batches = []
_defer_batches = False


def defer_batches() -> None:
    """From now until flush_batches() the counters are collected instead of bumped one by one (DGCNN.forward)."""
    global _defer_batches
    _defer_batches = True


def bump_batches(bn_mod) -> None:
    if bn_mod.num_batches_tracked is None:
        return
    if _defer_batches:
        _pending_batches.append(bn_mod.num_batches_tracked)
    else:
        bn_mod.num_batches_tracked.add_(1)


def flush_batches() -> None:
    global _defer_batches
    _defer_batches = False
    if _pending_batches:
        torch._foreach_add_(_pending_batches, 1)
        _pending_batches.clear()


class TcLinear(torch.autograd.Function):
    """y = x @ w^T (x [M,K], w [N,K]) with all three GEMMs (y, dx, dw) on the tensor cores.

    The operands are split into bf16 planes once: x for the forward (3 planes; the first GRAD_NSPLIT are reused by the
    weight gradient, which contracts over the ROWS of the same planes), dy once for both gradient GEMMs."""

    @staticmethod
    def forward(ctx, x, w):
        xp = Planes(x, 3)
        ctx.xp = xp
        ctx.save_for_backward(w)
        return gemm_planes(xp, Planes(w, 3), 3)

    @staticmethod
    def backward(ctx, dy):
        (w,) = ctx.saved_tensors
        xp = ctx.xp
        dx = dw = None
        # gradient GEMMs: GRAD_NSPLIT planes (2: relative error 2^-17, two orders below the gradient noise floor; 3: FP32-accurate)
        dyp = Planes(dy, GRAD_NSPLIT)
        if ctx.needs_input_grad[0]:
            dx = gemm_planes(dyp, Planes(w.t(), GRAD_NSPLIT), GRAD_NSPLIT)           # [M,K] = dy[M,N] @ w[N,K]
        if ctx.needs_input_grad[1]:
            dw = gemm_planes(dyp, xp, GRAD_NSPLIT, transposed=True)                    # [N,K] = dy^T[N,M] @ x[M,K]
        ctx.xp = None
        return dx, dw


class GlobalFeature(torch.autograd.Function):
    """g[B,E] = max_n LeakyReLU(BN(feat W^T)) per block (conv6 + bn6 + max, model.py:172-175), never materialising [M,E].

    Backward uses the linearity of the dense batch-statistics-BN gradient dy = a dz - b' - kappa y (dz is non-zero only
    at the B*E arg-max entries, y = feat W^T):
        dfeat = scatter(a dz W) - b' W - feat (W^T diag(kappa) W)
        dW    = a (dz^T feat) - b' colsum(feat)^T - diag(kappa) W (feat^T feat)
    i.e. two 192x192-wide GEMMs instead of three E-wide ones."""

    @staticmethod
    def forward(ctx, feat, w, gamma, beta, bn: BNState, training: bool, B: int, N: int):
        lib = _lib.load()
        M, K = feat.shape
        E = w.shape[0]
        dev = feat.device
        f32 = dict(dtype=torch.float32, device=dev)
        ws = _workspace(dev, lib.dgcnn_tc_gemm_colstats_workspace_bytes(M, E, K, B))
        g = torch.empty(B, E, **f32)
        ysel = torch.empty(B, E, **f32)
        nstar = torch.empty(B, E, dtype=torch.int32, device=dev)
        stat = torch.empty(2, E, **f32)
        fd, wd = feat.detach(), w.detach()
        _call("dgcnn_tc_gemm_bn_lrelu_segmax_f32", dev, _ptr(fd), fd.stride(0), fd.stride(1), _ptr(wd), wd.stride(0), wd.stride(1),
              M, E, K, N, _ptr(gamma.detach()), _ptr(beta.detach()), bn.eps, bn.momentum, 1 if training else 0,
              _ptr(bn.running_mean), _ptr(bn.running_var), _ptr(stat[0]), _ptr(stat[1]), _ptr(g), _ptr(ysel), _ptr(nstar),
              _ptr(ws), ws.numel())
        ctx.save_for_backward(fd, wd, gamma.detach(), beta.detach(), stat, ysel, nstar)
        ctx.training = training
        return g

    @staticmethod
    def backward(ctx, dg):
        feat, w, gamma, beta, stat, ysel, nstar = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError("GlobalFeature: the inference form (running statistics) is not differentiable")
        M = feat.shape[0]
        B, E = ysel.shape
        K = feat.shape[1]
        if K % 4 or feat.stride(1) != 1 or w.stride(1) != 1:
            raise RuntimeError("GlobalFeature: feature width must be a multiple of 4 with unit inner stride")
        dev = feat.device
        f32 = dict(dtype=torch.float32, device=dev)
        # per-channel coefficients in one launch: gdz [B,E] (dz at the arg-max entries), coef = a | kappa | b' | s1 | s2,
        # and -diag(kappa) W, the left operand of the small matrix below
        gdz_c = torch.empty(B, E, **f32)
        coef = torch.empty(5, E, **f32)
        wneg = torch.empty(E, K, **f32)
        dg = dg.contiguous()
        _call("dgcnn_global_feature_bwd_coef_f32", dev, _ptr(dg), _ptr(ysel), _ptr(stat[0]), _ptr(stat[1]), _ptr(gamma), _ptr(beta),
              _ptr(w), w.stride(0), B, E, K, M, 0.2, _ptr(gdz_c), _ptr(coef), _ptr(wneg))
        a_c, bprime, s1, s2 = coef[0], coef[2], coef[3], coef[4]
        # dense parts on the tensor cores: dfeat = -feat (W^T diag(kappa) W) - b' W  (the sign is folded into the small matrix)
        gram_neg = tc_gemm(wneg.t(), w.t(), nsplit=GRAD_NSPLIT)                          # [K,K] = -W^T diag(kappa) W (symmetric)
        featp = shared_planes(feat, GRAD_NSPLIT)                                         # the head node's planes of feat, or one split for both GEMMs
        dfeat = gemm_planes(featp, Planes(gram_neg, GRAD_NSPLIT), GRAD_NSPLIT)
        cvec = torch.empty(K, **f32)
        _call("dgcnn_vecmat_f32", dev, _ptr(bprime), _ptr(w), w.stride(0), E, K, -1.0, _ptr(cvec))
        dfeat.add_(cvec)
        # sparse parts in one pass: scatter of a dz W into dfeat, and t1 = dz^T feat
        t1 = torch.empty(E, K, **f32)
        if DETERMINISTIC:        # t1 only, then the scatter with one owner warp per target row
            _call("dgcnn_global_feature_bwd_sparse_f32", dev, _ptr(feat), feat.stride(0), _ptr(w), w.stride(0), _ptr(nstar), _ptr(gdz_c),
                  _ptr(a_c), B, E, K, _ptr(None), dfeat.stride(0), _ptr(t1))
            _call("dgcnn_global_feature_scatter_det_f32", dev, _ptr(w), w.stride(0), _ptr(nstar), _ptr(gdz_c), _ptr(a_c), B, E, K,
                  _ptr(dfeat), dfeat.stride(0))
        else:
            _call("dgcnn_global_feature_bwd_sparse_f32", dev, _ptr(feat), feat.stride(0), _ptr(w), w.stride(0), _ptr(nstar), _ptr(gdz_c),
                  _ptr(a_c), B, E, K, _ptr(dfeat), dfeat.stride(0), _ptr(t1))
        # dW = a dz^T feat - b' colsum(feat)^T - diag(kappa) W (feat^T feat)
        f2 = gemm_planes(featp, featp, GRAD_NSPLIT, transposed=True)            # [K,K] = feat^T feat (row contraction, split over the points)
        wf2 = tc_gemm(w, f2, nsplit=GRAD_NSPLIT)                                # f2 is symmetric: W f2 = W f2^T
        colsum = rows_colsum(feat, M).contiguous()                              # [1,K]
        dw = torch.empty(E, K, **f32)
        _call("dgcnn_global_feature_bwd_dw_f32", dev, _ptr(coef), _ptr(t1), _ptr(colsum), _ptr(wf2), E, K, _ptr(dw))
        return dfeat, dw, s2, s1, None, None, None, None


def rows_colsum(x: torch.Tensor, N: int) -> torch.Tensor:
    """x[M,C] -> [M/N, C]: column sums of every block of N rows (deterministic FP64 two-pass reduction)."""
    lib = _lib.load()
    M, Cn = x.shape
    if Cn % 4 or Cn > 1024 or M % N:
        return x.view(-1, N, Cn).sum(1)
    x = x.contiguous()
    out = torch.empty(M // N, Cn, dtype=torch.float32, device=x.device)
    ws = torch.empty(lib.dgcnn_rows_colsum_workspace_bytes(M, Cn, N), dtype=torch.uint8, device=x.device)
    _call("dgcnn_rows_colsum_f32", x.device, _ptr(x), M, Cn, N, _ptr(out), _ptr(ws), ws.numel())
    return out


class RowsBnLrelu(torch.autograd.Function):
    """h = LeakyReLU(BatchNorm1d(y + bias[block])) on point-major rows y[M,C] (model.py:131-139,179-184), two fused
    passes forward and two backward instead of the library's five and eight."""

    @staticmethod
    def forward(ctx, y, bias, gamma, beta, bn: BNState, training: bool, N: int):
        lib = _lib.load()
        M, Cn = y.shape
        dev = y.device
        y = y.contiguous()
        bias_c = None if bias is None else bias.detach().contiguous()
        ws = torch.empty(lib.dgcnn_rows_bn_workspace_bytes(Cn), dtype=torch.uint8, device=dev)
        stat = torch.empty(2, Cn, dtype=torch.float32, device=dev)
        aff = torch.empty(3 * Cn, dtype=torch.float32, device=dev)
        h = torch.empty_like(y)
        _call("dgcnn_rows_bn_lrelu_fwd", dev, _ptr(y.detach()), M, Cn, _ptr(bias_c), N, _ptr(gamma.detach()), _ptr(beta.detach()),
              bn.eps, bn.momentum, 1 if training else 0, _ptr(bn.running_mean), _ptr(bn.running_var), _ptr(stat[0]),
              _ptr(stat[1]), _ptr(aff), _ptr(h), _ptr(ws), ws.numel())
        ctx.save_for_backward(y.detach(), bias_c, gamma.detach(), stat, aff)
        ctx.training, ctx.N, ctx.ws = training, N, ws
        return h

    @staticmethod
    def backward(ctx, dh):
        y, bias, gamma, stat, aff = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError("RowsBnLrelu: the inference form (running statistics) is not differentiable")
        M, Cn = y.shape
        dh = dh.contiguous()
        dy = torch.empty_like(y)
        dgamma = torch.empty(Cn, dtype=torch.float32, device=y.device)
        dbeta = torch.empty_like(dgamma)
        _call("dgcnn_rows_bn_lrelu_bwd", y.device, _ptr(dh), _ptr(y), M, Cn, _ptr(bias), ctx.N, _ptr(gamma), _ptr(stat[1]), _ptr(aff),
              _ptr(dy), _ptr(dgamma), _ptr(dbeta), _ptr(ctx.ws), ctx.ws.numel())
        dbias = None if bias is None else rows_colsum(dy, ctx.N)
        return dy, dbias, dgamma, dbeta, None, None, None


def bn_lrelu_rows(y: torch.Tensor, bn_mod, training: bool, N: int, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LeakyReLU(bn_mod(y + bias[block])) for y[M,C]; `bn_mod` is the reference's nn.BatchNorm1d (state and parameters)."""
    mom = bn_mod.momentum
    if mom is None:                      # cumulative moving average (nn.BatchNorm with momentum=None)
        mom = 1.0 / float(int(bn_mod.num_batches_tracked) + 1)
    st = BNState(bn_mod.running_mean, bn_mod.running_var, bn_mod.eps, mom)
    h = RowsBnLrelu.apply(y, bias, bn_mod.weight, bn_mod.bias, st, training, N)
    if training:
        bump_batches(bn_mod)
    return h


class HeadMlp(torch.autograd.Function):
    """h8 = LReLU(bn8(conv8(LReLU(bn7(conv7_points(feat) + bias7[block]))))) (model.py:176-184, training form) as ONE node, so
    that every tensor between a BatchNorm pass and the GEMM next to it lives only as operand planes:
      forward   bn7's apply pass writes h7 as 3 planes (conv8's A operand; its first planes are conv8's weight-gradient operand)
      backward  bn8's / bn7's apply passes write dy8 / dy7 as GRAD_NSPLIT planes (operands of the data- and weight-gradient
                GEMMs); the per-block bias gradient comes out of bn7's apply pass.
    Saves three split passes, a column-sum pass and the FP32 copies of h7, dy7, dy8 (~0.9 GB of HBM traffic per step)."""

    @staticmethod
    def forward(ctx, feat, bias7, w7, g7, b7, w8, g8, b8, bn7: BNState, bn8: BNState, N: int):
        lib = _lib.load()
        dev = feat.device
        M = feat.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        fp = Planes(feat.detach(), 3)
        share_planes(feat.detach(), fp)                  # GlobalFeature.backward contracts over the same planes
        y7 = gemm_planes(fp, Planes(w7.detach(), 3), 3)
        C7 = y7.shape[1]
        bias_c = None if bias7 is None else bias7.detach().contiguous()
        ws7 = torch.empty(lib.dgcnn_rows_bn_workspace_bytes(C7), dtype=torch.uint8, device=dev)
        stat7, aff7 = torch.empty(2, C7, **f32), torch.empty(3 * C7, **f32)
        h7p = Planes(None, 3, shape=(M, C7), device=dev)
        _call("dgcnn_rows_bn_lrelu_fwd_planes", dev, _ptr(y7), M, C7, _ptr(bias_c), N, _ptr(g7.detach()), _ptr(b7.detach()), bn7.eps,
              bn7.momentum, _ptr(bn7.running_mean), _ptr(bn7.running_var), _ptr(stat7[0]), _ptr(stat7[1]), _ptr(aff7), _ptr(h7p.buf),
              h7p.buf.numel(), 3, _ptr(ws7), ws7.numel())
        y8 = gemm_planes(h7p, Planes(w8.detach(), 3), 3)
        C8 = y8.shape[1]
        ws8 = torch.empty(lib.dgcnn_rows_bn_workspace_bytes(C8), dtype=torch.uint8, device=dev)
        stat8, aff8 = torch.empty(2, C8, **f32), torch.empty(3 * C8, **f32)
        h8 = torch.empty_like(y8)
        _call("dgcnn_rows_bn_lrelu_fwd", dev, _ptr(y8), M, C8, _ptr(None), N, _ptr(g8.detach()), _ptr(b8.detach()), bn8.eps, bn8.momentum, 1,
              _ptr(bn8.running_mean), _ptr(bn8.running_var), _ptr(stat8[0]), _ptr(stat8[1]), _ptr(aff8), _ptr(h8), _ptr(ws8), ws8.numel())
        ctx.save_for_backward(y7, bias_c, w7.detach(), g7.detach(), stat7, aff7, y8, w8.detach(), g8.detach(), stat8, aff8)
        ctx.fp, ctx.h7p, ctx.ws7, ctx.ws8, ctx.N = fp, h7p, ws7, ws8, N
        return h8

    @staticmethod
    def backward(ctx, dh8):
        y7, bias7, w7, g7, stat7, aff7, y8, w8, g8, stat8, aff8 = ctx.saved_tensors
        fp, h7p, N = ctx.fp, ctx.h7p, ctx.N
        dev = y7.device
        M, C7 = y7.shape
        C8 = y8.shape[1]
        G = GRAD_NSPLIT
        f32 = dict(dtype=torch.float32, device=dev)
        dh8 = dh8.contiguous()
        dg8, db8 = torch.empty(C8, **f32), torch.empty(C8, **f32)
        dy8p = Planes(None, G, shape=(M, C8), device=dev)
        _call("dgcnn_rows_bn_lrelu_bwd_planes", dev, _ptr(dh8), _ptr(y8), M, C8, _ptr(None), N, _ptr(g8), _ptr(stat8[1]), _ptr(aff8),
              _ptr(dy8p.buf), dy8p.buf.numel(), G, _ptr(dg8), _ptr(db8), _ptr(None), _ptr(None), 0, _ptr(ctx.ws8), ctx.ws8.numel())
        dh7 = gemm_planes(dy8p, Planes(w8.t(), G), G)                      # [M,C7] = dy8 @ w8
        dw8 = gemm_planes(dy8p, h7p, G, transposed=True)                   # [C8,C7] = dy8^T h7
        dg7, db7 = torch.empty(C7, **f32), torch.empty(C7, **f32)
        dy7p = Planes(None, G, shape=(M, C7), device=dev)
        dbias7 = csw = None
        if bias7 is not None and ctx.needs_input_grad[1]:
            dbias7 = torch.empty(M // N, C7, **f32)
            csw = torch.empty(_lib.load().dgcnn_rows_colsum_workspace_bytes(M, C7, N), dtype=torch.uint8, device=dev)
        _call("dgcnn_rows_bn_lrelu_bwd_planes", dev, _ptr(dh7), _ptr(y7), M, C7, _ptr(bias7), N, _ptr(g7), _ptr(stat7[1]), _ptr(aff7),
              _ptr(dy7p.buf), dy7p.buf.numel(), G, _ptr(dg7), _ptr(db7), _ptr(dbias7), _ptr(csw), 0 if csw is None else csw.numel(),
              _ptr(ctx.ws7), ctx.ws7.numel())
        dfeat = gemm_planes(dy7p, Planes(w7.t(), G), G) if ctx.needs_input_grad[0] else None
        dw7 = gemm_planes(dy7p, fp, G, transposed=True)
        ctx.fp = ctx.h7p = None
        return dfeat, dbias7, dw7, dg7, db7, dw8, dg8, db8, None, None, None


def _bn_state_of(bn_mod) -> BNState:
    mom = bn_mod.momentum
    if mom is None:                      # cumulative moving average (nn.BatchNorm with momentum=None)
        mom = 1.0 / float(int(bn_mod.num_batches_tracked) + 1)
    return BNState(bn_mod.running_mean, bn_mod.running_var, bn_mod.eps, mom)


def head_mlp(feat: torch.Tensor, bias7: Optional[torch.Tensor], w7: torch.Tensor, bn7, w8: torch.Tensor, bn8, N: int) -> torch.Tensor:
    """Training-mode conv7(points part) + bn7 + LReLU + conv8 + bn8 + LReLU on point-major rows (see HeadMlp)."""
    h8 = HeadMlp.apply(feat, bias7, w7, bn7.weight, bn7.bias, w8, bn8.weight, bn8.bias, _bn_state_of(bn7), _bn_state_of(bn8), N)
    bump_batches(bn7)
    bump_batches(bn8)
    return h8


class LinearNarrow(torch.autograd.Function):
    """y = x @ w^T for at most 8 outputs (the classifier, model.py:141): streaming FP32-FMA passes (csrc/linear_narrow.cu)."""

    @staticmethod
    def forward(ctx, x, w):
        xd, wd = x.detach(), w.detach()
        if xd.stride(1) != 1 or xd.stride(0) % 4 or xd.data_ptr() % 16:
            xd = xd.contiguous()
        if wd.stride(1) != 1:
            wd = wd.contiguous()
        M, K = xd.shape
        NO = wd.shape[0]
        y = torch.empty(M, NO, dtype=torch.float32, device=xd.device)
        _call("dgcnn_linear_narrow_fwd_f32", xd.device, _ptr(xd), xd.stride(0), _ptr(wd), wd.stride(0), M, K, NO, _ptr(y), NO)
        ctx.save_for_backward(xd, wd)
        return y

    @staticmethod
    def backward(ctx, dy):
        xd, wd = ctx.saved_tensors
        M, K = xd.shape
        NO = wd.shape[0]
        dy = dy.contiguous()
        dx = torch.empty(M, K, dtype=torch.float32, device=xd.device) if ctx.needs_input_grad[0] else None
        dw = torch.empty(NO, K, dtype=torch.float32, device=xd.device)
        _call("dgcnn_linear_narrow_bwd_f32", xd.device, _ptr(xd), xd.stride(0), _ptr(dy), NO, _ptr(wd), wd.stride(0), M, K, NO,
              _ptr(dx), K, _ptr(dw))
        return dx, dw


def linear(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """x @ w^T without a library GEMM, whatever the shape: tensor cores (tcgen05, FP32-accurate split), or the streaming
    FP32-FMA kernel when the output is a handful of columns (the classifier: tensor-core tiles would be 96 % padding)."""
    _need_cuda(x, "linear")
    K, NO = x.shape[1], w.shape[0]
    if NARROW_LINEAR and NO <= 8 and K % 4 == 0 and K <= (256 if NO > 4 else 512):
        return LinearNarrow.apply(x, w)
    return TcLinear.apply(x, w)


# ------------------------------------------------------------------------------------------------
# EdgeConv block
# ------------------------------------------------------------------------------------------------
@dataclass
class BNState:
    """The non-differentiable part of one BatchNorm2d(64) (model.py:107-111)."""
    running_mean: Optional[torch.Tensor]
    running_var: Optional[torch.Tensor]
    eps: float = BN_EPS_DEFAULT
    momentum: float = 0.1


def _args(rows, ld, idx, B, Cn, N, k, w_a, w_b, out, ld_out, v_scratch=None, impl=None):
    return _lib.EdgeConvArgs(rows.data_ptr(), ld, idx.data_ptr(), B, Cn, N, k, w_a.data_ptr(),
                             0 if w_b is None else w_b.data_ptr(), 0 if out is None else out.data_ptr(), ld_out,
                             0 if v_scratch is None else v_scratch.data_ptr(), EDGECONV_IMPL if impl is None else impl)


def _v_scratch(B: int, Cn: int, N: int, dev) -> Optional[torch.Tensor]:
    """Scratch for the per-point centre term; its presence switches the 64-channel blocks to the tensor-core kernels."""
    return torch.empty(B * N, COUT, dtype=torch.float32, device=dev) if Cn == COUT else None


def _use_lin(Cn: int, w_b) -> bool:
    """Single-conv block on 64-channel features with the default kernel family: no per-edge GEMM (csrc/edgeconv_lin.cu)."""
    return w_b is None and Cn == COUT and EDGECONV_IMPL == 0


def _lin_weight(wa: torch.Tensor, Cn: int) -> torch.Tensor:
    """[Wa1 ; Wa2 - Wa1] ([128, C]): X @ this^T = [U | V] with y[(i,j)] = U_j + V_i."""
    wlin = torch.empty(2 * COUT, Cn, dtype=torch.float32, device=wa.device)
    _call("dgcnn_edgeconv_lin_weight_f32", wa.device, _ptr(wa), Cn, _ptr(wlin))
    return wlin


def set_deterministic(on: bool = True) -> None:
    """Bit-reproducible training step: every scatter becomes a fixed-order gather (csrc/graph_transpose.cu: the kNN table is
    transposed once per graph), the weight-gradient GEMMs reduce their row splits in a fixed order, and the layers whose
    streaming kernels accumulate with atomics go back to the tensor-core path; the model asks for its neighbour tables in the
    reference's best-first order (the order inside a neighbour SET depends on the low bits of the filter's channel means,
    which are summed with float atomics).  About 1.5x the step time; the forward pass is
    unchanged (it has no atomics whose order matters: the segmented max uses integer atomicMax)."""
    global DETERMINISTIC, FUSED_BWD, ATOMIC_WGRAD, NARROW_LINEAR, THIN_LINEAR
    DETERMINISTIC = on
    FUSED_BWD = ATOMIC_WGRAD = NARROW_LINEAR = THIN_LINEAR = not on


def _scatter_r_deterministic(a, idx, B, N, k, dy_edges, r, dev) -> None:
    """r[j] = sum of dy_edges over the in-edges of j in ascending edge id (CSR transpose of the kNN table)."""
    lib = _lib.load()
    M = B * N
    in_start = torch.empty(M + 1, dtype=torch.int32, device=dev)
    in_edge = torch.empty(M * k, dtype=torch.int32, device=dev)
    ws = torch.empty(lib.dgcnn_graph_transpose_workspace_bytes(B, N), dtype=torch.uint8, device=dev)
    _call("dgcnn_graph_transpose", dev, _ptr(idx), B, N, k, _ptr(in_start), _ptr(in_edge), _ptr(ws), ws.numel())
    _call("dgcnn_gather_in_edges_f32", dev, _ptr(dy_edges), _ptr(in_start), _ptr(in_edge), M, _ptr(r), r.stride(0))


def _uv_gemm(rows: torch.Tensor, ld: int, Cn: int, wlin: torch.Tensor, keep_planes: bool):
    """[U | V] = X [Wa1 ; Wa2 - Wa1]^T ([M,128]) and, for the tensor-core form, the planes of X (reused by the weight gradient).
    The 4-channel input block goes through the streaming thin-input kernel: K = 4 would be padded to a 64-wide atom."""
    M = rows.shape[0]
    if THIN_LINEAR and Cn <= 8 and _lib.load().dgcnn_linear_thin_supported(Cn, 2 * COUT):
        uv = torch.empty(M, 2 * COUT, dtype=torch.float32, device=rows.device)
        _call("dgcnn_linear_thin_fwd_f32", rows.device, _ptr(rows), ld, _ptr(wlin), wlin.stride(0), M, Cn, 2 * COUT, _ptr(uv), 2 * COUT)
        return uv, None
    if not keep_planes:
        return tc_gemm(rows[:, :Cn], wlin), None
    xp = Planes(rows[:, :Cn], 3)
    return gemm_planes(xp, Planes(wlin, 3), 3), xp


def eval_affine(gamma: torch.Tensor, beta: torch.Tensor, bn: BNState) -> torch.Tensor:
    """float[192] = running_mean | gamma/sqrt(running_var+eps) | beta."""
    return torch.cat((bn.running_mean, gamma * torch.rsqrt(bn.running_var + bn.eps), beta)).float().contiguous()


def edgeconv_eval(rows: torch.Tensor, idx: torch.Tensor, B: int, Cn: int, N: int, k: int,
                  w_a: torch.Tensor, aff_a: torch.Tensor,
                  w_b: Optional[torch.Tensor] = None, aff_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused inference EdgeConv block -> [B*N, 64].  Not differentiable."""
    _need_cuda(rows, "edgeconv_eval")
    rows, ld = _rows_view(rows)
    out = torch.empty(B * N, COUT, dtype=torch.float32, device=rows.device)
    wa = w_a.detach().reshape(COUT, -1).contiguous()
    if _use_lin(Cn, w_b):
        # single conv, 64 channels: y = U_j + V_i from ONE point-level GEMM, then a gather pass for the extrema over k
        M, dev = B * N, rows.device
        uv, _ = _uv_gemm(rows, ld, Cn, _lin_weight(wa, Cn), False)
        ymax, ymin = torch.empty(M, COUT, dtype=torch.float32, device=dev), torch.empty(M, COUT, dtype=torch.float32, device=dev)
        amax, amin = torch.empty(M, COUT, dtype=torch.uint8, device=dev), torch.empty(M, COUT, dtype=torch.uint8, device=dev)
        a = _args(rows, ld, idx, B, Cn, N, k, wa, None, None, COUT)
        _call("dgcnn_edgeconv_lin_extrema", dev, C.byref(a), _ptr(uv), uv.stride(0), _ptr(uv[:, COUT:]), uv.stride(0), _ptr(None),
              _ptr(ymax), _ptr(ymin), _ptr(amax), _ptr(amin))
        ysel, arg = torch.empty_like(ymax), torch.empty_like(amax)
        _call("dgcnn_edgeconv_select", dev, M, _ptr(ymax), _ptr(ymin), _ptr(amax), _ptr(amin), _ptr(aff_a), _ptr(out), COUT,
              _ptr(ysel), _ptr(arg))
        return out
    wb = None if w_b is None else w_b.detach().reshape(COUT, COUT).contiguous()
    if wb is not None and EDGECONV_IMPL == 0:
        # two convs: the training-forward passes with the running statistics as affines — conv_a as y_a = U_j + V_i
        # (point-level GEMM + gather pass), conv_b on the warp-specialised tcgen05 pipeline, extrema over k, BN_b +
        # LeakyReLU on the selected extremum.  Faster than the fully fused per-edge kernel (which runs BOTH convolutions
        # as per-edge MMAs) even though y_a takes a round trip through HBM.
        M, dev = B * N, rows.device
        partials, _ = _partials(dev)
        uv, _ = _uv_gemm(rows, ld, Cn, _lin_weight(wa, Cn), False)
        y_a = torch.empty(M * k, COUT, dtype=torch.float32, device=dev)
        ymax, ymin = torch.empty(M, COUT, dtype=torch.float32, device=dev), torch.empty(M, COUT, dtype=torch.float32, device=dev)
        amax, amin = torch.empty(M, COUT, dtype=torch.uint8, device=dev), torch.empty(M, COUT, dtype=torch.uint8, device=dev)
        a = _args(rows, ld, idx, B, Cn, N, k, wa, wb, None, COUT)
        _call("dgcnn_edgeconv_lin_conv_a", dev, C.byref(a), _ptr(uv), uv.stride(0), _ptr(uv[:, COUT:]), uv.stride(0), _ptr(y_a), _ptr(partials))
        _call("dgcnn_edgeconv_conv_b_train", dev, C.byref(a), _ptr(y_a), _ptr(aff_a), _ptr(partials), _ptr(ymax), _ptr(ymin), _ptr(amax),
              _ptr(amin))
        ysel, arg = torch.empty_like(ymax), torch.empty_like(amax)
        _call("dgcnn_edgeconv_select", dev, M, _ptr(ymax), _ptr(ymin), _ptr(amax), _ptr(amin), _ptr(aff_b), _ptr(out), COUT, _ptr(ysel),
              _ptr(arg))
        return out
    vs = _v_scratch(B, Cn, N, rows.device)
    a = _args(rows, ld, idx, B, Cn, N, k, wa, wb, out, COUT, vs)
    _call("dgcnn_edgeconv_fwd_eval", rows.device, C.byref(a), _ptr(aff_a), _ptr(aff_b))
    return out


class EdgeConvTrain(torch.autograd.Function):
    """Training-mode EdgeConv block (batch-statistics BN) with a hand-written backward.

    forward(x_rows[M,C], idx, meta, w_a, gamma_a, beta_a, w_b|None, gamma_b|None, beta_b|None, bn_a, bn_b)
    """

    @staticmethod
    def forward(ctx, x, idx, meta, w_a, gamma_a, beta_a, w_b, gamma_b, beta_b, bn_a: BNState, bn_b: Optional[BNState]):
        lib = _lib.load()
        B, Cn, N, k = meta
        M, E = B * N, B * N * k
        dev = x.device
        rows, ld = _rows_view(x.detach())
        two = w_b is not None
        wa = w_a.detach().reshape(COUT, 2 * Cn).contiguous()
        wb = w_b.detach().reshape(COUT, COUT).contiguous() if two else None
        partials, _ = _partials(dev)

        f32 = dict(dtype=torch.float32, device=dev)
        lin = _use_lin(Cn, w_b)
        y_a = None if lin else torch.empty(E, COUT, **f32)
        ymax = torch.empty(M, COUT, **f32)
        ymin = torch.empty(M, COUT, **f32)
        amax = torch.empty(M, COUT, dtype=torch.uint8, device=dev)
        amin = torch.empty(M, COUT, dtype=torch.uint8, device=dev)
        out = torch.empty(M, COUT, **f32)
        ysel = torch.empty(M, COUT, **f32)
        arg = torch.empty(M, COUT, dtype=torch.uint8, device=dev)
        stat_a = torch.empty(2 + 3, COUT, **f32)     # mean, rstd, affine(mean|scale|beta)
        vs = None if (lin or EDGECONV_IMPL == 0) else _v_scratch(B, Cn, N, dev)
        a = _args(rows, ld, idx, B, Cn, N, k, wa, wb, None, COUT, vs)

        uv = uv2 = fused = xp = wlin = None
        if two and EDGECONV_IMPL == 0:
            # first conv of a two-conv block: y_a = U_j + V_i written per edge (BN_a needs batch statistics before conv_b
            # can start), no per-edge GEMM.  Also for the input block: y_a carries the centre term Wa2 x_i (~|W||x|) either
            # way, so forming x_j - x_i first buys no accuracy at the level of y_a (both are within a few 1e-8 |W||x|).
            wlin = _lin_weight(wa, Cn)
            uv2, xp = _uv_gemm(rows, ld, Cn, wlin, True)     # X planes: this GEMM now, the weight gradient in the backward
            if FUSED_BWD and GRAD_NSPLIT == 2 and k >= 8 and N >= 16:
                # + the out-edge sums of y_a: with them (and the in-edge sums the backward accumulates) dz_a[E,64] is never formed
                fused = torch.empty(M, COUT, **f32)
                _call("dgcnn_edgeconv_lin_conv_a_train", dev, C.byref(a), _ptr(uv2), uv2.stride(0), _ptr(uv2[:, COUT:]), uv2.stride(0),
                      _ptr(y_a), _ptr(partials), _ptr(fused))
            else:
                _call("dgcnn_edgeconv_lin_conv_a", dev, C.byref(a), _ptr(uv2), uv2.stride(0), _ptr(uv2[:, COUT:]), uv2.stride(0), _ptr(y_a),
                      _ptr(partials))
        elif lin:
            # y = U_j + V_i: one point-level GEMM + a gather pass (extrema over k, batch sums); nothing per-edge in HBM
            wlin = _lin_weight(wa, Cn)
            uv, xp = _uv_gemm(rows, ld, Cn, wlin, True)
            _call("dgcnn_edgeconv_lin_extrema", dev, C.byref(a), _ptr(uv), uv.stride(0), _ptr(uv[:, COUT:]), uv.stride(0), _ptr(partials),
                  _ptr(ymax), _ptr(ymin), _ptr(amax), _ptr(amin))
        else:
            _call("dgcnn_edgeconv_conv_a_train", dev, C.byref(a), _ptr(y_a), _ptr(partials), _ptr(ymax), _ptr(ymin),
                                                       _ptr(amax), _ptr(amin))
        _call("dgcnn_bn_finalize", dev, _ptr(partials), COUT, float(E), _ptr(gamma_a.detach()), _ptr(beta_a.detach()),
                                         bn_a.eps, bn_a.momentum, _ptr(bn_a.running_mean), _ptr(bn_a.running_var),
                                         _ptr(stat_a[0]), _ptr(stat_a[1]), _ptr(stat_a[2:]))
        stat_b = None
        if two:
            stat_b = torch.empty(2 + 3, COUT, **f32)
            _call("dgcnn_edgeconv_conv_b_train", dev, C.byref(a), _ptr(y_a), _ptr(stat_a[2:]), _ptr(partials),
                                                       _ptr(ymax), _ptr(ymin), _ptr(amax), _ptr(amin))
            _call("dgcnn_bn_finalize", dev, _ptr(partials), COUT, float(E), _ptr(gamma_b.detach()), _ptr(beta_b.detach()),
                                             bn_b.eps, bn_b.momentum, _ptr(bn_b.running_mean), _ptr(bn_b.running_var),
                                             _ptr(stat_b[0]), _ptr(stat_b[1]), _ptr(stat_b[2:]))
        last = stat_b if two else stat_a
        _call("dgcnn_edgeconv_select", dev, M, _ptr(ymax), _ptr(ymin), _ptr(amax), _ptr(amin), _ptr(last[2:]),
                                             _ptr(out), COUT, _ptr(ysel), _ptr(arg))
        ctx.meta = meta
        ctx.two = two
        ctx.lin = lin
        ctx.fused = fused
        ctx.xp = xp
        ctx.thin = (uv is not None or uv2 is not None) and xp is None
        ctx.uv2 = uv2                                   # [U | V] of a two-conv block: the backward gathers it instead of re-reading y_a
        ctx.wlin = wlin                                 # [Wa1 ; Wa2 - Wa1]: the data gradient multiplies by it again
        ctx.save_for_backward(rows, idx, wa, wb if two else wa, gamma_a.detach(), gamma_b.detach() if two else gamma_a.detach(),
                              uv if lin else y_a, stat_a, stat_b if two else stat_a, ysel, arg)
        ctx.x_shape = tuple(x.shape)
        ctx.ld = ld
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        rows, idx, wa, wb, gamma_a, gamma_b, y_a, stat_a, stat_b, ysel, arg = ctx.saved_tensors
        B, Cn, N, k = ctx.meta
        two = ctx.two
        M, E = B * N, B * N * k
        dev = dout.device
        partials, dw_partials = _partials(dev)
        f32 = dict(dtype=torch.float32, device=dev)
        if dout.stride(1) != 1:
            dout = dout.contiguous()
        a = _args(rows, ctx.ld, idx, B, Cn, N, k, wa, wb if two else None, None, COUT)

        g = torch.empty(M, COUT, **f32)
        last = stat_b if two else stat_a
        gamma_last = gamma_b if two else gamma_a
        _call("dgcnn_edgeconv_bwd_prep", dev, M, _ptr(dout), dout.stride(0), _ptr(ysel), _ptr(last[2:]), _ptr(g),
                                               _ptr(partials))
        dgl = torch.empty(2, COUT, **f32)      # dgamma, dbeta of the last BN
        coef_l = torch.empty(3, COUT, **f32)
        _call("dgcnn_bn_bwd_reduce", dev, _ptr(partials), COUT, float(E), _ptr(gamma_last), _ptr(last[0]), _ptr(last[1]),
                                           _ptr(dgl[0]), _ptr(dgl[1]), _ptr(coef_l))
        dw_b = dga = None
        fused = ctx.fused if (two and GRAD_NSPLIT == 2) else None
        ctx.fused = None
        # [ r | s ]: scatter target and per-point sums, one GEMM operand; the fused form also needs the in-edge sums of y_a
        # [M,64] and the in-degrees [M]: one zero fill for all of them
        zbuf = torch.zeros(M * 2 * COUT + (M * COUT + M if fused is not None else 0), **f32)
        rs = zbuf[:M * 2 * COUT].view(M, 2 * COUT)
        r, s = rs[:, :COUT], rs[:, COUT:]
        if fused is not None:
            # conv_b backward with the scatter fused in (raw in-edge sums -> r, out-edge sums -> s), then BN_a's backward
            # per POINT: its batch sums from U, V, r, s and the affine map with the forward's graph sums of y_a
            uv = ctx.uv2
            dw_b = torch.empty(COUT, COUT, **f32)
            zin = zbuf[M * 2 * COUT:]                       # in-edge sums of y_a [M,64] | in-degrees [M]
            yin, cnt = zin[:M * COUT].view(M, COUT), zin[M * COUT:]
            _call("dgcnn_edgeconv_bwd_b_fused", dev, C.byref(a), _ptr(y_a), _ptr(stat_a[2:]), _ptr(g), _ptr(arg), _ptr(coef_l),
                  _ptr(r), _ptr(s), 2 * COUT, _ptr(yin), _ptr(cnt), _ptr(dw_b), _ptr(dw_partials))
            _call("dgcnn_edgeconv_bwd_a_point_stats", dev, M, _ptr(uv), uv.stride(0), _ptr(uv[:, COUT:]), uv.stride(0), _ptr(r), _ptr(s),
                  2 * COUT, _ptr(partials))
            dga = torch.empty(2, COUT, **f32)
            coef_a = torch.empty(3, COUT, **f32)
            _call("dgcnn_bn_bwd_reduce", dev, _ptr(partials), COUT, float(E), _ptr(gamma_a), _ptr(stat_a[0]), _ptr(stat_a[1]),
                                               _ptr(dga[0]), _ptr(dga[1]), _ptr(coef_a))
            _call("dgcnn_edgeconv_bwd_a_point_apply", dev, M, k, _ptr(r), _ptr(s), 2 * COUT, _ptr(yin), _ptr(fused), _ptr(cnt),
                  _ptr(coef_a))
            ctx.uv2 = None
        else:
            if two:
                dz_a = torch.empty(E, COUT, **f32)
                dw_b = torch.empty(COUT, COUT, **f32)
                # the tensor-core conv_b backward is a 2-plane (2^-17) kernel: the FP32-accurate setting takes the FP32-FMA one
                ab = a if GRAD_NSPLIT == 2 else _args(rows, ctx.ld, idx, B, Cn, N, k, wa, wb, None, COUT, impl=1)
                _call("dgcnn_edgeconv_bwd_b", dev, C.byref(ab), _ptr(y_a), _ptr(stat_a[2:]), _ptr(g), _ptr(arg), _ptr(coef_l),
                                                    _ptr(dz_a), _ptr(dw_b), _ptr(partials), _ptr(dw_partials))
                dga = torch.empty(2, COUT, **f32)
                coef_a = torch.empty(3, COUT, **f32)
                _call("dgcnn_bn_bwd_reduce", dev, _ptr(partials), COUT, float(E), _ptr(gamma_a), _ptr(stat_a[0]), _ptr(stat_a[1]),
                                                   _ptr(dga[0]), _ptr(dga[1]), _ptr(coef_a))
            else:
                dz_a, coef_a, dga = None, coef_l, dgl
            if ctx.lin or ctx.uv2 is not None:
                uv = y_a if ctx.lin else ctx.uv2            # [U | V] (saved in y_a's place for the single-conv block)
                if DETERMINISTIC:
                    dy_edges = torch.empty(E, COUT, **f32)
                    _call("dgcnn_edgeconv_bwd_a_lin_edges", dev, C.byref(a), _ptr(uv), uv.stride(0), _ptr(uv[:, COUT:]), uv.stride(0),
                          _ptr(dz_a), _ptr(g), _ptr(arg), _ptr(coef_a), _ptr(s), 2 * COUT, _ptr(dy_edges))
                    _scatter_r_deterministic(a, idx, B, N, k, dy_edges, r, dev)
                else:
                    _call("dgcnn_edgeconv_bwd_a_lin", dev, C.byref(a), _ptr(uv), uv.stride(0), _ptr(uv[:, COUT:]), uv.stride(0), _ptr(dz_a),
                          _ptr(g), _ptr(arg), _ptr(coef_a), _ptr(s), _ptr(r), 2 * COUT)
                ctx.uv2 = None
            else:
                _call("dgcnn_edgeconv_bwd_a", dev, C.byref(a), _ptr(y_a), _ptr(dz_a), _ptr(g), _ptr(arg), _ptr(coef_a),
                      _ptr(s), _ptr(r), 2 * COUT)
        # point-level GEMMs on the tensor cores (bf16x2 split: gradient path):
        #   dX  = r Wa1 + s (Wa2 - Wa1)            = [r|s] @ [Wa1 ; Wa2-Wa1]
        #   dWa = [(r - s)^T X | s^T X]            from G = [r|s]^T X  (contraction over the points, split-K)
        X = rows[:, :Cn]
        dx = None
        wcat = ctx.wlin if ctx.wlin is not None else _lin_weight(wa, Cn)     # [128, C] = [Wa1 ; Wa2 - Wa1]
        if getattr(ctx, "thin", False):
            # thin input (K <= 8): one streaming pass gives G = [r|s]^T X and, if asked for, dX = [r|s] [Wa1 ; Wa2-Wa1]
            C4 = (Cn + 3) // 4 * 4
            G = torch.empty(2 * COUT, Cn, **f32)
            dxp = torch.empty(M, C4, **f32) if ctx.needs_input_grad[0] else None
            _call("dgcnn_linear_thin_bwd_f32", dev, _ptr(rows), ctx.ld, _ptr(rs), 2 * COUT, _ptr(wcat), Cn, M, Cn, 2 * COUT, _ptr(dxp), C4,
                  _ptr(G))
            if dxp is not None:
                dx = dxp if dxp.shape == ctx.x_shape else torch.nn.functional.pad(dxp[:, :Cn], (0, ctx.x_shape[1] - Cn))
        else:
            rsp = Planes(rs, GRAD_NSPLIT)                               # split once: K-major for dX, row-contraction for dWa
            if ctx.needs_input_grad[0]:
                dx = gemm_planes(rsp, Planes(wcat.t(), GRAD_NSPLIT), GRAD_NSPLIT)
                if dx.shape != ctx.x_shape:                             # rows padded to a multiple of 4 channels
                    dx = torch.nn.functional.pad(dx, (0, ctx.x_shape[1] - dx.shape[1]))
            xp = ctx.xp if ctx.xp is not None else Planes(X, GRAD_NSPLIT)
            G = gemm_planes(rsp, xp, GRAD_NSPLIT, transposed=True)      # [128, C] = [r|s]^T X
        ctx.xp = None
        dwa = torch.empty(COUT, 2 * Cn, **f32)                          # [(r - s)^T X | s^T X]
        _call("dgcnn_edgeconv_lin_weight_bwd_f32", dev, _ptr(G), Cn, _ptr(dwa))
        # returns: x, idx, meta, w_a, gamma_a, beta_a, w_b, gamma_b, beta_b, bn_a, bn_b
        if two:
            return (dx, None, None, dwa.view(COUT, 2 * Cn, 1, 1), dga[0], dga[1], dw_b.view(COUT, COUT, 1, 1),
                    dgl[0], dgl[1], None, None)
        return dx, None, None, dwa.view(COUT, 2 * Cn, 1, 1), dga[0], dga[1], None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# the drivers' per-point passes on device: training criterion (train.py:180-197, util.py:17-34) and the
# tile-inference epilogue (test.py:123-127)
# ------------------------------------------------------------------------------------------------
UNCLASSIFIED = 31          # test.py:15


class SegLoss(torch.autograd.Function):
    """loss = cal_loss(focus(softmax(logits), mask), focus(labels, mask)) — forward and backward in ONE kernel pass."""

    @staticmethod
    def forward(ctx, logits, labels, mask, smoothing: float, want_pred: bool):
        lib = _lib.load()
        _need_cuda(logits, "seg_loss")
        if logits.dim() != 3:
            raise RuntimeError(f"seg_loss: expected logits [B, classes, N], got {tuple(logits.shape)}")
        B, Cn, N = logits.shape
        dev = logits.device
        labels = labels.to(device=dev, dtype=torch.int64).reshape(B, N).contiguous()
        if mask is not None:
            mask = mask.to(device=dev).reshape(B, N)
            if mask.dtype == torch.bool:
                mask = mask.contiguous().view(torch.uint8)             # same bytes; the kernel tests `!= 0`
            elif mask.dtype != torch.uint8:
                mask = mask.ne(0).to(torch.uint8)
            mask = mask.contiguous()
        z = logits.detach()
        loss = torch.empty((), dtype=torch.float32, device=dev)
        need_grad = logits.requires_grad
        dz = torch.empty_like(z, memory_format=torch.contiguous_format) if need_grad else None
        pred = torch.empty(B, N, dtype=torch.int32, device=dev) if want_pred else None
        nbytes = lib.dgcnn_seg_loss_workspace_bytes(B, N)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        ds = dz.stride() if need_grad else (0, 0, 0)
        _call("dgcnn_seg_loss_f32", dev, _ptr(z), z.stride(0), z.stride(1), z.stride(2), B, Cn, N, _ptr(labels), _ptr(mask),
              float(smoothing), 1.0, _ptr(loss), _ptr(dz), ds[0], ds[1], ds[2], _ptr(pred), _ptr(ws), nbytes)
        ctx.save_for_backward(dz)
        ctx.mark_non_differentiable(*([pred] if want_pred else []))
        return (loss, pred) if want_pred else loss

    @staticmethod
    def backward(ctx, dloss, *_unused):
        (dz,) = ctx.saved_tensors
        return (None if dz is None else dz * dloss), None, None, None, None


def seg_loss(logits: torch.Tensor, labels: torch.Tensor, mask: Optional[torch.Tensor] = None, smoothing: float = 0.2,
             return_pred: bool = False):
    """The reference's training criterion in one device pass.

    logits [B, classes, N] (the model's output), labels [B, N] (any integer/float dtype, as the reference's loaders
    deliver them), mask [B, N] or None.  With a mask: train.py:180-197 (softmax, focus class for unmasked points,
    cal_loss over classes + 1).  Without: the validation form train.py:265-270.  return_pred=True also returns the
    arg-max channel per point (`labels_pred.max(dim=2)[1]`, int32 [B, N])."""
    return SegLoss.apply(logits, labels, mask, smoothing, return_pred)


def seg_predict(logits: torch.Tensor, min_class_confidence: float = 0.8, unclassified: int = UNCLASSIFIED,
                return_conf: bool = False):
    """test.py:119-127 on device: softmax over the classes, (max probability, arg-max) per point, arg-max replaced by
    `unclassified` where the max probability is below min_class_confidence.  logits [B, classes, N] -> int64 [B, N]."""
    _need_cuda(logits, "seg_predict")
    B, Cn, N = logits.shape
    z = logits.detach()
    pred = torch.empty(B, N, dtype=torch.int64, device=z.device)
    conf = torch.empty(B, N, dtype=torch.float32, device=z.device) if return_conf else None
    _call("dgcnn_seg_predict_f32", z.device, _ptr(z), z.stride(0), z.stride(1), z.stride(2), B, Cn, N,
          float(min_class_confidence), int(unclassified), _ptr(pred), _ptr(conf))
    return (pred, conf) if return_conf else pred


if DETERMINISTIC:
    set_deterministic(True)
