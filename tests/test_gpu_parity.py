"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C-ABI library via the
drop-in `model.py`; the CPU oracle (oracle/dgcnn_oracle.py) and the committed golden vectors are the checkers.

Protocol:
  P1 kNN           oracle layer input -> neighbour SETS equal; every differing row must be an exact/near tie
  P2 EdgeConv      oracle x + oracle idx -> block output within 1e-5 * max|.|
  P3 model forced  all three graphs from the oracle -> logits within 1e-3 * max|logit| (tolerance of north_star)
  P4 free-running  neighbour-set overlap >= 90 % per layer (reported)
  P5 gradients     forced idx, dropout off -> all 25 gradients within 1e-3 rel-L2, BN running stats equal
"""
import os

import numpy as np
import pytest
import torch

import common
import __graft_entry__ as entry
from oracle import dgcnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    return entry.load_package()


def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _relmax(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _rel_l2(a, b):
    return (a - b).double().norm().item() / max(b.double().norm().item(), 1e-30)


# ---------------------------------------------------------------------------------------------- P1
@pytest.mark.parametrize("case", common.KNN_CASES, ids=[c[0] for c in common.KNN_CASES])
def test_knn_golden(pkg, golden_dir, case):
    name, B, C, N, k, kind, seed = case
    g = _golden(golden_dir, "knn.npz")
    x = common.points(seed, B, C, N, kind)
    ref = torch.from_numpy(g[f"{name}/idx"].astype(np.int64))
    got = pkg.knn(x.to(DEV), k)
    assert got.dtype == torch.int64 and tuple(got.shape) == (B, N, k)
    rep = O.compare_neighbour_sets(x, ref, got.cpu(), tol_ulps=4.0 if C > 16 else 0.0)
    assert rep["hard"] == 0, rep
    if C <= 16 and kind != "dup":
        assert rep["identical"] == rep["rows"], rep          # bit-exact scores => identical sets
    # non-contiguous caller view (train.py:176) gives the same table
    xv = x.permute(0, 2, 1).contiguous().permute(0, 2, 1).to(DEV)
    assert torch.equal(pkg.knn(xv, k), got)


@pytest.mark.parametrize("B,C,N,k,kind", [
    (2, 4, 2000, 30, "real"), (1, 4, 3000, 20, "dup"), (2, 64, 1500, 30, "randn"), (1, 64, 1000, 40, "randn"),
    (1, 128, 700, 20, "randn"), (1, 4, 600, 100, "unit"), (1, 3, 500, 16, "unit"), (1, 9, 400, 33, "unit"),
    (3, 4, 129, 64, "unit"), (1, 4, 64, 64, "unit"),
])
def test_knn_vs_oracle(pkg, B, C, N, k, kind):
    x = common.points(1000 + N + C, B, C, N, kind)
    ref = O.knn(x, k)
    got = pkg.knn(x.to(DEV), k).cpu()
    assert int(got.min()) >= 0 and int(got.max()) < N
    assert (torch.sort(got, -1)[0].diff(dim=-1) > 0).all(), "duplicate neighbour in a row"
    rep = O.compare_neighbour_sets(x, ref, got, tol_ulps=4.0 if C > 4 else 0.0)
    assert rep["hard"] == 0, rep
    assert rep["identical"] >= 0.99 * rep["rows"] or kind == "dup", rep


def test_knn_best_first_order_and_self(pkg):
    x = common.points(77, 2, 4, 800, "real")
    got = pkg.knn(x.to(DEV), 20).cpu()
    pd = O.pairwise_score(x)
    vals = torch.gather(pd, 2, got)
    assert (vals.diff(dim=-1) <= 0).all()                     # sorted like topk
    assert (got[:, :, 0] == torch.arange(800)).all()          # distinct points: self is the nearest


def test_knn_errors(pkg):
    x = torch.rand(1, 4, 10, device=DEV)
    with pytest.raises(RuntimeError, match="out of range"):
        pkg.knn(x, 11)
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.knn(torch.rand(1, 4, 10), 3)


def test_knn_large_properties(pkg):
    """cfg2 / cfg4-sized input graphs: size-independent properties + sampled rows against the oracle (bit-exact for C=4).
    The feature-space graphs (C=64) at these sizes are compared row by row in test_gpu_baseline_shapes.py."""
    for (B, C, N, k) in ((8, 4, 7000, 30), (1, 4, 32768, 30), (1, 4, 65536, 20)):
        x = common.points(5, B, C, N, "real")
        got = pkg.knn(x.to(DEV), k).cpu()
        assert int(got.min()) >= 0 and int(got.max()) < N
        assert (torch.sort(got, -1)[0].diff(dim=-1) > 0).all()
        assert (got[:, :, 0] == torch.arange(N)).all()
        rows = torch.randint(0, N, (64,), generator=torch.Generator().manual_seed(1)).tolist()
        pd = O.score_rows(x, B - 1, rows)
        ref = pd.topk(k, dim=-1)[1]
        a, b = torch.sort(ref, -1)[0], torch.sort(got[B - 1, rows], -1)[0]
        assert torch.equal(a, b)


def test_knn_permutation_equivariance(pkg):
    x = common.points(9, 1, 4, 1024, "unit")
    perm = torch.randperm(1024, generator=torch.Generator().manual_seed(3))
    a = pkg.knn(x.to(DEV), 16).cpu()
    b = pkg.knn(x[:, :, perm].contiguous().to(DEV), 16).cpu()
    # neighbours of permuted point i (= original perm[i]) are perm^-1 of the original neighbours
    mapped = perm[b[0]]
    assert torch.equal(torch.sort(mapped, -1)[0], torch.sort(a[0, perm], -1)[0])


# ---------------------------------------------------------------------------------------------- a2
@pytest.mark.parametrize("case", common.GRAPH_CASES, ids=[c[0] for c in common.GRAPH_CASES])
def test_graph_feature_golden(pkg, golden_dir, case):
    name, B, C, N, k, kind, seed = case
    g = _golden(golden_dir, "graph_feature.npz")
    x = common.points(seed, B, C, N, kind)
    f = pkg.get_graph_feature(x.to(DEV), k=k, use_cuda=True)
    assert tuple(f.shape) == (B, 2 * C, N, k) and tuple(f.stride()) == tuple(g[f"{name}/strides"])
    ref = torch.from_numpy(g[f"{name}/feature"])
    # neighbour ORDER may differ on ties only; compare as sets of edge vectors per point via sorting on all channels
    got = f.cpu().contiguous()
    if torch.equal(got, ref):
        return
    a = got.permute(0, 2, 3, 1).reshape(B * N, k, 2 * C)
    b = ref.permute(0, 2, 3, 1).reshape(B * N, k, 2 * C)
    assert torch.equal(torch.sort(a, dim=1)[0], torch.sort(b, dim=1)[0])


# ---------------------------------------------------------------------------------------------- P2 / P3
def _make(pkg, case, train):
    name, B, Fd, N, k, classes, emb, kind, seed, _ = case
    x = common.points(seed, B, Fd, N, kind)
    w = common.weights(seed + 100, classes, Fd, emb)
    net = pkg.DGCNN(classes, Fd, k, dropout=0.0 if train else 0.5, emb_dims=emb, cuda=True)
    missing = net.load_state_dict(w, strict=False)
    assert not missing.unexpected_keys and all(".1." in m for m in missing.missing_keys)
    net.train(train)
    return net, x, w


@pytest.mark.parametrize("case", [c for c in common.MODEL_CASES if not c[-1]], ids=lambda c: c[0])
def test_model_eval_golden(pkg, golden_dir, case):
    name, B, Fd, N, k, classes, emb, kind, seed, _ = case
    g = _golden(golden_dir, "model.npz")
    net, x, w = _make(pkg, case, False)
    forced = [torch.from_numpy(g[f"{name}/idx{i}"].astype(np.int64)) for i in (1, 2, 3)]
    xin = x.permute(0, 2, 1).contiguous().permute(0, 2, 1).to(DEV)
    logits = net(xin, forced_idx=forced)
    assert tuple(logits.shape) == (B, classes, N) and logits.is_contiguous()
    ref = torch.from_numpy(g[f"{name}/logits"])
    assert _relmax(logits.cpu(), ref) < 1e-3                  # north_star tolerance
    assert _relmax(logits.cpu(), ref) < 5e-5                  # what an FP32-accurate path achieves
    for d in range(1, 10):
        y = net(xin, depth=d, forced_idx=forced).cpu()
        assert tuple(y.shape) == tuple(g[f"{name}/depth{d}.shape"]), d
        flat = y.reshape(-1)
        sub = flat if flat.numel() <= 20000 else flat[::37]
        refd = torch.from_numpy(g[f"{name}/depth{d}"])
        assert (sub - refd).abs().max().item() <= 1e-5 * max(refd.abs().max().item(), 1.0) * (1 if d <= 4 else 20), d
    # free-running graph: layer-1 table must equal the reference's (bit-exact scores)
    net(xin)
    rep = O.compare_neighbour_sets(x, forced[0], net.last_idx[0].view(B, N, k).cpu().long())
    assert rep["hard"] == 0


def test_edgeconv_blocks_vs_oracle(pkg):
    """P2 on a mid-sized block, incl. a ragged last tile (M not a multiple of the points-per-tile)."""
    B, Fd, N, k, classes = 2, 4, 1003, 30, 5
    x = common.points(41, B, Fd, N, "real")
    w = common.weights(42, classes, Fd, 1024)
    ora = O.OracleDGCNN({kk: v.clone() for kk, v in w.items()}, k, training=False)
    ora(x, depth=3)
    forced = [ora.trace[f"idx{i}"] for i in (1, 2, 3)]
    net = pkg.DGCNN(classes, Fd, k, cuda=True)
    net.load_state_dict(w, strict=False)
    net.eval()
    for d in (1, 2, 3):
        y = net(x.to(DEV), depth=d, forced_idx=forced).cpu()
        ref = ora.trace[f"x{d}"]
        assert _relmax(y, ref) < 1e-5 * d, (d, _relmax(y, ref))


# ---------------------------------------------------------------------------------------------- P5
def _oracle_train(w, x, k, cot, forced, dtype):
    po = {kk: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for kk, v in w.items()}
    for key, v in po.items():
        if v.is_floating_point() and "running" not in key:
            v.requires_grad_(True)
    ora = O.OracleDGCNN(po, k, dropout=0.0, training=True)
    out = ora(x.to(dtype), forced_idx=forced)
    if cot is None:
        return out.detach(), po, ora
    (out * cot.to(dtype)).sum().backward()
    return out.detach(), po, ora


def _train_compare(pkg, B, Fd, N, k, classes, kind, seed):
    """Ours vs the oracle in FP32 (the reference's arithmetic) and in FP64 (ground truth).

    The reference's own FP32 autograd is only reproducible to a few 1e-3 rel-L2: one LeakyReLU pre-activation
    or max-pool candidate that rounds to the other side of a tie re-routes a gradient (measured on B200:
    CPU-FP32 oracle vs FP64 oracle differs by 1e-3..5e-3 on bn/conv grads while logits agree to 5e-6).
    Parity is therefore asserted against the FP64 oracle with the FP32 oracle's own error as the envelope."""
    x = common.points(seed, B, Fd, N, kind)
    w = common.weights(seed + 1, classes, Fd, 1024)
    ref, _, ora = _oracle_train(w, x, k, None, None, torch.float32)
    forced = [ora.trace[f"idx{i}"] for i in (1, 2, 3)]
    cot = common.cotangent(seed + 2, tuple(ref.shape))
    o32, p32, _ = _oracle_train(w, x, k, cot, forced, torch.float32)
    o64, p64, _ = _oracle_train(w, x, k, cot, forced, torch.float64)

    net = pkg.DGCNN(classes, Fd, k, dropout=0.0, cuda=True)
    net.load_state_dict(w, strict=False)
    net.train()
    out = net(x.to(DEV), forced_idx=forced)
    (out * cot.to(DEV)).sum().backward()
    res = {"logits": (_relmax(out.detach().cpu(), o64.float()), _relmax(o32, o64.float()))}
    for name, prm in net.named_parameters():
        res[name] = (_rel_l2(prm.grad.cpu(), p64[name].grad), _rel_l2(p32[name].grad, p64[name].grad))
    for i in range(1, 9):
        bn = getattr(net, f"bn{i}")
        for short, stat in (("rm", "running_mean"), ("rv", "running_var")):
            res[f"bn{i}.{short}"] = (_relmax(getattr(bn, stat).cpu(), p64[f"bn{i}.{stat}"].float()),
                                     _relmax(p32[f"bn{i}.{stat}"], p64[f"bn{i}.{stat}"].float()))
        assert int(bn.num_batches_tracked) == 1
    return res


@pytest.mark.parametrize("B,N,k,kind", [(2, 80, 6, "real"), (2, 515, 30, "real"), (3, 300, 20, "unit")])
def test_model_train_vs_oracle(pkg, B, N, k, kind):
    res = _train_compare(pkg, B, 4, N, k, 5, kind, 500 + N)
    assert res["logits"][0] < 1e-3, res                      # north_star tolerance ...
    assert res["logits"][0] < 5e-5, res                      # ... and what FP32-accurate kernels achieve
    # floor 4e-3: ONE LeakyReLU sign flip / max tie between two FP32 evaluations moves a bias gradient by ~1e-3
    # (measured: the CPU-FP32 oracle itself sits 1e-3..5e-3 from its FP64 run on most tensors)
    for name, (ours, ref32) in res.items():
        assert ours <= max(4e-3, 3.0 * ref32), (name, ours, ref32)


def test_model_train_golden(pkg, golden_dir):
    case = [c for c in common.MODEL_CASES if c[-1]][0]
    name, B, Fd, N, k, classes, emb, kind, seed, _ = case
    g = _golden(golden_dir, "model.npz")
    net, x, w = _make(pkg, case, True)
    forced = [torch.from_numpy(g[f"{name}/idx{i}"].astype(np.int64)) for i in (1, 2, 3)]
    out = net(x.to(DEV), forced_idx=forced)
    ref = torch.from_numpy(g[f"{name}/logits"])
    assert _relmax(out.detach().cpu(), ref) < 1e-3
    cot = common.cotangent(seed + 200, tuple(ref.shape)).to(DEV)
    (out * cot).sum().backward()
    for pname, prm in net.named_parameters():
        gr = prm.grad.detach().reshape(-1).cpu()
        sub = gr if gr.numel() <= 20000 else gr[::37]
        refg = torch.from_numpy(g[f"grad/{name}/{pname}"])
        # FP32-vs-FP32: tie flips in LeakyReLU / max re-route gradients at the few-1e-3 level (see _train_compare)
        assert _rel_l2(sub, refg) < 1e-2, pname
        assert abs(float(gr.double().norm()) - float(g[f"gradnorm/{name}/{pname}"])) <= 1e-2 * float(g[f"gradnorm/{name}/{pname}"]) + 1e-12
    for i in range(1, 9):
        bn = getattr(net, f"bn{i}")
        np.testing.assert_allclose(bn.running_mean.cpu().numpy(), g[f"{name}/bn{i}.running_mean"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(bn.running_var.cpu().numpy(), g[f"{name}/bn{i}.running_var"], rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------------------------------------------- P4
def test_free_running_overlap(pkg):
    B, Fd, N, k, classes = 2, 4, 2000, 30, 5
    x = common.points(61, B, Fd, N, "real")
    w = common.weights(62, classes, Fd, 1024)
    ora = O.OracleDGCNN({kk: v.clone() for kk, v in w.items()}, k, training=False)
    ref = ora(x)
    net = pkg.DGCNN(classes, Fd, k, cuda=True)
    net.load_state_dict(w, strict=False)
    net.eval()
    out = net(x.to(DEV)).cpu()
    for li in (1, 2, 3):
        a = torch.sort(ora.trace[f"idx{li}"], -1)[0]
        b = torch.sort(net.last_idx[li - 1].view(B, N, k).cpu().long(), -1)[0]
        same = (a == b).all(-1).float().mean().item()
        assert same >= 0.90, (li, same)
    assert _rel_l2(out, ref) < 5e-2


# ---------------------------------------------------------------------------------------------- properties
def test_batch_independence_eval(pkg):
    B, Fd, N, k = 3, 4, 700, 20
    x = common.points(71, B, Fd, N, "real").to(DEV)
    net = pkg.DGCNN(5, Fd, k, cuda=True)
    net.eval()
    full, full3 = net(x), net(x, depth=3)
    for b in range(B):
        # our kernels are per-block deterministic: EdgeConv output is bit-identical for any batch composition
        assert torch.equal(net(x[b:b + 1], depth=3), full3[b:b + 1])
        # the head's library GEMMs may pick another tiling for another M: equal to FP32 rounding only
        torch.testing.assert_close(net(x[b:b + 1]), full[b:b + 1], rtol=1e-4, atol=1e-5)


def test_state_dict_roundtrip(pkg):
    a = pkg.DGCNN(5, 4, 10, cuda=True)
    b = pkg.DGCNN(5, 4, 10, cuda=True)
    b.load_state_dict(a.state_dict())              # strict, 89 keys incl. aliases
    a.eval(), b.eval()
    x = common.points(81, 1, 4, 300, "unit").to(DEV)
    assert torch.equal(a(x), b(x))
    assert len(a.state_dict()) == 89
