"""ops.set_deterministic(True): every scatter of the backward becomes a fixed-order gather (csrc/graph_transpose.cu), so two
runs of the same step give bit-identical gradients; the default (float reductions) agrees with it to rounding."""
import pytest
import torch

import __graft_entry__ as entry
from golden import common

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _step(pkg, x, y, mask, w):
    torch.manual_seed(0)
    net = pkg.DGCNN(5, 4, 20, dropout=0.5, cuda=True)
    net.load_state_dict(w, strict=False)
    net.train()
    torch.manual_seed(123)                       # dropout mask
    loss = pkg.ops.seg_loss(net(x), y, mask)
    loss.backward()
    return loss.detach().clone(), [p.grad.clone() for p in net.parameters()]


def test_deterministic_mode_is_bit_reproducible_and_matches_the_default():
    pkg = entry.load_package()
    ops = pkg.ops
    B, N = 3, 1500
    x = common.points(301, B, 4, N, "real").to(DEV)
    g = torch.Generator().manual_seed(7)
    y = torch.randint(0, 5, (B, N), generator=g).to(DEV)
    mask = (torch.rand(B, N, generator=g) < 0.7).to(DEV)
    w = common.weights(302, 5, 4, 1024)
    saved = (ops.FUSED_BWD, ops.ATOMIC_WGRAD, ops.NARROW_LINEAR, ops.THIN_LINEAR, ops.DETERMINISTIC)
    try:
        ops.set_deterministic(True)
        l1, g1 = _step(pkg, x, y, mask, w)
        l2, g2 = _step(pkg, x, y, mask, w)
        assert torch.equal(l1, l2)
        for a, b in zip(g1, g2):
            assert torch.equal(a, b)
        # the default scatters (float reductions), with the same forward arithmetic: the narrow / thin streaming layers stay off
        # (their FP32-FMA forward differs from the bf16x3 GEMM in the last bit, which re-routes max-over-k ties)
        ops.set_deterministic(False)
        ops.NARROW_LINEAR = ops.THIN_LINEAR = False
        l3, g3 = _step(pkg, x, y, mask, w)
    finally:
        ops.FUSED_BWD, ops.ATOMIC_WGRAD, ops.NARROW_LINEAR, ops.THIN_LINEAR, ops.DETERMINISTIC = saved
    assert abs(float(l1) - float(l3)) <= 1e-5 * abs(float(l1))
    for a, b in zip(g1, g3):
        assert ((a - b).norm() / (a.norm() + 1e-30)).item() < 5e-4


@pytest.mark.parametrize("B,N", [(2, 1500), (1, 1000)])
def test_default_step_repeats_within_rounding(B, N):
    """Stress for rare pipeline races on the DEFAULT path (float reductions, neighbour sets in any order): 30 repetitions of
    the same step with a poisoned allocator in between must agree with the first one to rounding — a tile computed from a
    buffer that was overwritten too early shows up as a gross error in the loss or in some gradient."""
    pkg = entry.load_package()
    x = common.points(311 + N, B, 4, N, "real").to(DEV)
    g = torch.Generator().manual_seed(8)
    y = torch.randint(0, 5, (B, N), generator=g).to(DEV)
    mask = (torch.rand(B, N, generator=g) < 0.7).to(DEV)
    w = common.weights(312, 5, 4, 1024)
    l0, g0 = _step(pkg, x, y, mask, w)
    for it in range(30):
        junk = torch.full((1 << 23,), float("nan"), device=DEV)
        del junk
        l1, g1 = _step(pkg, x, y, mask, w)
        assert abs(float(l1) - float(l0)) <= 2e-5 * abs(float(l0)), (it, float(l0), float(l1))
        for i, (a, b) in enumerate(zip(g0, g1)):
            assert ((a - b).norm() / (a.norm() + 1e-30)).item() < 2e-3, (it, i)


@pytest.mark.parametrize("B,N,k", [(2, 700, 20), (1, 300, 7), (3, 257, 48)])
def test_graph_transpose_lists_every_in_edge_in_order(B, N, k):
    pkg = entry.load_package()
    ops = pkg.ops
    import ctypes as C
    lib = __import__(entry.PKG_NAME + "._lib", fromlist=["x"]).load()
    g = torch.Generator().manual_seed(B * N + k)
    idx = torch.stack([torch.stack([torch.randperm(N, generator=g)[:k] for _ in range(N)]) for _ in range(B)]).to(torch.int32).to(DEV)
    M = B * N
    in_start = torch.empty(M + 1, dtype=torch.int32, device=DEV)
    in_edge = torch.empty(M * k, dtype=torch.int32, device=DEV)
    ws = torch.empty(lib.dgcnn_graph_transpose_workspace_bytes(B, N), dtype=torch.uint8, device=DEV)
    ops._call("dgcnn_graph_transpose", idx.device, ops._ptr(idx), B, N, k, ops._ptr(in_start), ops._ptr(in_edge), ops._ptr(ws), ws.numel())
    torch.cuda.synchronize()
    tgt = (idx.view(B, N * k).long() + (torch.arange(B, device=DEV) * N)[:, None]).view(-1)      # global target of every edge
    order = torch.sort(tgt, stable=True)[1]                                                      # edges by target, ascending edge id
    assert torch.equal(in_edge.long(), order)
    counts = torch.bincount(tgt, minlength=M)
    assert torch.equal(in_start.long(), torch.cat((torch.zeros(1, dtype=torch.long, device=DEV), counts.cumsum(0))))
