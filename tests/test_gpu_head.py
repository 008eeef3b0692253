"""Shared-MLP head kernels (GPU): fused BatchNorm1d(+block bias)+LeakyReLU rows op against an FP64 torch composition
of the reference's layers (model.py:131-139,179-184)."""
import pytest
import torch
import torch.nn.functional as F

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    return entry.load_package()


def _reference(y, bias, N, gamma, beta, rm, rv, training, eps=1e-5, momentum=0.1):
    v = y if bias is None else y + bias.repeat_interleave(N, dim=0)
    z = F.batch_norm(v, rm, rv, gamma, beta, training, momentum, eps)
    return F.leaky_relu(z, 0.2)


@pytest.mark.parametrize("B,N,C,with_bias", [(3, 700, 512, True), (2, 1001, 256, False), (2, 37, 64, True), (4, 130, 1024, False)])
def test_rows_bn_lrelu_training_matches_fp64(pkg, B, N, C, with_bias):
    g = torch.Generator().manual_seed(B * N + C)
    M = B * N
    y = (torch.randn(M, C, generator=g) * 2 + 0.3).to(DEV)
    bias = torch.randn(B, C, generator=g).to(DEV) if with_bias else None
    bn = torch.nn.BatchNorm1d(C).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(C, generator=g))          # both signs of gamma
        bn.bias.copy_(torch.randn(C, generator=g))
    ct = torch.randn(M, C, generator=g).to(DEV)

    y1 = y.clone().requires_grad_(True)
    b1 = None if bias is None else bias.clone().requires_grad_(True)
    h = pkg.ops.bn_lrelu_rows(y1, bn, True, N, bias=b1)
    (h * ct).sum().backward()

    y2 = y.double().requires_grad_(True)
    b2 = None if bias is None else bias.double().requires_grad_(True)
    g2, be2 = bn.weight.detach().double().requires_grad_(True), bn.bias.detach().double().requires_grad_(True)
    rm, rv = torch.zeros(C, dtype=torch.float64, device=DEV), torch.ones(C, dtype=torch.float64, device=DEV)
    href = _reference(y2, b2, N, g2, be2, rm, rv, True)
    (href * ct.double()).sum().backward()

    def rel(a, b):
        return ((a.double() - b).norm() / (b.norm() + 1e-30)).item()

    assert rel(h, href) < 2e-6
    assert rel(y1.grad, y2.grad) < 2e-5
    assert rel(bn.weight.grad, g2.grad) < 2e-5 and rel(bn.bias.grad, be2.grad) < 2e-5
    if with_bias:
        # the block-bias gradient is a sum of N row gradients that nearly cancel (BN removes shifts): scale by the rows
        assert ((b1.grad.double() - b2.grad).norm() / (y2.grad.norm() * N ** 0.5)).item() < 2e-6
    # running statistics follow nn.BatchNorm1d (momentum 0.1, unbiased variance)
    torch.testing.assert_close(bn.running_mean.double(), rm, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(bn.running_var.double(), rv, rtol=1e-5, atol=1e-6)
    assert int(bn.num_batches_tracked) == 1


def test_rows_bn_lrelu_inference_uses_running_stats(pkg):
    g = torch.Generator().manual_seed(7)
    B, N, C = 2, 333, 256
    y = torch.randn(B * N, C, generator=g).to(DEV)
    bias = torch.randn(B, C, generator=g).to(DEV)
    bn = torch.nn.BatchNorm1d(C).to(DEV).eval()
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(C, generator=g))
        bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
        bn.weight.copy_(torch.randn(C, generator=g))
        bn.bias.copy_(torch.randn(C, generator=g))
    with torch.no_grad():
        h = pkg.ops.bn_lrelu_rows(y, bn, False, N, bias=bias)
    ref = _reference(y.double(), bias.double(), N, bn.weight.double(), bn.bias.double(), bn.running_mean.double(),
                     bn.running_var.double(), False)
    assert ((h.double() - ref).abs().max() / ref.abs().max()).item() < 2e-6
    assert int(bn.num_batches_tracked) == 0


def test_rows_bn_rejects_unsupported_width(pkg):
    lib = pkg._lib.load()
    assert lib.dgcnn_rows_bn_supported(512) == 1 and lib.dgcnn_rows_bn_supported(100) == 0


@pytest.mark.parametrize("B,N,C", [(8, 7000, 512), (3, 101, 64), (1, 5000, 256), (1, 56000, 192), (2, 333, 20)])
def test_rows_colsum_matches_fp64(pkg, B, N, C):
    g = torch.Generator().manual_seed(B + N + C)
    x = torch.randn(B * N, C, generator=g).to(DEV)
    out = pkg.ops.rows_colsum(x, N)
    ref = x.double().view(B, N, C).sum(1)
    assert out.shape == (B, C)
    assert ((out.double() - ref).abs().max() / ref.abs().max()).item() < 1e-6
    assert torch.equal(out, pkg.ops.rows_colsum(x, N))          # deterministic


@pytest.mark.parametrize("B,N", [(3, 700), (2, 1001), (8, 7000)])
def test_head_mlp_node_matches_fp64_and_the_unfused_chain(pkg, B, N):
    """ops.head_mlp (conv7 + bn7 + LReLU + conv8 + bn8 + LReLU as one node, activations between BN and GEMM only as operand
    planes) against an FP64 composition of the reference's layers (model.py:176-184) and against the op-by-op chain."""
    ops = pkg.ops
    g = torch.Generator().manual_seed(B * 31 + N)
    M = B * N
    feat0 = torch.randn(M, 192, generator=g).to(DEV)
    bias0 = torch.randn(B, 512, generator=g).to(DEV)
    w7_0 = (torch.randn(512, 192, generator=g) * 0.08).to(DEV)
    w8_0 = (torch.randn(256, 512, generator=g) * 0.05).to(DEV)
    up = torch.randn(M, 256, generator=g).to(DEV)

    def run(mode):
        dt = torch.float64 if mode == "fp64" else torch.float32
        feat, bias, w7, w8 = (t.clone().to(dt).requires_grad_(True) for t in (feat0, bias0, w7_0, w8_0))
        bn7, bn8 = torch.nn.BatchNorm1d(512).to(DEV).to(dt), torch.nn.BatchNorm1d(256).to(DEV).to(dt)
        with torch.no_grad():
            bn7.weight.copy_(torch.linspace(-1.0, 1.5, 512)); bn7.bias.copy_(torch.linspace(-0.3, 0.3, 512))
            bn8.weight.copy_(torch.linspace(1.2, -0.8, 256)); bn8.bias.copy_(torch.linspace(0.2, -0.2, 256))
        if mode == "fused":
            h8 = ops.head_mlp(feat, bias, w7, bn7, w8, bn8, N)
        elif mode == "chain":
            h7 = ops.bn_lrelu_rows(ops.linear(feat, w7), bn7, True, N, bias=bias)
            h8 = ops.bn_lrelu_rows(ops.linear(h7, w8), bn8, True, N)
        else:
            v7 = feat @ w7.t() + bias.repeat_interleave(N, dim=0)
            h7 = F.leaky_relu(bn7(v7), 0.2)
            h8 = F.leaky_relu(bn8(h7 @ w8.t()), 0.2)
        (h8 * up.to(dt)).sum().backward()
        grads = [feat.grad, bias.grad, w7.grad, w8.grad, bn7.weight.grad, bn7.bias.grad, bn8.weight.grad, bn8.bias.grad]
        state = [bn7.running_mean, bn7.running_var, bn8.running_mean, bn8.running_var, bn7.num_batches_tracked, bn8.num_batches_tracked]
        return h8.detach(), grads, state

    hf, gf, sf = run("fused")
    hc, gc, sc = run("chain")
    hr, gr, sr = run("fp64")
    assert ((hf.double() - hr).abs().max() / hr.abs().max()).item() < 2e-5
    assert ((hf - hc).abs().max() / hc.abs().max()).item() < 2e-6           # same FP32-accurate forward arithmetic
    for a, b in zip(sf, sr):
        torch.testing.assert_close(a.double(), b.double(), rtol=1e-5, atol=1e-6)
    names = ["feat", "bias7", "w7", "w8", "gamma7", "beta7", "gamma8", "beta8"]
    for n, a, c, r in zip(names, gf, gc, gr):
        ef = ((a.double() - r).norm() / (r.norm() + 1e-30)).item()
        ec = ((c.double() - r).norm() / (r.norm() + 1e-30)).item()
        assert ef < max(2e-4, 3 * ec), (n, ef, ec)                           # 2-plane gradient GEMMs: 2^-17 per product


@pytest.mark.parametrize("M,K,NO", [(56000, 256, 5), (1003, 256, 8), (777, 128, 1), (4100, 512, 3), (300, 64, 5), (50, 20, 2)])
def test_linear_narrow_matches_fp64(pkg, M, K, NO):
    """ops.linear with <= 8 outputs (the classifier conv9, model.py:141): streaming FP32-FMA kernels, forward and backward."""
    ops = pkg.ops
    g = torch.Generator().manual_seed(M + K + NO)
    x0 = torch.randn(M, K, generator=g).to(DEV)
    w0 = (torch.randn(NO, K, generator=g) * 0.1).to(DEV)
    up = torch.randn(M, NO, generator=g).to(DEV)
    x, w = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    y = ops.linear(x, w)
    assert y.grad_fn is not None and "LinearNarrow" in type(y.grad_fn).__name__
    (y * up).sum().backward()
    xd, wd = x0.double().requires_grad_(True), w0.double().requires_grad_(True)
    yd = xd @ wd.t()
    (yd * up.double()).sum().backward()
    scale = (x0.double().abs() @ w0.double().abs().t())
    assert ((y.double() - yd).abs() / scale).max().item() < 2e-6
    assert ((x.grad.double() - xd.grad).abs().max() / xd.grad.abs().max()).item() < 2e-6
    assert ((w.grad.double() - wd.grad).norm() / wd.grad.norm()).item() < 2e-6
    # a row-strided input (a column slice) goes through the same kernel
    xs = torch.randn(M, K + 8, generator=g).to(DEV)
    ys = ops.linear(xs[:, :K], w0)
    assert ((ys.double() - xs[:, :K].double() @ w0.double().t()).abs() / (xs[:, :K].double().abs() @ w0.double().abs().t())).max().item() < 2e-6
