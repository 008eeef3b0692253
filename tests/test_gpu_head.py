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
