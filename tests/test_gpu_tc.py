"""tcgen05 engine checks (GPU): the split-bf16 tensor-core GEMM against an FP64 matmul."""
import ctypes as C

import pytest
import torch

import __graft_entry__ as entry

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    return entry.load_package()


def _gemm(pkg, A, B, nsplit):
    lib = pkg._lib.load()
    M, K = A.shape
    N = B.shape[0]
    D = torch.full((M, N), float("nan"), device=DEV)
    ws_bytes = lib.dgcnn_tc_gemm_workspace_bytes(M, N, K, nsplit)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    rc = lib.dgcnn_tc_gemm_f32(C.c_void_p(A.data_ptr()), A.stride(0), A.stride(1), C.c_void_p(B.data_ptr()), B.stride(0),
                               B.stride(1), C.c_void_p(D.data_ptr()), D.stride(0), M, N, K, nsplit,
                               C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    pkg._lib.check(rc)
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 128), (256, 128, 192), (300, 200, 100), (1000, 1024, 192),
                                   (777, 512, 1024), (128, 5, 256)])
@pytest.mark.parametrize("nsplit", [3, 2])
def test_tc_gemm_matches_fp64(pkg, M, N, K, nsplit):
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 3).to(DEV)
    B = torch.randn(N, K, generator=g).to(DEV)
    D = _gemm(pkg, A, B, nsplit)
    ref = A.double() @ B.double().t()
    scale = (A.double().abs() @ B.double().abs().t())            # sum |a||b| : the natural error scale
    err = ((D.double() - ref).abs() / scale).max().item()
    assert torch.isfinite(D).all()
    assert err < (2e-7 if nsplit == 3 else 3e-5), err
    if nsplit == 3:                                                # as good as an FP32 FMA chain
        fp32 = ((A @ B.t()).double() - ref).abs().div(scale).max().item()
        assert err < max(4 * fp32, 2e-7), (err, fp32)


def test_tc_gemm_transposed_operands(pkg):
    g = torch.Generator().manual_seed(5)
    At = torch.randn(192, 300, generator=g).to(DEV)      # A = At^T  (M=300, K=192)
    Bt = torch.randn(192, 256, generator=g).to(DEV)      # B = Bt^T  (N=256)
    D = _gemm(pkg, At.t(), Bt.t(), 3)
    ref = (At.double().t() @ Bt.double())
    assert ((D.double() - ref).abs().max() / ref.abs().max()).item() < 1e-6
