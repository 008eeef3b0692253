import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import __graft_entry__ as entry
from test_gpu_tc import _gemm
pkg = entry.load_package()
for (M, N, K) in [(128,128,64),(128,128,128),(256,128,192),(300,200,100),(1000,1024,192),(777,512,1024),(128,5,256),(512,256,4096)]:
    for ns in (3, 2):
        g = torch.Generator().manual_seed(M+N+K)
        A = (torch.randn(M, K, generator=g)*3).cuda(); B = torch.randn(N, K, generator=g).cuda()
        D = _gemm(pkg, A, B, ns)
        ref = A.double() @ B.double().t(); scale = A.double().abs() @ B.double().abs().t()
        e = (D.double()-ref).abs()
        f = ((A@B.t()).double()-ref).abs()
        print(f"M{M} N{N} K{K} ns{ns}: max err/scale {(e/scale).max().item():.2e} rms {((e/scale)**2).mean().sqrt().item():.2e} | rel-to-max|ref| {(e.max()/ref.abs().max()).item():.2e} | fp32 sgemm: {(f/scale).max().item():.2e} rms {((f/scale)**2).mean().sqrt().item():.2e}  bias {((D.double()-ref)/scale).mean().item():.2e}")
