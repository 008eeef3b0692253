"""Time the tcgen05 split-bf16 GEMM on the head's shapes vs torch FP32 matmul (SIMT sgemm)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg._lib.load()
def run(M, N, K, ns, ta=False, tb=False):
    A = torch.randn(K, M, device="cuda").t() if ta else torch.randn(M, K, device="cuda")
    Bm = torch.randn(K, N, device="cuda").t() if tb else torch.randn(N, K, device="cuda")
    D = torch.empty(M, N, device="cuda")
    nb = lib.dgcnn_tc_gemm_workspace_bytes(M, N, K, ns); ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call():
        pkg._lib.check(lib.dgcnn_tc_gemm_f32(C.c_void_p(A.data_ptr()), A.stride(0), A.stride(1), C.c_void_p(Bm.data_ptr()), Bm.stride(0), Bm.stride(1),
                                             C.c_void_p(D.data_ptr()), D.stride(0), M, N, K, ns, C.c_void_p(ws.data_ptr()), nb, st))
    call(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): call()
    b.record(); torch.cuda.synchronize(); t = a.elapsed_time(b) / 5
    a.record()
    for _ in range(5): R = A @ Bm.t()
    b.record(); torch.cuda.synchronize(); t2 = a.elapsed_time(b) / 5
    fl = 2.0 * M * N * K
    print(f"M{M} N{N} K{K} ns{ns} ta{int(ta)} tb{int(tb)}: tc {t:.3f} ms ({fl/t/1e9:.1f} TFLOP/s alg) | torch fp32 {t2:.3f} ms ({fl/t2/1e9:.1f}) | maxerr {(D-R).abs().max().item():.2e}")
M = 56000
for ns in (2, 3):
    run(M, 1024, 192, ns); run(M, 512, 192, ns); run(M, 256, 512, ns)
    run(M, 192, 1024, ns, tb=True)          # dX of conv6
    run(1024, 192, M, ns, ta=True, tb=True) # dW of conv6 (contraction over points)
