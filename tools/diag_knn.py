"""Diagnostic: tensor-core kNN timing and fallback-row count on random vs real model features (cfg2 shape)."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg._lib.load(); ops = pkg.ops
B, N, k = 8, 7000, 30
def r1024(x): return (x + 1023) // 1024 * 1024
def run_tc(rows, tag):
    nbytes = lib.dgcnn_knn_tc_workspace_bytes(B, N)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    idx = torch.empty(B * N, k, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call():
        pkg._lib.check(lib.dgcnn_knn_tc_f32(C.c_void_p(rows.data_ptr()), rows.stride(0), B, 64, N, k, C.c_void_p(idx.data_ptr()),
                                            C.c_void_p(ws.data_ptr()), nbytes, st))
    call(); torch.cuda.synchronize()
    base = (ws.data_ptr() + 1023) // 1024 * 1024 - ws.data_ptr()
    off_cnt = base + r1024(B * N * 4) + r1024(B * 4)
    cnt = int(ws[off_cnt:off_cnt + 4].view(torch.int32).item())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): call()
    b.record(); torch.cuda.synchronize()
    t_tc = a.elapsed_time(b) / 5
    norms = torch.empty(B * N, device="cuda"); idx2 = torch.empty_like(idx)
    a.record()
    for _ in range(3):
        pkg._lib.check(lib.dgcnn_knn_f32(C.c_void_p(rows.data_ptr()), rows.stride(0), B, 64, N, k, C.c_void_p(idx2.data_ptr()), C.c_void_p(norms.data_ptr()), st))
    b.record(); torch.cuda.synchronize()
    print(f"{tag}: tc {t_tc:.3f} ms, exact {a.elapsed_time(b)/3:.3f} ms, fallback rows {cnt}/{B*N}, equal={torch.equal(idx, idx2)}, "
          f"norm range {norms.min().item():.3g}..{norms.max().item():.3g}")
torch.manual_seed(0)
run_tc(torch.randn(B * N, 64, device="cuda"), "randn")
run_tc(torch.rand(B * N, 64, device="cuda"), "rand")
model = pkg.DGCNN(5, 4, k, cuda=True).train()
x, y = bench.make_batch(0, B)
with torch.no_grad():
    x1 = model(x.cuda(), depth=1).permute(0, 2, 1).reshape(B * N, 64).contiguous()
    x2 = model(x.cuda(), depth=2).permute(0, 2, 1).reshape(B * N, 64).contiguous()
run_tc(x1, "model x1 (train-mode BN)")
run_tc(x2, "model x2 (train-mode BN)")
