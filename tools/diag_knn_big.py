import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import __graft_entry__ as entry
from bench_infer import blocks, timeit
pkg = entry.load_package(); lib = pkg._lib.load(); ops = pkg.ops
def r1024(x): return (x + 1023) // 1024 * 1024
def run(rows, B, N, k, tag):
    nbytes = lib.dgcnn_knn_tc_workspace_bytes(B, N)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    idx = torch.empty(B * N, k, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: pkg._lib.check(lib.dgcnn_knn_tc_f32(C.c_void_p(rows.data_ptr()), rows.stride(0), B, 64, N, k, C.c_void_p(idx.data_ptr()), C.c_void_p(ws.data_ptr()), nbytes, st))
    ms = timeit(call, n=2, warm=1)
    base = (ws.data_ptr() + 1023) // 1024 * 1024 - ws.data_ptr()
    off_cnt = base + r1024(B * N * 4) + r1024(B * 4)
    c = ws[off_cnt:off_cnt + 20].view(torch.int32).tolist()
    nrm = rows.square().sum(1)
    print(f"{tag}: B={B} N={N}: {ms:.2f} ms, fallback rows {c[0]}/{B*N} (overflow {c[1]}, too few {c[2]}, >64 survivors {c[3]}, certificate {c[4]}); "
          f"|x|^2 mean {nrm.mean().item():.3g} max {nrm.max().item():.3g}; |x-mu|^2 mean {(rows-rows.mean(0)).square().sum(1).mean().item():.3g}", flush=True)
torch.manual_seed(0)
for B, N in ((1, 32768),):
    run(torch.randn(B * N, 64, device="cuda"), B, N, 30, "randn")
net = pkg.DGCNN(5, 4, 30, cuda=True).eval()
for B in (1, 4):
    x = blocks(B, 32768)
    with torch.no_grad():
        x1 = net(x, depth=1).permute(0, 2, 1).reshape(B * 32768, 64).contiguous()
        x2 = net(x, depth=2).permute(0, 2, 1).reshape(B * 32768, 64).contiguous()
    run(x1, B, 32768, 30, "model x1"); run(x2, B, 32768, 30, "model x2")
