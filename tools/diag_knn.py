"""Diagnostic: tensor-core kNN vs the exact FP32-FMA kernel — identical tables, timing, fallback counters.

  python tools/diag_knn.py [quick]

Shapes: the model's three graphs at cfg2 (C=4 real-like, C=64 features), cfg1, cfg4 (N=32768) and a C=128 case."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import __graft_entry__ as entry

pkg = entry.load_package()
ops = pkg.ops
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def run(rows, B, Cn, N, k, tag, exact=True):
    stats, stats_u = [], []
    idx = ops.knn_rows(rows, B, Cn, N, k, method="tc", stats=stats)
    idx_u = ops.knn_rows(rows, B, Cn, N, k, method="tc", stats=stats_u, ordered=False)
    t_tc = timed(lambda: ops.knn_rows(rows, B, Cn, N, k, method="tc"), 5)
    t_u = timed(lambda: ops.knn_rows(rows, B, Cn, N, k, method="tc", ordered=False), 5)
    msg = (f"{tag}: B={B} C={Cn} N={N} k={k}: tc set {t_u:.3f} ms ({2.0 * B * N * N * Cn / t_u / 1e9:.1f} TFLOP/s) fallback {stats_u}, "
           f"ordered {t_tc:.3f} ms fallback {stats}, set==ordered: {torch.equal(idx_u.sort(1)[0], idx.sort(1)[0])}")
    if exact:
        idx2 = ops.knn_rows(rows, B, Cn, N, k, method="exact")
        t_ex = timed(lambda: ops.knn_rows(rows, B, Cn, N, k, method="exact"), 2)
        same = torch.equal(idx, idx2)
        msg += f", exact {t_ex:.3f} ms, equal={same}"
        if not same:
            bad = (idx != idx2).any(1).nonzero().flatten()
            msg += f" ({bad.numel()} rows differ, first {bad[:5].tolist()})"
        if not torch.equal(idx_u.sort(1)[0], idx2.sort(1)[0]):
            bad = (idx_u.sort(1)[0] != idx2.sort(1)[0]).any(1).nonzero().flatten()
            msg += f" SET MISMATCH ({bad.numel()} rows, first {bad[:5].tolist()})"
    if Cn <= 16 and N >= 512:
        t_sm = timed(lambda: ops.knn_rows(rows, B, Cn, N, k, method="small"), 3)
        msg += f", small {t_sm:.3f} ms"
    print(msg, flush=True)


torch.manual_seed(0)
B, N, k = 8, 7000, 30
x, _, _ = bench.make_batch(0, B)
rows4 = ops.pack_rows(x.cuda())
run(rows4, B, 4, N, k, "cfg2 layer 1 (real-like xyz+agl)")
run(torch.randn(B * N, 64, device="cuda"), B, 64, N, k, "randn64")
model = pkg.DGCNN(5, 4, k, cuda=True).train()
with torch.no_grad():
    x1 = model(x.cuda(), depth=1).permute(0, 2, 1).reshape(B * N, 64).contiguous()
    x2 = model(x.cuda(), depth=2).permute(0, 2, 1).reshape(B * N, 64).contiguous()
run(x1, B, 64, N, k, "cfg2 layer 2 (model x1)")
run(x2, B, 64, N, k, "cfg2 layer 3 (model x2)")
run(torch.rand(1000, 4, device="cuda"), 1, 4, 1000, 20, "cfg1 layer 1")
run(torch.randn(1000, 64, device="cuda"), 1, 64, 1000, 20, "cfg1 C=64")
run(torch.randn(2 * 600, 128, device="cuda"), 2, 128, 600, 20, "C=128 small")
run(torch.randn(3 * 300, 64, device="cuda"), 3, 64, 300, 40, "N=300 k=40")
run(torch.randn(2 * 200, 12, device="cuda"), 2, 12, 200, 48, "N=200 (no threshold) k=48")
if not quick:
    run(torch.randn(8 * 7000, 128, device="cuda"), 8, 128, 7000, 30, "C=128 cfg2-size")
    run(torch.randn(2 * 32768, 64, device="cuda"), 2, 64, 32768, 30, "cfg4 C=64 (2 blocks)")
    xr = torch.cat((torch.rand(2, 32768, 2) * 30, 100 + torch.rand(2, 32768, 1) * 30, torch.rand(2, 32768, 1) * 20), 2).reshape(-1, 4).cuda()
    run(xr.contiguous(), 2, 4, 32768, 30, "cfg4 C=4 (2 blocks)")
    run(torch.randn(1 * 65536, 64, device="cuda"), 1, 64, 65536, 20, "N=64k C=64", exact=False)
