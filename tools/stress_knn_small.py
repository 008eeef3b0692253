"""Stress: the tensor-core kNN (set mode) against the exact kernel at the cfg1 shape, many calls in one process with the
caching allocator full of garbage.  python tools/stress_knn_small.py [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
import common
import __graft_entry__ as entry
pkg = entry.load_package(); ops = pkg.ops
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = "cuda:0"
bad = {}
SHAPES = ((1, 4, 1000, 20, "unit"), (1, 64, 1000, 20, "randn"), (2, 4, 515, 30, "real"), (1, 4, 1000, 20, "real"))
if len(sys.argv) > 2 and sys.argv[2] == "big":      # the training shapes and a C = 128 case
    SHAPES = ((8, 64, 7000, 30, "randn"), (8, 4, 7000, 30, "real"), (3, 64, 2000, 30, "randn"), (2, 128, 3000, 20, "randn"))
for (B, C, N, k, kind) in SHAPES:
    x = common.points(5, B, C, N, kind if kind != "randn" else "unit").to(dev)
    if kind == "randn":
        x = torch.randn(B, C, N, device=dev)
    rows = ops.pack_rows(x)
    ref = ops.knn_rows(rows, B, C, N, k, method="exact").sort(1)[0]
    nb = 0
    for it in range(iters):
        junk = [torch.full((1 << 22,), float("nan"), device=dev) for _ in range(4)]      # poison the allocator's free blocks
        junk2 = torch.randint(-2**31, 2**31 - 1, (1 << 22,), device=dev, dtype=torch.int32)
        del junk, junk2
        stats = []
        idx = ops.knn_rows(rows, B, C, N, k, method="tc", ordered=(it % 2 == 1), stats=stats).sort(1)[0]
        if not torch.equal(idx, ref):
            nb += 1
            rowsbad = (idx != ref).any(1).nonzero().flatten()
            if nb <= 3:
                print(f"MISMATCH B={B} C={C} N={N} k={k} {kind} it={it} ordered={it % 2 == 1}: {rowsbad.numel()} rows {rowsbad[:6].tolist()} fallback {stats}", flush=True)
    bad[(B, C, N, k, kind)] = nb
    print(f"B={B} C={C} N={N} k={k} {kind}: {nb} / {iters} calls differ from the exact kernel", flush=True)
sys.exit(1 if any(bad.values()) else 0)
