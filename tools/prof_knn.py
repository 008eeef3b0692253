"""One tensor-core kNN call on model-like features (cfg2 shape) for ncu captures: python tools/prof_knn.py [C]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
pkg = entry.load_package()
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B, N, k = 8, 7000, 30
torch.manual_seed(0)
rows = torch.randn(B * N, max(C, 4), device="cuda")
for _ in range(3):
    idx = pkg.ops.knn_rows(rows, B, C, N, k, method="tc")
torch.cuda.synchronize()
print("ok", int(idx.sum()))
