import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
pkg = entry.load_package(); ops = pkg.ops
mode = sys.argv[1]
B, N, k = 2, 96, 8
torch.manual_seed(0)
rows = torch.randn(B * N, 64, device="cuda")
idx = torch.randint(0, N, (B * N, k), dtype=torch.int32, device="cuda")
wa = torch.randn(64, 128, device="cuda") * 0.1
wb = torch.randn(64, 64, device="cuda") * 0.1
aff = torch.cat((torch.zeros(64), torch.ones(64), torch.zeros(64))).cuda()
if mode == "single":
    y = ops.edgeconv_eval(rows, idx, B, 64, N, k, wa, aff)
else:
    y = ops.edgeconv_eval(rows, idx, B, 64, N, k, wa, aff, wb, aff)
torch.cuda.synchronize(); print(mode, "ok", y.abs().max().item())
pkg._lib.load().dgcnn_set_edgeconv_tc(0)
y2 = ops.edgeconv_eval(rows, idx, B, 64, N, k, wa, aff) if mode == "single" else ops.edgeconv_eval(rows, idx, B, 64, N, k, wa, aff, wb, aff)
print("max diff vs ffma", (y - y2).abs().max().item())
