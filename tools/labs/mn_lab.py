import sys, os, ctypes as C, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
pkg = entry.load_package(); lib = pkg._lib.load()
torch.manual_seed(0)
X = torch.randn(128, 64, device="cuda"); Y = torch.randn(128, 64, device="cuda")
Xb, Yb = X.bfloat16().float(), Y.bfloat16().float()
S = torch.cat((Xb, Yb), 1)                      # [128 rows, 128 m]
ref = S.t() @ Yb                               # [128 m, 64 n]
ref_top_only = Xb.t() @ Yb
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
best = []
for lbo_a, sbo_a, lbo_b, sbo_b, kstep in itertools.product((16384, 1024, 128, 16), (1024, 16384, 128), (16384, 1024, 128, 16), (1024, 128), (2048, 32, 256)):
    D = torch.zeros(128, 64, device="cuda")
    rc = lib.dgcnn_debug_mn_lab(C.c_void_p(X.data_ptr()), C.c_void_p(Y.data_ptr()), C.c_void_p(D.data_ptr()), lbo_a, sbo_a, lbo_b, sbo_b, kstep, 1, 1, st)
    torch.cuda.synchronize()
    e_all = (D - ref).abs().max().item(); e_top = (D[:64] - ref_top_only).abs().max().item()
    best.append((e_all, e_top, lbo_a, sbo_a, lbo_b, sbo_b, kstep))
best.sort()
for b in best[:8]: print("err_all %.3e err_top %.3e  lbo_a %d sbo_a %d lbo_b %d sbo_b %d kstep %d" % b)
best.sort(key=lambda t: t[1])
for b in best[:4]: print("TOP: err_all %.3e err_top %.3e  lbo_a %d sbo_a %d lbo_b %d sbo_b %d kstep %d" % b)
print("ref scale", ref.abs().max().item())
