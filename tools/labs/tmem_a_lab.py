"""tcgen05.mma with A in tensor memory: checks the TMEM operand layout on one 128x64x64 tile."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
pkg = entry.load_package()
lib = pkg._lib.load()
torch.manual_seed(0)
X = torch.randn(128, 64, device="cuda"); Y = torch.randn(64, 64, device="cuda")
ref = X.bfloat16().double() @ Y.bfloat16().double().t()
for swap in (0, 1):
    D = torch.full((128, 64), float("nan"), device="cuda")
    rc = lib.dgcnn_debug_tmem_a_lab(C.c_void_p(X.data_ptr()), C.c_void_p(Y.data_ptr()), C.c_void_p(D.data_ptr()), swap,
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (D.double() - ref).abs().max().item()
    print("swap", swap, "rc", rc, "max err", err, "ref scale", ref.abs().max().item())
