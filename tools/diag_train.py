"""Diagnostic: gradient parity table (ours vs CPU-fp32 oracle vs CPU-fp64 oracle vs eager-GPU oracle)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, common
import __graft_entry__ as entry
from oracle import dgcnn_oracle as O

def rel(a, b): return (a.double() - b.double()).norm().item() / max(b.double().norm().item(), 1e-30)

def run_oracle(w, x, k, cot, forced, dtype=torch.float32, dev="cpu"):
    po = {kk: (v.clone().to(dtype) if v.is_floating_point() else v.clone()).to(dev) for kk, v in w.items()}
    for key, v in po.items():
        if v.is_floating_point() and "running" not in key: v.requires_grad_(True)
    ora = O.OracleDGCNN(po, k, dropout=0.0, training=True)
    if dev != "cpu":
        # graph_feature builds idx_base on CPU; patch arange device
        import types
    out = ora(x.to(dtype).to(dev), forced_idx=forced)
    (out * cot.to(dtype).to(dev)).sum().backward()
    return out.detach(), {n: po[n].grad.detach().cpu() for n in po if po[n].requires_grad}, ora

B, N, k, kind, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
pkg = entry.load_package()
x = common.points(seed, B, 4, N, kind); w = common.weights(seed + 1, 5, 4, 1024)
o32, g32, ora = run_oracle(w, x, k, torch.zeros(1), None) if False else (None, None, None)
ora0 = O.OracleDGCNN({kk: v.clone() for kk, v in w.items()}, k, dropout=0.0, training=True)
ref = ora0(x); forced = [ora0.trace[f"idx{i}"] for i in (1, 2, 3)]
cot = common.cotangent(seed + 2, tuple(ref.shape))
o32, g32, _ = run_oracle(w, x, k, cot, forced, torch.float32)
o64, g64, _ = run_oracle(w, x, k, cot, forced, torch.float64)
net = pkg.DGCNN(5, 4, k, dropout=0.0, cuda=True); net.load_state_dict(w, strict=False); net.train()
out = net(x.cuda(), forced_idx=forced); (out * cot.cuda()).sum().backward()
print(f"logits: ours-vs-f64 {rel(out.detach().cpu(), o64):.2e}  cpu32-vs-f64 {rel(o32, o64):.2e}")
print(f"{'param':18s} {'ours/f64':>10s} {'cpu32/f64':>10s} {'ours/cpu32':>10s}")
for n, p in net.named_parameters():
    print(f"{n:18s} {rel(p.grad.cpu(), g64[n]):10.2e} {rel(g32[n], g64[n]):10.2e} {rel(p.grad.cpu(), g32[n]):10.2e}")
