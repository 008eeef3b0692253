import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import __graft_entry__ as entry
from golden import common
pkg = entry.load_package(); ops = pkg.ops
DEV = "cuda:0"
def run(fused, N, k, B=2):
    ops.FUSED_BWD = fused
    x = common.points(51, B, 4, N, "real").to(DEV)
    w = common.weights(52, 5, 4, 1024)
    net = pkg.DGCNN(5, 4, k, dropout=0.0, cuda=True); net.load_state_dict(w, strict=False); net.train()
    y = net(x); y.square().mean().backward(); torch.cuda.synchronize()
    return y.detach(), [p.grad.clone() for p in net.parameters()]
for N, k in ((1003, 30), (517, 20), (300, 48), (260, 7)):
    y1, g1 = run(True, N, k); y0, g0 = run(False, N, k)
    print(N, k, float((y1 - y0).abs().max()), max(float((a - b).norm() / (a.norm() + 1e-30)) for a, b in zip(g0, g1)), flush=True)
# timing at the training shape
B, N, k = 8, 7000, 30
x = common.points(1, B, 4, N, "real").to(DEV)
net = pkg.DGCNN(5, 4, k, dropout=0.5, cuda=True).train()
for fused in (False, True, False, True):
    ops.FUSED_BWD = fused
    for _ in range(3):
        net(x).square().mean().backward()
    torch.cuda.synchronize(); t = time.time()
    for _ in range(10):
        net(x).square().mean().backward()
    torch.cuda.synchronize(); print("fused", fused, (time.time() - t) / 10 * 1e3, "ms fwd+bwd", flush=True)
ops.LOG.reset(timing=True)
for _ in range(5):
    net(x).square().mean().backward()
torch.cuda.synchronize()
for kname, v in sorted(ops.LOG.summary().items(), key=lambda kv: -kv[1][1]):
    print(f"{kname:44s} {v[0] // 5:3d} calls/step {v[1] / 5 * 1e3:8.1f} us/step")
