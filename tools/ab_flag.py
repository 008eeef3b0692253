import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import __graft_entry__ as entry
from golden import common
pkg = entry.load_package(); ops = pkg.ops
B, N, k = 8, 7000, 30
x = common.points(1, B, 4, N, "real").to("cuda:0")
net = pkg.DGCNN(5, 4, k, dropout=0.5, cuda=True).train()
name = sys.argv[1]
for val in (False, True, False, True):
    setattr(ops, name, val)
    for _ in range(3): net(x).square().mean().backward()
    torch.cuda.synchronize(); t = time.time()
    for _ in range(20): net(x).square().mean().backward()
    torch.cuda.synchronize(); print(name, val, (time.time() - t) / 20 * 1e3, "ms fwd+bwd", flush=True)
