import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, common
import __graft_entry__ as entry
pkg = entry.load_package()
x = common.points(31, 2, 4, 96, "real").cuda()
net = pkg.DGCNN(5, 4, 8, cuda=True).eval()
for d in (1, 2, 3, 10):
    y = net(x, depth=d); torch.cuda.synchronize(); print("depth", d, "ok", y.abs().max().item(), flush=True)
