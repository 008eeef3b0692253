import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import __graft_entry__ as entry
import bench
pkg = entry.load_package(); ops = pkg.ops
x, y, m = bench.make_batch(0, 8); x, y, m = x.cuda(), y.cuda(), m.cuda()
net = pkg.DGCNN(5, 4, 30, dropout=0.5, cuda=True).train()
for det in (False, True):
    ops.set_deterministic(det)
    for _ in range(3): ops.seg_loss(net(x), y, m).backward()
    torch.cuda.synchronize(); t = time.time()
    for _ in range(10): ops.seg_loss(net(x), y, m).backward()
    torch.cuda.synchronize(); print("deterministic", det, (time.time() - t) / 10 * 1e3, "ms fwd+criterion+bwd (eager)", flush=True)
