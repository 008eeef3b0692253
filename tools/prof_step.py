"""One warm-up + STEPS training steps of the bench workload (for ncu launch lists / captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import __graft_entry__ as entry

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.B_PER_GPU
pkg = entry.load_package()
torch.manual_seed(0)
model = pkg.DGCNN(bench.CLASSES, bench.FEAT, bench.KNN, dropout=0.5, emb_dims=bench.EMB, cuda=True).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
x, y, m = bench.make_batch(0, B)
x, y, m = x.cuda(), y.cuda(), m.cuda()
for i in range(1 + steps):
    opt.zero_grad()
    loss = pkg.ops.seg_loss(model(x), y, m)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
