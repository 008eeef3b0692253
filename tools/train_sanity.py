"""Short training run on a learnable synthetic task (label = height band of the point): the loss must fall and stay
finite.  Exercises the full step (fwd + loss + bwd + Adam) repeatedly, with dropout on and BN statistics drifting."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import __graft_entry__ as entry

pkg = entry.load_package()
torch.manual_seed(0)
B, N, steps = 4, bench.NPTS, int(sys.argv[1]) if len(sys.argv) > 1 else 150
model = pkg.DGCNN(bench.CLASSES, bench.FEAT, 20, dropout=0.5, emb_dims=1024, cuda=True).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
losses = []
for it in range(steps):
    x, _ = bench.make_batch(it, B)
    x = x.cuda()
    z = x[:, 2, :]
    q = torch.quantile(z.flatten().float(), torch.linspace(0, 1, bench.CLASSES + 1, device="cuda")[1:-1])
    y = torch.bucketize(z, q).long()                       # 5 height bands
    opt.zero_grad()
    logits = model(x)
    loss = bench.seg_loss(logits, y)
    loss.backward()
    opt.step()
    losses.append(float(loss))
    if it % 25 == 0 or it == steps - 1:
        acc = (logits.argmax(1) == y).float().mean().item()
        print(f"step {it:4d} loss {losses[-1]:.4f} acc {acc:.3f}", flush=True)
assert all(l == l and l < 1e4 for l in losses), "non-finite loss"
# the reference's loss is a label-smoothed CE on softmax(logits) (train.py:180,197): its floor is ~1.1, so accuracy is the criterion
assert losses[-1] < losses[0] and acc > 0.9, "the model did not learn the height bands"
print("ok: loss", losses[0], "->", losses[-1], "accuracy", acc)
