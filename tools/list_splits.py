"""Print every operand split / plane GEMM of one training step with its shape (which producers are worth fusing)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import __graft_entry__ as entry
pkg = entry.load_package(); ops = pkg.ops
model = pkg.DGCNN(bench.CLASSES, bench.FEAT, bench.KNN, dropout=0.5, emb_dims=bench.EMB, cuda=True).train()
x, y, m = bench.make_batch(0, bench.B_PER_GPU)
x, y, m = x.cuda(), y.cuda(), m.cuda()
for _ in range(2):
    ops.seg_loss(model(x), y, m).backward()
torch.cuda.synchronize()
orig = ops._call
rows = []
def spy(name, dev, *args):
    if name in ("dgcnn_tc_split_planes_f32", "dgcnn_tc_gemm_planes_f32", "dgcnn_tc_gemm_f32", "dgcnn_rows_bn_lrelu_fwd", "dgcnn_rows_bn_lrelu_bwd", "dgcnn_rows_colsum_f32"):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); orig(name, dev, *args); b.record(); torch.cuda.synchronize()
        ints = [v for v in args if isinstance(v, int) and 0 < v < 10**6]
        rows.append((name, ints[:8], a.elapsed_time(b) * 1e3))
    else:
        orig(name, dev, *args)
ops._call = spy
ops.seg_loss(model(x), y, m).backward()
torch.cuda.synchronize()
for r in rows:
    print(f"{r[0]:32s} {str(r[1]):48s} {r[2]:7.1f} us")
