"""Diagnostic (2 ranks): where does the overlapped bucket exchange stall?  Phases print as they complete; a phase that
takes longer than 25 s dumps every thread's stack and exits."""
import faulthandler, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench
import __graft_entry__ as entry

pkg = entry.load_package()
parallel = __import__(entry.PKG_NAME + ".parallel", fromlist=["x"])
rank, local, world = parallel.init_from_env()
dev = torch.device("cuda", local)
torch.manual_seed(0)
model = pkg.DGCNN(bench.CLASSES, bench.FEAT, bench.KNN, dropout=0.0, emb_dims=bench.EMB, cuda=True).train()
parallel.broadcast_parameters(model)
x, y, m = (t.to(dev) for t in bench.make_batch(1000 * rank, bench.B_PER_GPU))
bucket = parallel.GradBucket(model.parameters())
tail = bucket._offsets[id(model.conv6[0].weight)]


def say(msg):
    print(f"[rank {rank}] {msg}", flush=True)


def phase(name, fn, n=2):
    faulthandler.dump_traceback_later(25, exit=True)
    t = time.time()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    faulthandler.cancel_dump_traceback_later()
    say(f"{name}: ok, {1e3 * (time.time() - t) / n:.2f} ms per step (host clock, synchronised)")


def fwd_bwd():
    bucket.zero()
    loss = pkg.ops.seg_loss(model(x), y, m)
    loss.backward()


def plain():
    fwd_bwd()
    dist.all_reduce(bucket.flat)


def async_main_thread():          # async collective issued from the main thread after the backward (no overlap)
    fwd_bwd()
    w = dist.all_reduce(bucket.flat[tail:], async_op=True)
    dist.all_reduce(bucket.flat[:tail])
    w.wait()


def hook_sync_only():             # hook fires in the autograd thread but only records that it ran
    fired = []
    h = model.conv6[0].weight.register_post_accumulate_grad_hook(lambda p: fired.append(torch.cuda.current_stream().cuda_stream))
    fwd_bwd()
    h.remove()
    dist.all_reduce(bucket.flat)
    say(f"hook fired {len(fired)}x, stream in hook {fired[:1]}, main stream {torch.cuda.current_stream().cuda_stream}")


def hook_overlap():
    bucket.overlap_from(model.conv6[0].weight)
    fwd_bwd()
    bucket.all_reduce_mean()


say(f"start: tail at {tail} of {bucket.flat.numel()}")
phase("plain", plain)
phase("hook + early all-reduce", hook_overlap)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True, capturable=True)


def full_step(x, y, m):
    fwd_bwd()
    bucket.all_reduce_mean()
    opt.step()
    return bucket.flat


phase("5 overlapped steps + Adam without a host sync in between", lambda: full_step(x, y, m), n=5)
graphs = __import__(entry.PKG_NAME + ".graphs", fromlist=["x"])
faulthandler.dump_traceback_later(40, exit=True)
try:
    g = graphs.GraphedTrainStep(full_step, (x, y, m))
    say("captured")
    for _ in range(3):
        g.run(x, y, m)
    torch.cuda.synchronize()
    say("graph replays: ok")
except Exception as e:
    import traceback
    say("capture failed: " + repr(e))
    traceback.print_exc()
faulthandler.cancel_dump_traceback_later()
dist.destroy_process_group()
