"""Inference timings (eval mode): cfg1 README example, cfg2 shape, cfg4 tile inference; kNN sweep (cfg5)."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as entry
pkg = entry.load_package(); ops = pkg.ops
graphs = __import__(entry.PKG_NAME + ".graphs", fromlist=["x"])

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def blocks(B, N, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.cat((torch.rand(B, 2, N, generator=g) * 30, 100 + torch.rand(B, 1, N, generator=g) * 30, torch.rand(B, 1, N, generator=g) * 20), 1).cuda()

out = {}
for name, B, N, k in (("cfg1 B=1 N=1000 k=20", 1, 1000, 20), ("cfg2-shape B=8 N=7000 k=30", 8, 7000, 30), ("cfg4 B=16 N=32768 k=30", 16, 32768, 30)):
    torch.manual_seed(0)
    net = pkg.DGCNN(5, 4, k, cuda=True).eval()
    x = blocks(B, N)
    ms = timeit(lambda: net(x))
    ops.LOG.reset(timing=True); net(x); torch.cuda.synchronize(); ker = {kk: round(v[1], 3) for kk, v in ops.LOG.summary().items()}; ops.LOG.reset()
    out[name] = {"ms": round(ms, 3), "points_per_s": round(B * N / ms * 1e3), "kernels_ms": ker}
    if B * N <= 100000:          # launch-bound shapes: the same forward replayed from a CUDA graph (graphs.GraphedForward)
        gf = graphs.GraphedForward(net, x)
        msg = timeit(lambda: gf.run(x), n=20, warm=3)
        out[name]["ms_cuda_graph"] = round(msg, 3)
        del gf
    print(name, out[name], flush=True)
    del net, x; torch.cuda.empty_cache()
# kNN sweep
sweep = []
for C in (() if os.environ.get("NO_SWEEP") else (4, 64, 128)):
    for N in (1000, 2000, 4000, 7000, 8000, 16000, 32000, 64000):
        for k in (20, 30, 40):
            # B = 1 (the survey's single-cloud sweep) and a batch that fills the GPU (8 clouds up to N = 8000)
            for B in ((1, 8) if N <= 8000 else (1,)):
                x = (torch.randn(B, C, N, generator=torch.Generator().manual_seed(N + C)) ).cuda()
                rows = ops.pack_rows(x)
                try:
                    ms = timeit(lambda: ops.knn_rows(rows, B, C, N, k), n=5, warm=3)
                except Exception as e:
                    ms = None
                sweep.append({"C": C, "N": N, "k": k, "B": B, "ms": None if ms is None else round(ms, 3),
                              "tflops_alg": None if ms is None else round(2.0 * B * N * N * C / ms / 1e9, 2)})
    print("knn sweep C", C, [(s["B"], s["N"], s["k"], s["ms"]) for s in sweep if s["C"] == C], flush=True)
out["knn_sweep"] = sweep
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "infer_r2.json"), "w"), indent=1)
