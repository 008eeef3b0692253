#!/usr/bin/env python
"""bench.py — headline benchmark of the DGCNN EdgeConv hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|infer]

Workload (BASELINE.json configs[1]/[2]): AHN3 training shape, B=8 blocks per GPU of N=7000 points, F=4, k=30,
5 classes, emb 1024.  One step = forward + the reference's masked criterion (softmax -> focus class for unmasked points
-> label-smoothed CE on the probabilities, train.py:180-197 / util.py:17-34; class-balanced masks as data.py:80-105
makes them) + backward + gradient all-reduce (N>1) + Adam update (reference train.py:118).  `--workload infer` is
BASELINE configs[3]: tile inference, batches of 16 blocks of N=32768 points, k=30, blocks sharded over the GPUs with no
collective (test.py:281-330), softmax/max/threshold on device; metric inference points/s.  Synthetic
"real-like" blocks (xy in [0,30) m, z in [100,130) m, agl in [0,20) m), random-init weights.
Metric: training points/s, whole job (all GPUs).

`value`   : inputs resident in HBM, device-timed (CUDA events), max over ranks.
`e2e`     : same step through the public API with pinned-host inputs copied H2D and the loss read back D2H
            inside the timed region.
`roofline`: dominant kernel's algorithmic FLOP/s (CUDA events around its launches inside the timed region).
`cpu_baseline` / `--impl reference`: the CPU oracle (a restatement of the reference's model.py; the reference
            itself is Python-on-ATen and cannot travel to the GPU box) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, NPTS, FEAT, KNN, CLASSES, EMB = 8, 7000, 4, 30, 5, 1024
METRIC, UNIT = "train points/sec (N=7000,k=30)", "points/s"


# ------------------------------------------------------------------------------------------------ data / loss
def make_batch(seed: int, B: int):
    g = torch.Generator().manual_seed(seed)
    x = torch.cat((torch.rand(B, 2, NPTS, generator=g) * 30, 100 + torch.rand(B, 1, NPTS, generator=g) * 30,
                   torch.rand(B, 1, NPTS, generator=g) * 20), dim=1)
    y = torch.randint(0, CLASSES, (B, NPTS), generator=g)
    # class-balanced training mask of every block, as the loader side builds it (data.py:80-105): n = min(smallest class
    # count, N // classes) random points per class
    n_per = min(int(torch.stack([(y == c).sum(1) for c in range(CLASSES)]).min()), NPTS // CLASSES)
    keys = torch.rand(B, NPTS, generator=g)
    mask = torch.zeros(B, NPTS, dtype=torch.uint8)
    for c in range(CLASSES):
        kc = torch.where(y == c, keys, torch.full_like(keys, 2.0))
        mask.scatter_(1, kc.argsort(dim=1)[:, :n_per], 1)
    # the reference's loaders produce [B,N,F]; the model sees the permute(0,2,1) view (train.py:176)
    return x.permute(0, 2, 1).contiguous().permute(0, 2, 1), y, mask


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons via NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        while self.ok and not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_model_name() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_reference_steps(steps: int, warmup: int, sample_blocks: int):
    """Time the CPU oracle's training step (fwd + masked criterion + bwd + Adam) on `sample_blocks` blocks of the workload."""
    from oracle import dgcnn_oracle as O
    from oracle import drivers_oracle as D
    torch.set_num_threads(os.cpu_count() or 1)
    params = O.init_params(CLASSES, FEAT, EMB, seed=0)
    leaves = []
    for name, v in params.items():
        if v.is_floating_point() and "running" not in name:
            v.requires_grad_(True)
            leaves.append(v)
    net = O.OracleDGCNN(params, KNN, dropout=0.5, training=True)
    opt = torch.optim.Adam(leaves, lr=1e-3, weight_decay=1e-4)
    x, y, mask = make_batch(0, sample_blocks)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss, _ = D.train_criterion(net(x), y, mask.numpy(), CLASSES)      # vectorised form of train.py:180-197
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def cpu_reference_infer(steps: int, warmup: int, n_points: int):
    """The reference's predict() loop body (test.py:291-306, B=1) on one block of `n_points` points."""
    from oracle import dgcnn_oracle as O
    from oracle import drivers_oracle as D
    torch.set_num_threads(os.cpu_count() or 1)
    params = O.init_params(CLASSES, FEAT, EMB, seed=0)
    net = O.OracleDGCNN(params, KNN, training=False)
    x = make_blocks(0, 1, n_points).permute(0, 2, 1).float()
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            D.predict_labels(net(x), 0.8)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def make_blocks(seed: int, nb: int, n: int) -> torch.Tensor:
    """[nb, n, F] float32 blocks in the reference's block layout (test.py:276-299)."""
    g = torch.Generator().manual_seed(seed)
    return torch.cat((torch.rand(nb, n, 2, generator=g) * 30, 100 + torch.rand(nb, n, 1, generator=g) * 30,
                      torch.rand(nb, n, 1, generator=g) * 20), dim=2)


INFER_B, INFER_N = 16, 32768


def workload_config(world: int, workload: str = "train") -> dict:
    """`config` of the JSON line: identical for our arm and for the reference arm."""
    if workload == "infer":
        return {"workload": f"AHN3 tile inference (predict(): fwd + softmax/max/threshold) {INFER_B} blocks/GPU/batch F={FEAT} N={INFER_N} k={KNN} classes={CLASSES} emb={EMB}",
                "parallelism": f"blocks sharded over {world} GPU(s), no collective", "global_batch": INFER_B * world,
                "l2": "inputs larger than L2: a batch streams > 8 GB of per-edge work; 2 rotating block sets"}
    return {"workload": f"AHN3 train step (fwd+masked criterion+bwd+allreduce+Adam) B={B_PER_GPU}/GPU F={FEAT} N={NPTS} k={KNN} classes={CLASSES} emb={EMB}",
            "parallelism": f"dp{world}", "global_batch": B_PER_GPU * world,
            "l2": "no explicit flush: each step streams >1.5 GB of per-edge intermediates (L2 is 126 MB); 4 rotating input batches"}


def run_reference(args, rank: int):
    if rank != 0:
        return
    if args.workload == "infer":
        sample_n = 8192                      # the reference's B x N x N matrices: 32768^2 floats x 3 per layer do not fit the step budget
        times, cores = cpu_reference_infer(args.steps, args.warmup, sample_n)
        total = sum(times)
        value = sample_n * len(times) / total
        sample = (f"one block of N={sample_n} points per step (B=1 like test.py:291-306; N={INFER_N} needs 3 x 4.3 GB distance "
                  f"matrices per layer and ~16x the time), {args.warmup} warm-up + {args.steps} timed steps")
        metric = INFER_METRIC
    else:
        sample_blocks = B_PER_GPU
        times, cores = cpu_reference_steps(args.steps, args.warmup, sample_blocks)
        total = sum(times)
        value = sample_blocks * NPTS * len(times) / total
        sample = f"the full batch: {sample_blocks} blocks (N={NPTS}) per step, {args.warmup} warm-up + {args.steps} timed steps"
        metric = METRIC
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, args.workload),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "cpu": cpu_model_name()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ roofline bookkeeping
def kernel_flops(name: str, B: int, N: int = NPTS):
    """Algorithmic FLOPs of ALL launches of one entry point in one step."""
    M = B * N
    E = M * KNN
    if name == "dgcnn_knn_tc_f32[C=64]":              # the two feature-space graphs (layers 2 and 3): 2*B*N^2*C each
        return 2.0 * B * N * N * (64 + 64)
    if name == f"dgcnn_knn_tc_f32[C={FEAT}]":         # the input-space graph (same kernel, C padded to one 64-wide K atom)
        return 2.0 * B * N * N * FEAT
    if name == "dgcnn_edgeconv_conv_a_train":         # conv1 (2F->64), conv3, conv5 (128->64) when run as per-edge GEMMs
        return 2.0 * E * 64 * (2 * FEAT + 128 + 128)
    if name == "dgcnn_edgeconv_lin_conv_a":           # conv1 + conv3 as y = U_j + V_i (algorithmic work of the per-edge convolution)
        return 2.0 * E * 64 * (2 * FEAT + 128)
    if name == "dgcnn_edgeconv_lin_extrema":          # conv5 + max over k
        return 2.0 * E * 64 * 128
    if name == "dgcnn_edgeconv_conv_b_train":         # conv2, conv4 (64->64)
        return 2.0 * E * 64 * 64 * 2
    if name in ("dgcnn_edgeconv_bwd_b", "dgcnn_edgeconv_bwd_b_fused"):   # per block: dW_b and dh GEMMs
        return 2.0 * E * 64 * 64 * 2 * 2
    if name == "dgcnn_tc_gemm_planes_f32":            # conv7/8/9 forward + dgrad + wgrad, conv6 backward, EdgeConv point-level GEMMs
        head = 2.0 * M * (192 * 512 + 512 * 256 + 256 * CLASSES) * 3 + 2.0 * M * 192 * 192 * 2
        edge = 2.0 * M * 128 * (64 + 64) * 2 + 2.0 * M * 128 * FEAT
        return head + edge
    if name == "dgcnn_tc_gemm_bn_lrelu_segmax_f32":   # conv6 + bn6 + LeakyReLU + global max
        return 2.0 * M * 192 * EMB
    if name == "dgcnn_edgeconv_fwd_eval":
        return 2.0 * E * 64 * (2 * FEAT + 64 + 128 + 64 + 128)
    return 0.0


def roofline_of(per_kernel: dict, steps: int, B: int, N: int, peaks: dict, sm_mhz: float):
    """Roofline entry of the entry point with the largest device time among those with an algorithmic FLOP count."""
    # the gather passes that REPLACE per-edge GEMMs are L2/HBM-bound by construction: they are not rated against the tensor roof
    dom, dom_ms = None, 0.0
    for name, (calls, tot) in per_kernel.items():
        if name.startswith("dgcnn_edgeconv_lin_"):
            continue
        if kernel_flops(name, B, N) > 0 and tot > dom_ms:
            dom, dom_ms = name, tot
    if not dom:
        return None
    fl = kernel_flops(dom, B, N)
    achieved = fl / (dom_ms / steps * 1e-3) / 1e12
    tensor_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    kind = ("measured sustained dense bf16 tensor peak (MEASURED_PEAKS.json); the kernels issue 3 (bf16x2 split) or 6 (bf16x3) "
            "kind::f16 MMAs per algorithmic product for FP32-level accuracy, so the ceiling for ALGORITHMIC flops is peak/3 .. peak/6")
    traffic = None
    try:        # measured DRAM bytes per launch from the committed `ncu --set full` capture of the TRAINING shape (not a live number)
        if (B, N) != (B_PER_GPU, NPTS):
            raise KeyError("no capture at this shape")
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(dom, {}).get("bytes_per_launch")
    except Exception:
        pass
    return {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
            "frac": achieved / tensor_peak, "traffic": traffic, "peak_kind": kind,
            "launches_per_step": per_kernel[dom][0] / steps, "ms_per_step": dom_ms / steps}


# ------------------------------------------------------------------------------------------------ extras
def _event_ms(fn, iters):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def layer_times(pkg, dev):
    """kNN + EdgeConv block (train mode) per layer at the bench shape: forward and forward+backward, microseconds, with the
    achieved algorithmic TFLOP/s of the fused layer (kNN Gram + EdgeConv convs, fwd+bwd = 3x the conv FLOPs)."""
    ops = pkg.ops
    g = torch.Generator().manual_seed(1)
    B, N, k = B_PER_GPU, NPTS, KNN
    E = B * N * k
    out = {}
    for name, Cn, two in (("L1 (C=4, conv1+conv2)", FEAT, True), ("L2 (C=64, conv3+conv4)", 64, True), ("L3 (C=64, conv5)", 64, False)):
        x = (torch.randn(B, Cn, N, generator=g) if Cn > 4 else make_batch(7, B)[0]).to(dev)
        rows = ops.pack_rows(x)
        wa = (torch.randn(64, 2 * Cn, 1, 1, generator=g) * 0.1).to(dev).requires_grad_(True)
        wb = (torch.randn(64, 64, 1, 1, generator=g) * 0.1).to(dev).requires_grad_(True) if two else None
        ga = torch.ones(64, device=dev, requires_grad=True)
        be = torch.zeros(64, device=dev, requires_grad=True)

        def fwd(req=False):
            r = rows.detach().requires_grad_(req)
            idx = ops.knn_rows(r.detach(), B, Cn, N, k, ordered=False)
            return r, ops.EdgeConvTrain.apply(r, idx, (B, Cn, N, k), wa, ga, be, wb, ga if two else None, be if two else None,
                                              ops.BNState(None, None), ops.BNState(None, None) if two else None)

        def fwd_bwd():
            r, o = fwd(True)
            o.sum().backward()

        for _ in range(2):
            fwd_bwd()
        t_f, t_fb = _event_ms(lambda: fwd(), 5), _event_ms(fwd_bwd, 5)
        gram = 2.0 * B * N * N * Cn
        conv = 2.0 * E * 64 * (2 * Cn + (64 if two else 0))
        out[name] = {"fwd": round(1e3 * t_f, 1), "fwd_bwd": round(1e3 * t_fb, 1),
                     "fwd_tflops": round((gram + conv) / t_f / 1e9, 1), "fwd_bwd_tflops": round((gram + 3 * conv) / t_fb / 1e9, 1)}
    return out


def eager_gpu_oracle(dev):
    """The CPU oracle's torch code run eagerly on this GPU (FP32, TF32 off): the library composition the kernels replace."""
    from oracle import dgcnn_oracle as O
    from oracle import drivers_oracle as D
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        params = {kk: v.to(dev) for kk, v in O.init_params(CLASSES, FEAT, EMB, seed=0).items()}
        leaves = []
        for name, v in params.items():
            if v.is_floating_point() and "running" not in name:
                v.requires_grad_(True)
                leaves.append(v)
        net = O.OracleDGCNN(params, KNN, dropout=0.5, training=True)
        opt = torch.optim.Adam(leaves, lr=1e-3, weight_decay=1e-4, fused=True)
        x, y, _ = make_batch(0, B_PER_GPU)
        x, y = x.to(dev), y.to(dev)

        def step():
            opt.zero_grad()
            D.validation_criterion(net(x), y, CLASSES)[0].backward()       # (the masked form's host loops would only add to it)
            opt.step()

        step()
        ms = _event_ms(step, 2)
        return {"ms_per_step": ms, "value": B_PER_GPU * NPTS / (ms * 1e-3), "unit": UNIT,
                "what": "oracle/dgcnn_oracle.py (torch ops of reference model.py) on this GPU, eager, FP32, same shape"}
    except Exception as e:                       # informative only: never fail the bench line on it
        return {"error": repr(e)[:200]}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        torch.cuda.empty_cache()


def load_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------ our arm: training
def run_ours(args, rank: int, local: int, world: int):
    import __graft_entry__ as entry
    import torch.distributed as dist
    pkg = entry.load_package()
    ops, parallel = pkg.ops, __import__(entry.PKG_NAME + ".parallel", fromlist=["x"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = pkg.DGCNN(CLASSES, FEAT, KNN, dropout=0.5, emb_dims=EMB, cuda=True)
    parallel.broadcast_parameters(model)
    model.train()
    bucket = parallel.GradBucket(model.parameters())
    overlap = world > 1 and args.overlap_collective and ops.FUSED_HEAD
    if overlap:
        # opt-in: conv6..conv9 (97 % of the bucket) are exchanged while the EdgeConv backward is still running
        bucket.overlap_from(model.conv6[0].weight)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True, capturable=not args.no_graph)

    nbuf = 4
    host = [make_batch(1000 * rank + i, B_PER_GPU) for i in range(nbuf)]
    host = [(x.pin_memory(), y.pin_memory(), m.pin_memory()) for x, y, m in host]
    resident = [(x.to(dev), y.to(dev), m.to(dev)) for x, y, m in host]
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 8 + host[0][2].numel()

    def step(x, y, m, update=True):
        bucket.zero()
        # callers hand the model the [B,N,F] -> permute(0,2,1) view (train.py:176); criterion = train.py:180-197 on device
        loss = ops.seg_loss(model(x), y, m)
        loss.backward()
        bucket.all_reduce_mean()
        if update:
            opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(max(args.warmup, 3)):
        step(*resident[i % nbuf])
    eager_step = step

    # ---- eager passes first (a captured graph ties the gradient accumulators to its capture stream; eager steps after the
    # capture pay cross-stream synchronisations) -------------------------------------------------------------------------
    # (1) per-entry-point device time: CUDA event pairs around every C-ABI call
    ops.LOG.reset(timing=True)
    ms_instr = timed(lambda i: eager_step(*resident[i % nbuf]), args.steps)
    per_kernel = ops.LOG.summary()
    # (2) launches per step and the optimizer alone (SURVEY §8d metric i): the eager step with and without the Adam update
    ops.LOG.reset(timing=False)
    ms_eager = timed(lambda i: eager_step(*resident[i % nbuf]), args.steps)
    launches = ops.LOG.launches
    ms_noopt = timed(lambda i: eager_step(*resident[i % nbuf], update=False), args.steps)
    # (3) the FP32-accurate gradient setting
    ms3 = None
    if world == 1 and not args.no_extras:
        old = ops.GRAD_NSPLIT
        try:
            ops.GRAD_NSPLIT = 3
            for i in range(3):
                eager_step(*resident[i % nbuf])
            ms3 = timed(lambda i: eager_step(*resident[i % nbuf]), args.steps)
        finally:
            ops.GRAD_NSPLIT = old

    # ---- CUDA graph of the step (one launch per step instead of ~220 kernel launches); with a gradient exchange (N > 1)
    # the graph covers forward + criterion + backward and the all-reduce + Adam stay eager --------------------------------
    graphed, graph_scope = None, "fwd+criterion+bwd+adam"
    if not args.no_graph:
        graphs = __import__(entry.PKG_NAME + ".graphs", fromlist=["x"])
        try:
            if world == 1:
                graphed = graphs.GraphedTrainStep(lambda x, y, m: eager_step(x, y, m), resident[0])
                step = lambda x, y, m: graphed.run(x, y, m)
            else:
                # NCCL collectives are capturable: the whole step (forward .. backward, the bucket all-reduce, Adam) is ONE graph
                # on every rank; if this build of torch / NCCL refuses, the graph covers forward .. backward only
                try:
                    if args.no_graph_collective:
                        raise RuntimeError("--no-graph-collective")
                    graphed = graphs.GraphedTrainStep(lambda x, y, m: eager_step(x, y, m), resident[0])
                    step = lambda x, y, m: graphed.run(x, y, m)
                    graph_scope = "fwd+criterion+bwd+allreduce+adam"
                except Exception as e2:
                    print(f"note: all-reduce not captured ({e2!r}); graph = forward..backward", file=sys.stderr)
                    torch.cuda.synchronize()

                    def fwd_bwd(x, y, m):
                        bucket.zero()
                        loss = ops.seg_loss(model(x), y, m)
                        loss.backward()
                        return loss
                    graphed = graphs.GraphedTrainStep(fwd_bwd, resident[0])
                    graph_scope = "fwd+criterion+bwd"

                    def step(x, y, m):
                        loss = graphed.run(x, y, m)
                        bucket.all_reduce_mean()
                        opt.step()
                        return loss
        except Exception as e:               # capture problems must not cost the bench line: fall back to eager launches
            print(f"warning: CUDA graph capture failed ({e!r}); eager launches", file=sys.stderr)
            torch.cuda.synchronize()
            step, graphed = eager_step, None
        for i in range(3):
            step(*resident[i % nbuf])

    # ---- device-resident throughput: no per-call instrumentation inside this region ---------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(lambda i: step(*resident[i % nbuf]), args.steps)
    clocks = sampler.stop()
    points = B_PER_GPU * NPTS * world * args.steps
    value = points / (ms * 1e-3)

    # ---- end-to-end through the public API with host buffers ------------------------------------------
    # Every step copies ONE batch from pinned host memory and reads its loss back.  The copy of batch i+1 is issued on a
    # copy stream before the loss of step i is read (double buffering, what a DataLoader with pin_memory does), so it
    # overlaps the tail of step i instead of serialising with the read-back.
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}

    def stage(i):
        with torch.cuda.stream(copy_stream):
            xh, yh, mh = host[i % nbuf]
            t = (xh.to(dev, non_blocking=True), yh.to(dev, non_blocking=True), mh.to(dev, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        staged[i] = (t, ev)

    def e2e_step(i):
        if i not in staged:
            stage(i)
        (x, y, m), ev = staged.pop(i)
        torch.cuda.current_stream(dev).wait_event(ev)
        for t in (x, y, m):
            t.record_stream(torch.cuda.current_stream(dev))
        loss = step(x, y, m)                 # (graph replay: device-to-device copy into the static inputs, then one launch)
        stage(i + 1)                         # H2D of the next batch, inside this step's timed region
        return float(loss.item())            # D2H read of the step's result

    e2e_step(0)
    staged.clear()
    ms_e2e = timed(e2e_step, args.steps)
    staged.clear()
    e2e_value = points / (ms_e2e * 1e-3)

    if rank != 0:
        return
    peaks = load_peaks()
    sm_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
    breakdown = {}
    for name, (calls, tot) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
        breakdown[name] = {"calls_per_step": calls / args.steps, "ms_per_step": tot / args.steps}
    roof = roofline_of(per_kernel, args.steps, B_PER_GPU, NPTS, peaks, sm_mhz)

    # ---- extras: kNN + EdgeConv per layer, the FP32-accurate gradient setting, eager oracle on this GPU ----
    extras = None
    if world == 1 and not args.no_extras:
        extras = {"layers_us": layer_times(pkg, dev)}
        extras["grad_nsplit3"] = {"ms_per_step": ms3 / args.steps, "value": points / (ms3 * 1e-3), "unit": UNIT,
                                  "what": "the eagerly launched step with every gradient GEMM on 3 bf16 planes (FP32-accurate, 6 MMAs per "
                                          "product) and the conv_b backward on the FP32-FMA kernel; compare optimizer.ms_per_step_eager "
                                          "(2 planes, 2^-17, eager launches)"}
        extras["torch_eager_gpu"] = eager_gpu_oracle(dev)

    # ---- CPU baseline (bounded sample) -------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        times, cores = cpu_reference_steps(2, 1, B_PER_GPU)
        cpu = {"value": B_PER_GPU * NPTS * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port", "cpu": cpu_model_name(),
               "sample": f"the full batch of {B_PER_GPU} blocks (N={NPTS}), 1 warm-up + 2 timed training steps of the CPU oracle"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "precision": ("FP32 I/O; forward GEMMs/Gram tiles as bf16x3 split MMAs (FP32-accurate, ~2^-24) or exact FP32 FMA re-scores; "
                      f"gradient GEMMs on {ops.GRAD_NSPLIT} bf16 planes ({'2^-17 relative' if ops.GRAD_NSPLIT == 2 else 'FP32-accurate'}), "
                      "see extras.grad_nsplit3 for the all-FP32-accurate step"),
        "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "cuda_graph": graphed is not None, "cuda_graph_scope": graph_scope if graphed is not None else None,
        "collective": None if world == 1 else ("all-reduce of conv6..conv9 (97 % of the 3.92 MB bucket) started from an autograd hook while the "
                                               "EdgeConv backward runs, the rest after the backward" if overlap else
                                               "one all-reduce of the 3.92 MB bucket after the backward"),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "optimizer": {"ms_per_step": max(ms_eager - ms_noopt, 0.0) / args.steps,
                      "what": "fused Adam (train.py:118): eagerly launched step time minus the same step without opt.step()",
                      "ms_per_step_eager": ms_eager / args.steps, "ms_per_step_eager_without": ms_noopt / args.steps},
        "kernels": breakdown,
        "kernels_pass_ms_per_step": ms_instr / args.steps,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm: tile inference
INFER_METRIC = "inference points/sec (N=32768,k=30, 16 blocks/batch)"


def run_infer(args, rank: int, local: int, world: int):
    import __graft_entry__ as entry
    import torch.distributed as dist
    pkg = entry.load_package()
    ops = pkg.ops
    inference = __import__(entry.PKG_NAME + ".inference", fromlist=["x"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = pkg.DGCNN(CLASSES, FEAT, KNN, dropout=0.5, emb_dims=EMB, cuda=True).eval()
    nb = INFER_B * world                                           # one batch per GPU and step; the tile's blocks are shared by all ranks
    host_sets = [make_blocks(10 + i, nb, INFER_N).pin_memory() for i in range(2)]
    dev_sets = [h.to(dev) for h in host_sets]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(i)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def resident_step(i):
        blocks = dev_sets[i % 2][rank::world]                      # this rank's shard, already in HBM
        x = inference.shift_to_block_origin(blocks.permute(0, 2, 1)).float()
        with torch.no_grad():
            return ops.seg_predict(model(x), 0.8)

    def e2e_step(i):
        return inference.predict_blocks(model, host_sets[i % 2], batch_size=INFER_B, min_class_confidence=0.8, rank=rank, world=world)

    for i in range(max(args.warmup, 3)):
        resident_step(i)
    sampler = ClockSampler(local)
    sampler.start()
    ops.LOG.reset(timing=False)
    ms = timed(resident_step, args.steps)
    launches = ops.LOG.launches
    clocks = sampler.stop()
    points = INFER_B * INFER_N * world * args.steps
    e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps)
    ops.LOG.reset(timing=True)
    timed(resident_step, args.steps)
    per_kernel = ops.LOG.summary()
    ops.LOG.reset(timing=False)
    if rank != 0:
        return
    peaks = load_peaks()
    breakdown = {name: {"calls_per_step": calls / args.steps, "ms_per_step": tot / args.steps}
                 for name, (calls, tot) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1])}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        times, cores = cpu_reference_infer(1, 1, 8192)
        cpu = {"value": 8192 * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port", "cpu": cpu_model_name(),
               "sample": "one block of N=8192 points (B=1, the reference's predict() loop body), 1 warm-up + 1 timed forward"}
    line = {
        "metric": INFER_METRIC, "value": points / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(world, "infer"),
        "e2e": {"value": points / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": INFER_B * INFER_N * FEAT * 4,
                "d2h_bytes_per_step": INFER_B * INFER_N * 8, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": roofline_of(per_kernel, args.steps, INFER_B, INFER_N, peaks, clocks["sm_mhz"] or 1965.0),
        "cpu_baseline": cpu, "kernels": breakdown,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "infer"],
                    help="train = BASELINE configs[1]/[2] (the headline metric); infer = configs[3], sharded tile inference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--overlap-collective", action="store_true",
                    help="N > 1: exchange the head's weight gradients from an autograd hook while the EdgeConv backward runs "
                         "(GradBucket.overlap_from; measured: no gain, see DESIGN.md §6) instead of one all-reduce after the backward")
    ap.add_argument("--no-graph-collective", action="store_true", help="N > 1: keep the gradient all-reduce and Adam outside the graph")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-layer timings, the nsplit=3 step and the eager-GPU oracle run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import __graft_entry__ as entry
    entry.load_package()
    parallel = __import__(entry.PKG_NAME + ".parallel", fromlist=["x"])
    rank, local, world = parallel.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    (run_infer if args.workload == "infer" else run_ours)(args, rank, local, world)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
