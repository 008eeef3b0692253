#!/usr/bin/env python
"""bench.py — headline benchmark of the DGCNN EdgeConv hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]/[2]): AHN3 training shape, B=8 blocks per GPU of N=7000 points, F=4, k=30,
5 classes, emb 1024.  One step = forward + loss (softmax -> label-smoothed CE, reference train.py:180-197 /
util.py:17-34) + backward + gradient all-reduce (N>1) + Adam update (reference train.py:118).  Synthetic
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
    # the reference's loaders produce [B,N,F]; the model sees the permute(0,2,1) view (train.py:176)
    return x.permute(0, 2, 1).contiguous().permute(0, 2, 1), y


def seg_loss(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """The reference's training loss: label-smoothed CE (eps 0.2) applied to softmax(logits) (train.py:180,197)."""
    p = torch.softmax(logits, dim=1).permute(0, 2, 1).reshape(-1, CLASSES)
    gold = labels.reshape(-1)
    one_hot = torch.zeros_like(p).scatter(1, gold.view(-1, 1), 1)
    one_hot = one_hot * 0.8 + (1 - one_hot) * 0.2 / (CLASSES - 1)
    return -(one_hot * torch.log_softmax(p, dim=1)).sum(dim=1).mean()


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
def cpu_reference_steps(steps: int, warmup: int, sample_blocks: int):
    """Time the CPU oracle's training step (fwd + loss + bwd + Adam) on `sample_blocks` blocks of the workload."""
    from oracle import dgcnn_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    params = O.init_params(CLASSES, FEAT, EMB, seed=0)
    leaves = []
    for name, v in params.items():
        if v.is_floating_point() and "running" not in name:
            v.requires_grad_(True)
            leaves.append(v)
    net = O.OracleDGCNN(params, KNN, dropout=0.5, training=True)
    opt = torch.optim.Adam(leaves, lr=1e-3, weight_decay=1e-4)
    x, y = make_batch(0, sample_blocks)
    xin = x
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = seg_loss(net(xin), y)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def workload_config(world: int) -> dict:
    """`config` of the JSON line: identical for our arm and for the reference arm."""
    return {"workload": f"AHN3 train step (fwd+loss+bwd+allreduce+Adam) B={B_PER_GPU}/GPU F={FEAT} N={NPTS} k={KNN} classes={CLASSES} emb={EMB}",
            "parallelism": f"dp{world}", "global_batch": B_PER_GPU * world,
            "l2": "no explicit flush: each step streams >1.5 GB of per-edge intermediates (L2 is 126 MB); 4 rotating input batches"}


def run_reference(args, rank: int):
    if rank != 0:
        return
    sample_blocks = 2
    times, cores = cpu_reference_steps(args.steps, args.warmup, sample_blocks)
    total = sum(times)
    value = sample_blocks * NPTS * len(times) / total
    sample = f"{sample_blocks} of {B_PER_GPU} blocks (N={NPTS}) per step, {args.warmup} warm-up + {args.steps} timed steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ roofline bookkeeping
def kernel_flops(name: str, calls_per_step: int, B: int):
    """Algorithmic FLOPs of ALL launches of one entry point in one step."""
    M = B * NPTS
    E = M * KNN
    if name == "dgcnn_knn_f32":                       # xyz/agl graph (C=4): 2*B*N^2*C
        return 2.0 * B * NPTS * NPTS * FEAT
    if name == "dgcnn_knn_tc_f32":                    # two feature-space graphs (C=64)
        return 2.0 * B * NPTS * NPTS * (64 + 64)
    if name == "dgcnn_edgeconv_conv_a_train":         # conv1 (2F->64), conv3, conv5 (128->64)
        return 2.0 * E * 64 * (2 * FEAT + 128 + 128)
    if name == "dgcnn_edgeconv_conv_b_train":         # conv2, conv4 (64->64)
        return 2.0 * E * 64 * 64 * 2
    if name == "dgcnn_edgeconv_bwd_b":                # per block: dW_b and dh GEMMs (recompute of y_b not counted)
        return 2.0 * E * 64 * 64 * 2 * 2
    if name == "dgcnn_tc_gemm_f32":                   # conv6/7/8 forward + dgrad + wgrad, and the EdgeConv point-level GEMMs
        head = 2.0 * M * (192 * EMB + 192 * 512 + 512 * 256) * 3
        edge = 2.0 * M * 128 * (64 + 64) * 2 + 2.0 * M * 128 * FEAT
        return head + edge
    if name == "dgcnn_edgeconv_fwd_eval":
        return 2.0 * E * 64 * (2 * FEAT + 64 + 128 + 64 + 128)
    return 0.0


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
    """kNN + EdgeConv block (train mode) per layer at the bench shape: forward and forward+backward, microseconds."""
    ops = pkg.ops
    g = torch.Generator().manual_seed(1)
    B, N, k = B_PER_GPU, NPTS, KNN
    M = B * N
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
            idx = ops.knn_rows(r.detach(), B, Cn, N, k)
            return r, ops.EdgeConvTrain.apply(r, idx, (B, Cn, N, k), wa, ga, be, wb, ga if two else None, be if two else None,
                                              ops.BNState(None, None), ops.BNState(None, None) if two else None)

        def fwd_bwd():
            r, o = fwd(True)
            o.sum().backward()

        for _ in range(2):
            fwd_bwd()
        out[name] = {"fwd": round(1e3 * _event_ms(lambda: fwd(), 5), 1), "fwd_bwd": round(1e3 * _event_ms(fwd_bwd, 5), 1)}
    return out


def eager_gpu_oracle(dev):
    """The CPU oracle's torch code run eagerly on this GPU (FP32, TF32 off): the library composition the kernels replace."""
    from oracle import dgcnn_oracle as O
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
        x, y = make_batch(0, B_PER_GPU)
        x, y = x.to(dev), y.to(dev)

        def step():
            opt.zero_grad()
            seg_loss(net(x), y).backward()
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


# ------------------------------------------------------------------------------------------------ our arm
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
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)

    nbuf = 4
    host = [make_batch(1000 * rank + i, B_PER_GPU) for i in range(nbuf)]
    host = [(x.pin_memory(), y.pin_memory()) for x, y in host]
    resident = [(x.to(dev), y.to(dev)) for x, y in host]
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 8

    def step(x, y):
        bucket.zero()
        # callers hand the model the [B,N,F] -> permute(0,2,1) view (train.py:176)
        loss = seg_loss(model(x), y)
        loss.backward()
        bucket.all_reduce_mean()
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
    # ---- device-resident throughput --------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    ops.LOG.reset(timing=True)
    ms = timed(lambda i: step(*resident[i % nbuf]), args.steps)
    launches = ops.LOG.launches
    per_kernel = ops.LOG.summary()
    clocks = sampler.stop()
    ops.LOG.reset(timing=False)
    points = B_PER_GPU * NPTS * world * args.steps
    value = points / (ms * 1e-3)

    # ---- end-to-end through the public API with host buffers ------------------------------------------
    def e2e_step(i):
        xh, yh = host[i % nbuf]
        x = xh.to(dev, non_blocking=True)
        y = yh.to(dev, non_blocking=True)
        loss = step(x, y)
        return float(loss.item())            # D2H read of the step's result

    e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = points / (ms_e2e * 1e-3)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel -----------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # all of our GEMM-shaped kernels run on the FP32 FMA pipe in this round: peak = 148 SM x 128 lanes x 2 x clock
    sm_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    dom, dom_ms = None, 0.0
    breakdown = {}
    for name, (calls, tot) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
        breakdown[name] = {"calls_per_step": calls / args.steps, "ms_per_step": tot / args.steps}
        if kernel_flops(name, 0, B_PER_GPU) > 0 and tot > dom_ms:
            dom, dom_ms = name, tot
    roof = None
    if dom:
        fl = kernel_flops(dom, 0, B_PER_GPU)
        achieved = fl / (dom_ms / args.steps * 1e-3) / 1e12
        on_tensor = dom in ("dgcnn_tc_gemm_f32", "dgcnn_knn_tc_f32", "dgcnn_edgeconv_bwd_b", "dgcnn_edgeconv_conv_b_train")
        tensor_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        peak = tensor_peak if on_tensor else fp32_peak
        kind = ("measured sustained dense bf16 tensor peak (MEASURED_PEAKS.json); the kernel issues 3 (bf16x2) or 6 (bf16x3) "
                "kind::f16 MMAs per algorithmic product for FP32-level accuracy, so the ceiling for ALGORITHMIC flops is "
                "peak/3 .. peak/6" if on_tensor else
                f"FP32 FMA pipe at the median SM clock under load ({sm_mhz:.0f} MHz); kernel is FFMA-based")
        # measured DRAM bytes per launch of that entry point from the committed `ncu --set full` capture (not a live number)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(dom, {}).get("bytes_per_launch")
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "peak_kind": kind,
                "launches_per_step": per_kernel[dom][0] / args.steps, "ms_per_step": dom_ms / args.steps}

    # ---- extras: kNN + EdgeConv per layer, and the oracle run eagerly on this GPU -----------------
    extras = None
    if world == 1 and not args.no_extras:
        extras = {"layers_us": layer_times(pkg, dev), "torch_eager_gpu": eager_gpu_oracle(dev)}

    # ---- CPU baseline (bounded sample) -------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sb = 2
        times, cores = cpu_reference_steps(1, 1, sb)
        cpu = {"value": sb * NPTS * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{sb} of {B_PER_GPU} blocks (N={NPTS}), 1 warm-up + 1 timed training step of the CPU oracle"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "kernels": breakdown,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the per-layer timings and the eager-GPU oracle run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import __graft_entry__ as entry
    pkg = entry.load_package()
    parallel = __import__(entry.PKG_NAME + ".parallel", fromlist=["x"])
    rank, local, world = parallel.init_from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    run_ours(args, rank, local, world)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
