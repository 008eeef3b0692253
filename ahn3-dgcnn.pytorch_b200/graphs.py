"""CUDA-graph replay of the training step and of a fixed-shape inference forward.

The step is ~220 kernel launches of 5-500 us each; launched eagerly the GPU waits for the host after every host-side
read of the loss (train.py:203 logs `loss.item()` each step).  Capturing forward + criterion + backward (+ the optimizer
update when there is no gradient exchange in between) into ONE graph makes a step a single launch.  Everything on the
path is capture-safe: the C-ABI only enqueues work on the current stream, scratch comes from the caching allocator (kept
alive by the graph's pool), BatchNorm counters are device tensors.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import torch


class GraphedTrainStep:
    """fn(*static_inputs) -> loss, captured once; `run(*inputs)` copies the inputs into the static buffers and replays.

    `fn` must do the whole device-side step (zero the gradient bucket, forward, loss, backward, optional optimizer step
    with a capturable optimizer) and return the loss tensor."""

    def __init__(self, fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        self.static_in = [t.clone() for t in example_inputs]
        dev = self.static_in[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                    # warm-up off the default stream (allocator pools, lazy init, autotune-free)
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = fn(*self.static_in)

    def run(self, *inputs: torch.Tensor, non_blocking: bool = True) -> torch.Tensor:
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=non_blocking)
        self.graph.replay()
        return self.static_loss


class GraphedForward:
    """Eval-mode `model(x)` (optionally followed by the inference epilogue) for ONE input shape, captured once.

    A forward pass of a small block is ~60 launches of 2-60 us each; launched eagerly the host is the bottleneck (the
    README example, B=1 x 1000 points: 1.5 ms per call against 0.3 ms of device time).  `run(x)` copies x into the static
    input and replays; the returned tensors are the graph's static outputs (clone them to keep them across calls)."""

    def __init__(self, model: torch.nn.Module, example_x: torch.Tensor, post: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 warmup: int = 2):
        if model.training:
            raise RuntimeError("GraphedForward: call model.eval() first")
        self.static_x = example_x.clone()
        dev = self.static_x.device
        fn = (lambda x: model(x)) if post is None else (lambda x: post(model(x)))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                fn(self.static_x)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = fn(self.static_x)

    def run(self, x: torch.Tensor, non_blocking: bool = True) -> torch.Tensor:
        if x.shape != self.static_x.shape:
            raise RuntimeError(f"GraphedForward: captured for {tuple(self.static_x.shape)}, got {tuple(x.shape)}")
        if x is not self.static_x:
            self.static_x.copy_(x, non_blocking=non_blocking)
        self.graph.replay()
        return self.static_out
