"""Multi-GPU plumbing for the hot path: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch).

The path shards by 7000-point blocks: every kernel is per block, BatchNorm statistics stay
per replica (B=8 per GPU, the reference's single-device semantics), so training has exactly one exchange step —
a sum-all-reduce of the 980,864-element gradient bucket — and tile inference has none.
The reference has no distributed code at all; this module is new.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise the default process group from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).

    Returns (rank, local_rank, world_size).  With WORLD_SIZE unset or 1 nothing is initialised."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)      # model.py's `cuda=True` uses the *current* device (reference model.py:60)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, local, world


class GradBucket:
    """All parameter gradients as views of ONE flat buffer, so the data-parallel exchange is one all-reduce.

    3.92 MB for the reference configuration: latency-bound on NVLink 5 (tens of microseconds), so one collective per
    step, launched after the last backward kernel.  `overlap_from(p)` is the opt-in alternative: the bucket's tail (97 %
    of it: the head's weights) is exchanged from an autograd hook as soon as those gradients are complete, on the process
    group's own stream, while the EdgeConv backward is still running.  Measured on 2 x B200 (tools/diag_overlap.py): 4.23
    ms per eagerly launched step either way — the exchange is too short to matter — so bench.py keeps the single
    all-reduce.  (A CUDA graph that captured a collective must be destroyed before the process group is.)
    """

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket: no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self._offsets = {}
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)      # autograd accumulates in place into the bucket
            self._offsets[id(p)] = off
            off += n
        self._tail_start = total       # [tail_start, total) is exchanged early (overlap_from)
        self._tail_work = None
        self._hook = None

    def zero(self) -> None:
        self.flat.zero_()

    @staticmethod
    def _distributed() -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def overlap_from(self, first: torch.nn.Parameter) -> None:
        """Exchange the gradients of `first` and of every parameter after it in the bucket as soon as `first`'s gradient has
        been accumulated, concurrently with the rest of the backward pass.

        Precondition (the caller's knowledge of its model): the parameters after `first` belong to layers that run after
        `first`'s layer in the forward pass, so their gradients are complete before `first`'s.  For DGCNN `first` =
        conv6's weight: conv6..conv9 are registered last and are the first to receive gradients."""
        if id(first) not in self._offsets:
            raise ValueError("overlap_from: parameter is not in the bucket")
        if self._hook is not None:
            self._hook.remove()
            self._hook = None
        self._tail_start = self._offsets[id(first)]
        if not self._distributed():
            return

        def start_tail(_param):
            # the collective runs on the process group's own stream, ordered after the work enqueued so far
            self._tail_work = dist.all_reduce(self.flat[self._tail_start:], op=dist.ReduceOp.SUM, async_op=True)

        self._hook = first.register_post_accumulate_grad_hook(start_tail)

    def all_reduce_mean(self) -> None:
        """Average the bucket across ranks (no-op without a process group)."""
        if not self._distributed():
            return
        if self._tail_work is not None:
            if self._tail_start > 0:
                dist.all_reduce(self.flat[:self._tail_start], op=dist.ReduceOp.SUM)
            self._tail_work.wait()           # the current stream waits for the early exchange
            self._tail_work = None
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.mul_(1.0 / dist.get_world_size())


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """Replicate rank `src`'s parameters and buffers (one flat broadcast per dtype)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    by_dtype: dict = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for group in by_dtype.values():
        flat = torch.cat([t.reshape(-1) for t in group])
        dist.broadcast(flat, src)
        off = 0
        for t in group:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


def shard_blocks(num_blocks: int, rank: int, world: int) -> Sequence[int]:
    """Round-robin assignment of independent point-cloud blocks (reference test.py:291 loop) to ranks.
    Inference needs no collective: each rank runs its blocks and the host concatenates predictions."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return range(rank, num_blocks, world)
