"""Batched, sharded tile inference — the `predict()` path of the reference (test.py:281-330) without its per-block
host work.

The reference classifies a tile block by block (B=1): `np.savetxt` of every block, a host-side shift of x/y to the
block's lower-left corner (test.py:293-296), one forward pass, softmax/max/threshold on the host tensor (test.py:303-306),
`np.vstack` and a LAS file per block.  Blocks are independent, so here they are
  * sharded round-robin over the ranks (`parallel.shard_blocks`, no collective — SURVEY.md §8e),
  * batched `batch_size` blocks per forward (cfg4: 16 blocks of 32768 points),
  * shifted, classified and thresholded on the device (`ops.seg_predict`, UNCLASSIFIED = 31),
and only the int64 label rows travel back to the host, in block order.  File formats stay with the caller.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops
from .parallel import shard_blocks


def shift_to_block_origin(x: torch.Tensor) -> torch.Tensor:
    """test.py:293-296: `X -= [min x, min y, 0, 0]` per block, for a batch [B, F, N] (returns a new tensor)."""
    x = x.clone()
    x[:, :2, :] -= x[:, :2, :].amin(dim=2, keepdim=True)
    return x


@torch.no_grad()
def predict_blocks(model: torch.nn.Module, blocks: torch.Tensor, batch_size: int = 16, min_class_confidence: float = 0.8,
                   rank: int = 0, world: int = 1, shift_origin: bool = True,
                   device: Optional[torch.device] = None, return_index: bool = False):
    """Classify this rank's share of `blocks`.

    blocks: [num_blocks, N, F] (the reference's block layout, test.py:276-299; float32/float64, host — pinned for
    asynchronous copies — or device).  Returns int64 [n_mine, N] labels on the host in the order of
    `shard_blocks(num_blocks, rank, world)` (and that index list if return_index).  `model` must be in eval mode.
    """
    if model.training:
        raise RuntimeError("predict_blocks: call model.eval() first (test.py:267)")
    if blocks.dim() != 3:
        raise RuntimeError(f"predict_blocks: expected blocks [num_blocks, N, F], got {tuple(blocks.shape)}")
    dev = device or next(model.parameters()).device
    mine: Sequence[int] = list(shard_blocks(blocks.size(0), rank, world))
    N = blocks.size(1)
    out = torch.empty(len(mine), N, dtype=torch.int64, pin_memory=True)
    # two pinned staging buffers: the gather of batch i+1 on the host overlaps the copy / compute of batch i
    stage = None if blocks.is_cuda else [torch.empty(batch_size, N, blocks.size(2), dtype=blocks.dtype, pin_memory=True) for _ in range(2)]
    staged = [None, None]
    for bi, s in enumerate(range(0, len(mine), batch_size)):
        sel = mine[s:s + batch_size]
        index = torch.as_tensor(sel, dtype=torch.long, device=blocks.device)
        if blocks.is_cuda:
            xb = blocks.index_select(0, index)
        else:
            slot = bi & 1
            if staged[slot] is not None:
                staged[slot].synchronize()                                     # the copy that last used this buffer is done
            torch.index_select(blocks, 0, index, out=stage[slot][:len(sel)])
            xb = stage[slot][:len(sel)].to(dev, non_blocking=True)
            staged[slot] = torch.cuda.Event()
            staged[slot].record(torch.cuda.current_stream(dev))
        xb = xb.permute(0, 2, 1)                                               # test.py:299-301: [B, F, N] view of [B, N, F]
        if shift_origin:
            xb = shift_to_block_origin(xb)                                     # in the blocks' own dtype (float64 upstream), THEN .float()
        xb = xb.float()
        logits = model(xb)                                                     # test.py:302
        pred = ops.seg_predict(logits, min_class_confidence)                   # test.py:303-306
        out[s:s + len(sel)].copy_(pred, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return (out, mine) if return_index else out
