"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: flat gradient bucket all-reduce, parameter
broadcast, block sharding.  The data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as entry


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    pkg = entry.load_package()
    parallel = __import__(entry.PKG_NAME + ".parallel", fromlist=["x"])
    r, _, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(rank)                      # different init per rank ...
    model = pkg.DGCNN(5, 4, 20)
    parallel.broadcast_parameters(model)         # ... made identical here
    flat0 = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    bucket = parallel.GradBucket(model.parameters())
    assert bucket.flat.numel() == 980864
    # autograd accumulates into the bucket views
    loss = sum((p * (rank + 1)).sum() for p in model.parameters())
    loss.backward()
    assert torch.all(bucket.flat == rank + 1)
    bucket.all_reduce_mean()
    expect = sum(range(1, world + 1)) / world
    ok = bool(torch.allclose(bucket.flat, torch.full_like(bucket.flat, expect)))
    # every parameter's .grad still aliases the bucket
    ok &= all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())
    gathered = [torch.empty_like(flat0) for _ in range(world)]
    dist.all_gather(gathered, flat0)
    ok &= all(torch.equal(g, gathered[0]) for g in gathered)
    bucket.zero()
    ok &= float(model.conv9.weight.grad.abs().sum()) == 0.0
    out[rank] = ok
    dist.destroy_process_group()


def test_grad_bucket_allreduce_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
        [p.start() for p in procs]
        [p.join(120) for p in procs]
        assert all(p.exitcode == 0 for p in procs)
        assert dict(out) == {0: True, 1: True}


def test_shard_blocks_partition():
    parallel = __import__(entry.PKG_NAME + ".parallel", fromlist=["x"]) if entry.load_package() else None
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in parallel.shard_blocks(37, r, world))
        assert seen == list(range(37))
        sizes = [len(parallel.shard_blocks(37, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_blocks(4, 4, 4)
