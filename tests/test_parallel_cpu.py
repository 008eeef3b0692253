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


def _overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    entry.load_package()
    parallel = __import__(entry.PKG_NAME + ".parallel", fromlist=["x"])
    parallel.init_from_env("gloo")
    torch.manual_seed(0)
    # a chain, so that the gradients of the later layers are complete before the trigger layer's (the precondition)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8, bias=False), torch.nn.Tanh(), torch.nn.Linear(8, 8, bias=False), torch.nn.Tanh(),
                              torch.nn.Linear(8, 3, bias=False))
    x = torch.randn(16, 4, generator=torch.Generator().manual_seed(100 + rank))
    results = []
    for overlap in (False, True):
        bucket = parallel.GradBucket(net.parameters())
        if overlap:
            bucket.overlap_from(net[2].weight)
            assert bucket._tail_start == net[0].weight.numel()
        for _ in range(2):                         # two steps: the hook must re-arm
            bucket.zero()
            net(x).square().sum().backward()
            if overlap:
                assert bucket._tail_work is not None
            bucket.all_reduce_mean()
            assert bucket._tail_work is None
        results.append(bucket.flat.clone())
        if bucket._hook is not None:
            bucket._hook.remove()
    # reference: mean over ranks of the local gradients
    local = parallel.GradBucket(net.parameters())
    local.zero()
    net(x).square().sum().backward()
    gathered = [torch.empty_like(local.flat) for _ in range(world)]
    dist.all_gather(gathered, local.flat)
    mean = sum(gathered) / world
    out[rank] = bool(torch.allclose(results[0], mean, rtol=1e-6, atol=1e-7) and torch.allclose(results[1], mean, rtol=1e-6, atol=1e-7))
    dist.destroy_process_group()


def test_grad_bucket_overlapped_tail_gloo_world2():
    """The early exchange of the bucket's tail (started from an autograd hook) gives the same averaged gradients."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        out = mgr.dict()
        procs = [ctx.Process(target=_overlap_worker, args=(r, world, port, out)) for r in range(world)]
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
