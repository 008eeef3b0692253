"""CUDA-graph replay of the training step (graphs.GraphedTrainStep): the replayed step must do what the eagerly launched
step does — same losses step after step, parameters moving, BatchNorm counters advancing."""
import pytest
import torch

import __graft_entry__ as entry
from golden import common

DEV = "cuda:0"


def _setup(pkg, seed):
    torch.manual_seed(seed)
    m = pkg.DGCNN(5, 4, 8, dropout=0.0, cuda=True).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
    return m, opt


@pytest.mark.gpu
def test_graphed_step_matches_eager_steps():
    pkg = entry.load_package()
    from importlib import import_module
    graphs = import_module(entry.PKG_NAME + ".graphs")
    ops = pkg.ops
    torch.cuda.set_device(0)
    B, N = 2, 600
    batches = []
    for s in range(4):
        x = common.points(20 + s, B, 4, N, "unit").to(DEV)
        g = torch.Generator().manual_seed(s)
        y = torch.randint(0, 5, (B, N), generator=g).to(DEV)
        mask = (torch.rand(B, N, generator=g) < 0.8).to(DEV)
        batches.append((x, y, mask))

    def make_step(m, opt):
        def step(x, y, mask):
            opt.zero_grad(set_to_none=False)
            loss = ops.seg_loss(m(x), y, mask)
            loss.backward()
            opt.step()
            return loss
        return step

    # eager arm: 3 warm-up steps on batch 0 (what the capture's warm-up does) + one per batch
    m1, o1 = _setup(pkg, 7)
    for p in m1.parameters():
        p.grad = torch.zeros_like(p)
    s1 = make_step(m1, o1)
    for _ in range(4):                                   # 3 warm-up + the captured call itself
        s1(*batches[0])
    eager = [float(s1(*b)) for b in batches]

    m2, o2 = _setup(pkg, 7)
    for p in m2.parameters():
        p.grad = torch.zeros_like(p)
    g = graphs.GraphedTrainStep(make_step(m2, o2), batches[0], warmup=3)   # capture itself does not execute; replay does
    g.run(*batches[0])
    replay = [float(g.run(*b)) for b in batches]
    for a, b in zip(eager, replay):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(a)), (eager, replay)
    assert len(set(round(v, 6) for v in replay)) > 1                    # the graph sees the new inputs, the weights move
    bn = [mod for mod in m2.modules() if hasattr(mod, "num_batches_tracked")]
    if bn:
        assert int(bn[0].num_batches_tracked) == int([mod for mod in m1.modules() if hasattr(mod, "num_batches_tracked")][0].num_batches_tracked)


@pytest.mark.gpu
def test_graphed_forward_matches_eager_forward():
    """graphs.GraphedForward: the replayed eval forward (+ epilogue) equals the eagerly launched one, call after call."""
    pkg = entry.load_package()
    from importlib import import_module
    graphs = import_module(entry.PKG_NAME + ".graphs")
    torch.cuda.set_device(0)
    torch.manual_seed(3)
    m = pkg.DGCNN(5, 4, 20, cuda=True).eval()
    xs = [common.points(40 + s, 1, 4, 1000, "unit").to(DEV) for s in range(3)]        # the README shape (cfg1)
    g = graphs.GraphedForward(m, xs[0])
    for x in xs:
        want = m(x)
        got = g.run(x).clone()
        assert got.shape == want.shape == (1, 5, 1000)
        assert float((got - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
    gp = graphs.GraphedForward(m, xs[0], post=lambda z: pkg.ops.seg_predict(z, 0.3))
    assert torch.equal(gp.run(xs[1]), pkg.ops.seg_predict(m(xs[1]), 0.3))
    with pytest.raises(RuntimeError):
        g.run(torch.zeros(1, 4, 999, device=DEV))
    with pytest.raises(RuntimeError):
        graphs.GraphedForward(m.train(), xs[0])
