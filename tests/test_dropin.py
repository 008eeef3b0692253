"""The drop-in boundary: the package must be usable EXACTLY the way the reference's drivers use
`model.py` — `sys.path` points at the directory and they say `from model import DGCNN` (train.py:12, test.py:11) —
and their call sequences must run unchanged.  The GPU tests below replay those sequences line by line."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import common
import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "ahn3-dgcnn.pytorch_b200")
DEV = "cuda:0"


def _run_fresh(code: str) -> str:
    """A fresh interpreter whose only knowledge of the repo is `sys.path.insert(0, <package dir>)`."""
    env = dict(os.environ, PYTHONPATH="")
    res = subprocess.run([sys.executable, "-c", f"import sys; sys.path.insert(0, {PKG_DIR!r})\n" + code],
                         capture_output=True, text=True, cwd="/tmp", env=env, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    return res.stdout


def test_import_site_of_the_reference_drivers_works():
    entry.build()
    out = _run_fresh(
        "from model import DGCNN\n"                      # train.py:12 / test.py:11
        "from model import knn, get_graph_feature\n"
        "import model\n"
        "m = DGCNN(5, 4, 20, ropout = 0.5, emb_dims = 1024, cuda = False)\n"      # train.py:102-105 (incl. the `ropout` typo)
        "print(len(m.state_dict()), sum(p.numel() for p in m.parameters()), model.DGCNN is DGCNN, str(m).count('Conv'))\n")
    assert out.split() == ["89", "980864", "True", "9"]


def test_checkpoint_written_like_train_py_loads_strictly(tmp_path):
    """train.py:301-312 saves {'epoch', 'mBPCA', 'model_state_dict', ...}; test.py:93-100 loads it with strict keys."""
    pkg = entry.load_package()
    torch.manual_seed(3)
    a = pkg.DGCNN(5, 4, 20, dropout=0.5, emb_dims=1024, cuda=False)
    opt = torch.optim.Adam(a.parameters(), lr=1e-3, weight_decay=1e-4)               # train.py:118
    state = {"epoch": 3, "mBPCA": 0.5, "model_state_dict": a.state_dict(), "optimizer_state_dict": opt.state_dict()}
    path = os.path.join(tmp_path, "dgcnn_model.t7")
    torch.save(state, path)
    b = pkg.DGCNN(5, 4, 20, dropout=0.5, emb_dims=1024, cuda=False)
    checkpoint = torch.load(path, map_location=torch.device("cpu"))                    # test.py:96
    b.load_state_dict(checkpoint["model_state_dict"])                                  # test.py:99 (strict)
    b = b.eval()                                                                       # test.py:100
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    # the reference's own state_dict surface (names recorded from /root/reference/model.py in the golden file)
    g = np.load(os.path.join(ROOT, "tests", "golden", "model.npz"))
    assert sorted(a.state_dict().keys()) == [str(s) for s in g["eval_small/state_keys"]]
    assert [n for n, _ in a.named_parameters()] == [str(s) for s in g["eval_small/param_order"]]


# ------------------------------------------------------------------------------------------------ GPU: call sites, verbatim
@pytest.mark.gpu
def test_train_py_call_sequence():
    """train.py:102-118,159,175-179 then the criterion of :180-197 (device form) and :198-201."""
    pkg = entry.load_package()
    from model import DGCNN as TopLevelDGCNN                     # noqa: F401  (sys.path contains the package dir, see conftest)
    torch.cuda.set_device(0)
    num_classes, num_features, k, dropout, emb_dims, cuda = 5, 4, 20, 0.5, 1024, True
    device = torch.device("cuda" if cuda else "cpu")                                   # train.py:99
    model = TopLevelDGCNN(num_classes, num_features, k,
                          ropout = dropout,
                          emb_dims = emb_dims,
                          cuda = cuda)                                                 # train.py:102-105
    assert "DGCNN(" in str(model) and next(model.parameters()).is_cuda                 # train.py:107
    opt = torch.optim.Adam(model.parameters(), lr=0.001, weight_decay=1e-4)            # train.py:118
    model.train()                                                                      # train.py:159
    # the loader's collate_pcs returns float64 [B, N, F] points and float64 [B, N] labels (data.py:169-182)
    rs = np.random.RandomState(0)
    data = torch.tensor(np.concatenate([rs.rand(4, 600, 2) * 30, 100 + rs.rand(4, 600, 1) * 30, rs.rand(4, 600, 1) * 20], 2))
    labels = torch.tensor(rs.randint(0, num_classes, size=(4, 600)).astype(np.float64))
    mask = torch.tensor((rs.rand(4, 600) < 0.5).astype(np.float64))
    assert data.dtype == torch.float64
    data, labels, mask = data.to(device), labels.to(device), mask.to(device)           # train.py:175
    data = data.permute(0, 2, 1).float()                                               # train.py:176
    assert not data.is_contiguous()
    batch_size = data.size()[0]                                                        # train.py:177
    opt.zero_grad()                                                                    # train.py:178
    labels_pred = model(data)                                                          # train.py:179
    assert tuple(labels_pred.shape) == (batch_size, num_classes, 600) and labels_pred.is_contiguous()
    loss, pred = pkg.ops.seg_loss(labels_pred, labels, mask, return_pred=True)         # train.py:180-197, 201 on device
    loss.backward()                                                                    # train.py:198
    opt.step()                                                                         # train.py:199
    assert torch.isfinite(loss) and tuple(pred.shape) == (batch_size, 600)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    # validation leg: train.py:256,264-271
    model.eval()
    with torch.no_grad():
        out = model(data)
    vloss, vpred = pkg.ops.seg_loss(out, labels, return_pred=True)
    assert torch.isfinite(vloss) and int(vpred.max()) < num_classes


@pytest.mark.gpu
def test_test_py_call_sequences(tmp_path):
    """test.py:91-100 (strict checkpoint load), :118-127 (batched evaluation) and :298-308 (per-block predict with the
    [N,F,1] -> permute(2,1,0) view whose batch stride is what numpy left behind)."""
    pkg = entry.load_package()
    from model import DGCNN as TopLevelDGCNN                     # noqa: F401
    from oracle import dgcnn_oracle as O
    from oracle import drivers_oracle as D
    torch.cuda.set_device(0)
    num_classes, num_features, k, dropout, emb_dims, cuda = 5, 4, 20, 0.5, 1024, True
    w = common.weights(71, num_classes, num_features, emb_dims)
    src = pkg.DGCNN(num_classes, num_features, k, dropout=dropout, emb_dims=emb_dims, cuda=False)
    src.load_state_dict(w, strict=False)
    path = os.path.join(tmp_path, "dgcnn_model.t7")
    torch.save({"epoch": 0, "mBPCA": 0.0, "model_state_dict": src.state_dict()}, path)  # train.py:301-312

    model = TopLevelDGCNN(num_classes, num_features, k, dropout = dropout, emb_dims = emb_dims, cuda = cuda)   # test.py:91
    checkpoint = torch.load(path)                                                      # test.py:94
    model.load_state_dict(checkpoint['model_state_dict'])                              # test.py:99
    model = model.eval()                                                               # test.py:100
    device = torch.device("cuda")

    # ---- batched evaluation loop body, test.py:118-127
    data = torch.tensor(common.points(72, 3, num_features, 500, "real").permute(0, 2, 1).numpy().astype(np.float64))   # [B,N,F] fp64
    data = data.to(device)                                                             # test.py:118
    data = data.permute(0, 2, 1).float()                                               # test.py:119
    seg_pred = model(data)                                                             # test.py:122
    pred = pkg.ops.seg_predict(seg_pred, 0.8)                                          # test.py:123-127 on device
    ora = O.OracleDGCNN({kk: v.clone() for kk, v in w.items()}, k, training=False)
    ref_logits = ora(data.cpu())
    assert ((seg_pred.cpu() - ref_logits).abs().max() / ref_logits.abs().max()).item() < 1e-3
    ref_pred, vals = D.predict_labels(seg_pred.cpu(), 0.8)
    assert not ((pred.cpu() != ref_pred) & ((vals - 0.8).abs() > 1e-6)).any()

    # ---- predict(): one block at a time, test.py:298-308
    X = common.points(73, 1, num_features, 400, "real")[0].t().contiguous().numpy().astype(np.float64)     # [N,F] block
    X = X[:, :, np.newaxis]                                                            # test.py:299
    X = torch.tensor(X)                                                                # test.py:300
    X = X.permute(2, 1, 0).float()                                                     # test.py:301  -> [1,F,N]
    logit_pred = model(X.to(device))                                                   # test.py:302 (+ the .to a GPU run needs)
    assert tuple(logit_pred.shape) == (1, num_classes, 400)
    ref = O.OracleDGCNN({kk: v.clone() for kk, v in w.items()}, k, training=False)(X)
    assert ((logit_pred.cpu() - ref).abs().max() / ref.abs().max()).item() < 1e-3
    pred1 = pkg.ops.seg_predict(logit_pred, 0.8)                                       # test.py:303-306
    assert tuple(pred1.shape) == (1, 400) and pred1.dtype == torch.int64


@pytest.mark.gpu
def test_model_on_non_current_device_guard():
    """launches must follow the tensors' device, not the caller's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    pkg = entry.load_package()
    torch.cuda.set_device(0)
    m = pkg.DGCNN(5, 4, 8).to("cuda:1").eval()
    x = common.points(5, 1, 4, 300, "unit")
    y1 = m(x.to("cuda:1"))
    assert y1.device.index == 1
    torch.cuda.set_device(1)
    y2 = m(x.to("cuda:1"))
    torch.cuda.set_device(0)
    assert torch.equal(y1, y2)


@pytest.mark.gpu
def test_input_gradient_and_small_blocks():
    """the reference is differentiable w.r.t. the input cloud and works for any N >= k."""
    pkg = entry.load_package()
    torch.cuda.set_device(0)
    m = pkg.DGCNN(5, 4, 6, dropout=0.0, cuda=True).train()
    x = common.points(9, 2, 4, 20, "unit").to(DEV).requires_grad_(True)               # N = 20 < 32: fallback global feature
    m(x).square().mean().backward()
    assert x.grad is not None and tuple(x.grad.shape) == (2, 4, 20) and torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0
    m2 = pkg.DGCNN(5, 3, 6, dropout=0.0, cuda=True).train()                           # F = 3: rows padded to 4 channels
    x3 = common.points(10, 2, 3, 64, "unit").to(DEV).requires_grad_(True)
    m2(x3).square().mean().backward()
    assert tuple(x3.grad.shape) == (2, 3, 64) and torch.isfinite(x3.grad).all()
    m.bn7.momentum = None                                                             # cumulative moving average is legal
    m(common.points(11, 2, 4, 40, "unit").to(DEV))


@pytest.mark.gpu
def test_no_library_gemm_on_the_model_path():
    """Round-1 VERDICT item 6: cfg1 (B=1, N=1000, k=20) inference and a training step must not launch a single
    cuBLAS / CUTLASS / cuDNN kernel — every GEMM-shaped op is one of this library's tcgen05 kernels."""
    pkg = entry.load_package()
    try:
        from torch.profiler import ProfilerActivity, profile
    except Exception:                                   # pragma: no cover
        pytest.skip("torch.profiler unavailable")
    torch.cuda.set_device(0)
    x = torch.rand(1, 4, 1000, device=DEV)                                      # README example (cfg1)
    net = pkg.DGCNN(5, 4, 20, dropout=0.5, emb_dims=1024, cuda=True).eval()
    net(x)
    labels = torch.randint(0, 5, (2, 600), device=DEV)
    xt = common.points(3, 2, 4, 600, "real").to(DEV)
    tr = pkg.DGCNN(5, 4, 20, dropout=0.5, emb_dims=1024, cuda=True).train()
    pkg.ops.seg_loss(tr(xt), labels).backward()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net(x)
        pkg.ops.seg_loss(tr(xt), labels).backward()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    if not names:
        pytest.skip("no CUDA activity records (CUPTI unavailable)")
    bad = [n for n in names if any(t in n.lower() for t in ("cutlass", "cublas", "gemm", "gemv", "cudnn", "xmma"))
           and not n.startswith(("dg::", "void dg::"))]
    assert not bad, bad
    assert any("tc_gemm_kernel" in n for n in names) and any("knn_tc_scan_kernel" in n for n in names)
