"""Generate tests/golden/*.npz by running the REFERENCE's own model.py (build container only).

Usage (here, not on the GPU box):  python tests/golden/make_golden.py
Imports /root/reference/model.py unmodified.  The only shim: `DGCNN.forward` passes a stray
`dim9=True` to `get_graph_feature` (model.py:147 vs :43), so the module-level function is wrapped to
drop that kwarg — the arithmetic is untouched (SURVEY.md §0 finding 1).  The wrapper also records the
neighbour tables so per-layer parity can be checked.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")
import common  # noqa: E402
import model as ref  # noqa: E402  (the reference)

torch.set_num_threads(8)
_orig_ggf = ref.get_graph_feature
_orig_knn = ref.knn
_seen_idx = []


def _knn_rec(x, k):
    idx = _orig_knn(x, k)
    _seen_idx.append(idx.clone())
    return idx


def _ggf(x, k=20, use_cuda=True, **_ignored):
    return _orig_ggf(x, k=k, use_cuda=use_cuda)


ref.knn = _knn_rec
ref.get_graph_feature = _ggf

BIG = 20000


def _store_grad(out, name, g):
    g = g.detach().reshape(-1)
    out[f"gradnorm/{name}"] = np.array(float(g.double().norm()))
    out[f"grad/{name}"] = (g if g.numel() <= BIG else g[::37]).numpy()


def main():
    # ---- knn -------------------------------------------------------------------------------
    out = {}
    for name, B, C, N, k, kind, seed in common.KNN_CASES:
        x = common.points(seed, B, C, N, kind)
        out[f"{name}/idx"] = _orig_knn(x, k).numpy().astype(np.int32)
        # the model hands knn a permuted view (train.py:176); record that too (must be identical)
        xv = x.permute(0, 2, 1).contiguous().permute(0, 2, 1)
        assert torch.equal(_orig_knn(xv, k), torch.from_numpy(out[f"{name}/idx"]).long())
    np.savez_compressed(os.path.join(HERE, "knn.npz"), **out)

    # ---- get_graph_feature -----------------------------------------------------------------
    out = {}
    for name, B, C, N, k, kind, seed in common.GRAPH_CASES:
        x = common.points(seed, B, C, N, kind)
        f = _orig_ggf(x, k=k, use_cuda=False)
        out[f"{name}/feature"] = f.contiguous().numpy()
        out[f"{name}/strides"] = np.array(f.stride())
    np.savez_compressed(os.path.join(HERE, "graph_feature.npz"), **out)

    # ---- whole model -----------------------------------------------------------------------
    out = {}
    for name, B, Fdim, N, k, classes, emb, kind, seed, train in common.MODEL_CASES:
        x = common.points(seed, B, Fdim, N, kind)
        w = common.weights(seed + 100, classes, Fdim, emb)
        m = ref.DGCNN(classes, Fdim, k, dropout=0.0 if train else 0.5, emb_dims=emb, cuda=False)
        missing = m.load_state_dict(w, strict=False)
        assert not missing.unexpected_keys
        assert all(".1." in key for key in missing.missing_keys), missing.missing_keys  # BN aliases only
        m.train(train)
        _seen_idx.clear()
        xin = x.permute(0, 2, 1).contiguous().permute(0, 2, 1)   # the callers' non-contiguous view
        logits = m(xin)
        out[f"{name}/logits"] = logits.detach().numpy()
        for li, idx in enumerate(_seen_idx[:3], start=1):
            out[f"{name}/idx{li}"] = idx.numpy().astype(np.int32)
        if train:
            cot = common.cotangent(seed + 200, tuple(logits.shape))
            (logits * cot).sum().backward()
            for pname, prm in m.named_parameters():
                _store_grad(out, f"{name}/{pname}", prm.grad)
            for i in range(1, 9):
                bn = getattr(m, f"bn{i}")
                out[f"{name}/bn{i}.running_mean"] = bn.running_mean.numpy().copy()
                out[f"{name}/bn{i}.running_var"] = bn.running_var.numpy().copy()
                out[f"{name}/bn{i}.num_batches_tracked"] = np.array(int(bn.num_batches_tracked))
        else:
            with torch.no_grad():
                for d in range(1, 10):
                    y = m(xin, depth=d)
                    out[f"{name}/depth{d}.shape"] = np.array(y.shape)
                    out[f"{name}/depth{d}.norm"] = np.array(float(y.double().norm()))
                    y = y.reshape(-1)
                    out[f"{name}/depth{d}"] = (y if y.numel() <= BIG else y[::37]).numpy()
        out[f"{name}/state_keys"] = np.array(sorted(m.state_dict().keys()))
        out[f"{name}/param_order"] = np.array([n for n, _ in m.named_parameters()])
    np.savez_compressed(os.path.join(HERE, "model.npz"), **out)
    for f in ("knn.npz", "graph_feature.npz", "model.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
