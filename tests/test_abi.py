"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/dgcnn_b200.h declares, validates arguments before touching the GPU, and the Python host mirrors
the reference's module surface.  No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    entry.build()
    return entry.load_package()


def _declared():
    text = open(os.path.join(ROOT, "include", "dgcnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dgcnn_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(pkg):
    names = _declared()
    assert len(names) >= 18
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(pkg._lib.SIGNATURES) == names            # the ctypes binding covers the whole header


def test_version_and_argument_validation(pkg):
    lib = pkg._lib.load()
    assert lib.dgcnn_version() == 100
    assert lib.dgcnn_edgeconv_partials_bytes() > 0 and lib.dgcnn_edgeconv_dw_partials_bytes() > 0
    # null pointers / k > N are rejected on the host before any CUDA call
    assert lib.dgcnn_knn_f32(None, 4, 1, 4, 10, 3, None, None, None) == -1
    assert b"null" in lib.dgcnn_last_error()
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.dgcnn_knn_f32(p, 4, 1, 4, 10, 11, p, p, None) == -1
    assert b"out of range" in lib.dgcnn_last_error()
    assert lib.dgcnn_knn_f32(p, 256, 1, 200, 10, 3, p, p, None) == -2      # C > 128 not compiled
    with pytest.raises(pkg._lib.DgcnnError):
        pkg._lib.check(-1)


def test_module_surface_matches_reference_contract(pkg):
    m = pkg.DGCNN(5, 4, 20)
    sd = m.state_dict()
    assert len(sd) == 89
    assert sd["bn1.weight"].data_ptr() == sd["conv1.1.weight"].data_ptr()         # BN aliases share storage
    names = [n for n, _ in m.named_parameters()]
    assert len(names) == 25 and names[0] == "bn1.weight" and names[-1] == "conv9.weight"
    assert names[16:] == [f"conv{i}.0.weight" for i in range(1, 9)] + ["conv9.weight"]
    assert sum(p.numel() for p in m.parameters()) == 980864
    assert tuple(sd["conv1.0.weight"].shape) == (64, 8, 1, 1) and tuple(sd["conv7.0.weight"].shape) == (512, 1216, 1)
    # reference train.py:103 passes the mistyped keyword `ropout`
    assert pkg.DGCNN(5, 4, 20, ropout=0.3).dp1.p == 0.3
    with pytest.raises(TypeError):
        pkg.DGCNN(5, 4, 20, bogus=1)


def test_no_cpu_fallback(pkg):
    m = pkg.DGCNN(5, 4, 20)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 4, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        pkg.knn(torch.rand(1, 4, 64), 5)


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "ahn3-dgcnn.pytorch_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"
