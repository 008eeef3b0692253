"""B200-native EdgeConv hot path of DGCNN (drop-in for BenCurran98/AHN3-dgcnn.pytorch `model.py`).

The directory name (`ahn3-dgcnn.pytorch_b200`) is not a Python identifier; load it with
`__graft_entry__.load_package()` (importlib, module name `ahn3_dgcnn_pytorch_b200`), or put this
directory on `sys.path` and `from model import DGCNN, knn, get_graph_feature` exactly like the reference's
drivers do (train.py:12, test.py:11): `model.py` bootstraps the package when it is imported top-level
(`tests/test_dropin.py` executes the reference call sites that way).
"""
from . import _lib, ops  # noqa: F401
from .model import DGCNN, get_graph_feature, knn  # noqa: F401

__all__ = ["DGCNN", "knn", "get_graph_feature", "ops", "_lib"]
