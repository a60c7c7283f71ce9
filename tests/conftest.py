import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scalable-e3-gnn_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(path):
    z = np.load(path)
    rec = {k: z[k] for k in z.files if k != "meta"}
    rec["meta"] = json.loads(bytes(z["meta"]).decode())
    return rec


def golden_l1tp_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "l1tp_*.npz")))


@pytest.fixture(scope="session")
def has_cuda():
    import torch
    return torch.cuda.is_available()
