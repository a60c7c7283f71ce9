"""The C-ABI library loads and exports every symbol include/se3gnn_b200.h declares (no compute)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "se3gnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(se3_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    from se3gnn_b200 import capi
    L = ctypes.CDLL(capi.LIB_PATH)
    names = _declared()
    assert "se3_l1tp_forward" in names and "se3_l1tp_backward" in names
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert sorted(e[0] for e in capi.EXPORTS) == names
    assert L.se3_version() >= 100
