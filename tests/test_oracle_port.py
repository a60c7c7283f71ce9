"""Pin the torch CPU port (oracle/l1tp_port.py) against the reference's golden vectors, and against
the unmodified reference itself when /root/reference is mounted (build container only)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_l1tp_files, load_golden
from oracle.l1tp_port import L1TPPort


@pytest.mark.parametrize("path", golden_l1tp_files(), ids=lambda p: os.path.basename(p)[5:-4])
def test_port_matches_golden(path):
    rec = load_golden(path)
    meta = rec["meta"]
    tp = L1TPPort(meta["in1"], meta["out"], **meta["kwargs"]).double()
    sd = {k[3:]: torch.from_numpy(rec[k]).double() for k in rec if k.startswith("sd_")}
    tp.load_state_dict(sd)
    x = torch.from_numpy(rec["x"]).double().requires_grad_(True)
    y = torch.from_numpy(rec["y"]).double().requires_grad_(True)
    o = tp(x, y)
    np.testing.assert_allclose(o.detach().numpy(), rec["out_f64"], rtol=1e-10, atol=1e-12)
    o.backward(torch.from_numpy(rec["gout"]).double())
    np.testing.assert_allclose(x.grad.numpy(), rec["gx_f64"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(y.grad.numpy(), rec["gy_f64"], rtol=1e-10, atol=1e-11)
    for k, p in tp.named_parameters():
        np.testing.assert_allclose(p.grad.numpy(), rec[f"gw_{k}_f64"], rtol=1e-10, atol=1e-11)


@pytest.mark.skipif(not os.path.exists("/root/reference/models/segnn/l1_tensor_prod.py"),
                    reason="reference mount absent (GPU box)")
def test_port_matches_live_reference():
    import importlib.util
    sys.path.insert(0, os.path.join(ROOT, "oracle", "e3nn_shim"))
    try:
        spec = importlib.util.spec_from_file_location("_ref_l1tp", "/root/reference/models/segnn/l1_tensor_prod.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        from e3nn.o3 import Irreps
    finally:
        sys.path.pop(0)
    torch.manual_seed(3)
    in1, out = "5x0e+2x0o+3x1e+4x1o", "3x0e+1x0o+2x1e+2x1o"
    ref = mod.L1TensorProduct(Irreps(in1), Irreps(out))
    port = L1TPPort(in1, out)
    port.load_state_dict(ref.state_dict())
    x, y = torch.randn(50, ref.in1_dim), torch.randn(50, 4)
    assert torch.allclose(ref(x, y), port(x, y), rtol=1e-6, atol=1e-6)
