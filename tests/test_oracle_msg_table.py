"""The "msg1 by linearity" formulation (oracle/msg_table_oracle.py, the CPU restatement of csrc/msg_table.cu) against the
reference-pinned tensor product port + gate under autograd: forward, input gradient and both weight gradients."""
import numpy as np
import pytest
import torch

from oracle import msg_table_oracle as M
from oracle.l1tp_port import L1TPPort
from oracle.segnn_oracle import SIGMOID_CST, SILU_CST


@pytest.mark.parametrize("ns,nv", [(34, 10), (8, 4), (16, 8)])
def test_table_formulation_equals_port(ns, nv):
    torch.manual_seed(ns)
    rng = np.random.default_rng(nv)
    n, e = 23, 301
    h = f"{ns}x0e+{nv}x1o"
    hg = f"{ns + nv}x0e+{nv}x1o"
    tp = L1TPPort(f"{h}+{h}+2x0e", hg).double()
    # per-channel norms that differ between irreps (the real module's are uniform inside an irrep)
    x = torch.randn(n, ns + 3 * nv, dtype=torch.float64, requires_grad=True)
    dst = torch.from_numpy(np.sort(rng.integers(0, n, e))).int()
    src = torch.from_numpy(rng.integers(0, n, e)).int()
    y = torch.randn(e, 4, dtype=torch.float64)
    extra = torch.randn(e, 2, dtype=torch.float64)
    gpost = torch.randn(e, ns + 3 * nv, dtype=torch.float64)

    raw = tp(torch.cat([x[dst.long()], x[src.long()], extra], 1), y)
    s, g, v = raw[:, :ns], raw[:, ns:ns + nv], raw[:, ns + nv:].reshape(e, nv, 3)
    post_ref = torch.cat([SILU_CST * torch.nn.functional.silu(s),
                          (SIGMOID_CST * torch.sigmoid(g)[:, :, None] * v).reshape(e, -1)], 1)
    (post_ref * gpost).sum().backward()

    with torch.no_grad():
        post, gx, gwz, gwv = M.msg1_forward_backward(ns, nv, x.detach(), tp.weights_l0e.detach(), tp.weights_l1o.detach(),
                                                     tp.norm_l0e.double(), tp.norm_l1o.double(), y, extra, dst, src,
                                                     gpost, SILU_CST, SIGMOID_CST)
    close = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert close(post, post_ref.detach()) < 1e-12
    assert close(gx, x.grad) < 1e-12
    assert close(gwz, tp.weights_l0e.grad) < 1e-12
    assert close(gwv, tp.weights_l1o.grad) < 1e-12
