"""CUDA kernels of the message layer's first tensor product by linearity (csrc/msg_table.cu) against the CPU restatement
(oracle/msg_table_oracle.py, itself pinned to the reference-pinned port in tests/test_oracle_msg_table.py), through the
C ABI; plus the graph helpers (CSR row pointers of a sorted index, stable transposed edge order) bit-exact vs numpy."""
import numpy as np
import pytest
import torch

from oracle import msg_table_oracle as M
from oracle.segnn_oracle import SIGMOID_CST, SILU_CST

pytestmark = pytest.mark.gpu


def _graph(rng, n_dst, n_all, e):
    dst = np.sort(rng.integers(0, max(n_dst, 1), e)).astype(np.int32)
    src = rng.integers(0, max(n_all, 1), e).astype(np.int32)
    return dst, src


@pytest.mark.parametrize("n_dst,n_all,e", [(50, 50, 0), (1, 1, 7), (200, 260, 4001), (3000, 3000, 60000), (17, 40, 2000)])
def test_edge_index_bit_exact(n_dst, n_all, e):
    from se3gnn_b200.msg import build_edge_index
    rng = np.random.default_rng(e + n_dst)
    dst, src = _graph(rng, n_dst, n_all, e)
    ei = build_edge_index(torch.from_numpy(dst).cuda(), torch.from_numpy(src).cuda(), n_dst, n_all)
    torch.cuda.synchronize()
    rp = np.searchsorted(dst, np.arange(n_dst + 1), side="left")
    np.testing.assert_array_equal(ei.rowptr.cpu().numpy(), rp)
    tp = np.concatenate([[0], np.cumsum(np.bincount(src, minlength=n_all))])
    np.testing.assert_array_equal(ei.tptr.cpu().numpy(), tp)
    if e:
        np.testing.assert_array_equal(ei.perm.cpu().numpy()[:e], np.argsort(src, kind="stable"))


@pytest.mark.parametrize("ns,nv,n_dst,n_all,e", [(34, 10, 300, 300, 5003), (34, 10, 2000, 2100, 40000), (8, 4, 64, 80, 999),
                                                 (16, 8, 500, 500, 7000), (34, 10, 10, 10, 1)])
def test_msg1_kernels_vs_restatement(ns, nv, n_dst, n_all, e):
    from se3gnn_b200 import msg
    rng = np.random.default_rng(ns * 1000 + e)
    torch.manual_seed(e)
    d, mz = ns + 3 * nv, ns + nv
    rows = 2 * ns + 2 + 2 * nv
    dst, src = _graph(rng, n_dst, n_all, e)
    x = torch.randn(n_all, d, dtype=torch.float64)
    wz = torch.rand(rows, mz, dtype=torch.float64) * 2 - 1
    wv = torch.rand(rows, nv, dtype=torch.float64) * 2 - 1
    nz = torch.full((mz,), 0.31, dtype=torch.float64)
    nvn = torch.full((3 * nv,), 0.17, dtype=torch.float64)
    y = torch.randn(e, 4, dtype=torch.float64)
    extra = torch.randn(e, 2, dtype=torch.float64)
    gpost = torch.randn(e, d, dtype=torch.float64)
    dt, st = torch.from_numpy(dst), torch.from_numpy(src)
    post_r, gx_r, gwz_r, gwv_r = M.msg1_forward_backward(ns, nv, x, wz, wv, nz, nvn, y, extra, dt, st, gpost,
                                                         SILU_CST, SIGMOID_CST)
    c = lambda t: t.float().cuda().contiguous()
    xg, wzg, wvg = c(x).requires_grad_(), c(wz).requires_grad_(), c(wv).requires_grad_()
    ei = msg.build_edge_index(dt.cuda(), st.cuda(), n_dst, n_all)
    post = msg.msg1(xg, wzg, wvg, c(nz), c(nvn), c(y), c(extra), ei, ns, nv, SILU_CST, SIGMOID_CST)
    (post * c(gpost)).sum().backward()
    rel = lambda a, b: float((a.detach().cpu().double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert rel(post, post_r) < 1e-5
    assert rel(xg.grad, gx_r) < 1e-5
    assert rel(wzg.grad, gwz_r) < 2e-5
    assert rel(wvg.grad, gwv_r) < 2e-5


def _gate(raw, ns, nv):
    e = raw.shape[0]
    s, g, v = raw[:, :ns], raw[:, ns:ns + nv], raw[:, ns + nv:].reshape(e, nv, 3)
    return torch.cat([SILU_CST * torch.nn.functional.silu(s), (SIGMOID_CST * torch.sigmoid(g)[:, :, None] * v).reshape(e, -1)], 1)


@pytest.mark.parametrize("ns,nv,n_dst,n_all,e", [(34, 10, 300, 300, 5003), (34, 10, 2000, 2100, 40000), (16, 8, 500, 500, 7000),
                                                 (34, 10, 10, 10, 1), (34, 10, 40, 40, 64), (34, 10, 5000, 5000, 64 * 148 * 3 + 17)])
def test_fused_message_layer_vs_port(ns, nv, n_dst, n_all, e):
    """The fused tcgen05 message layer (tables -> message 1 -> message 2 -> segment sum) and its backward against the
    reference-pinned tensor-product port applied twice with the gate in between (fp64, autograd)."""
    from oracle.l1tp_port import L1TPPort
    from se3gnn_b200 import capi, msg
    from se3gnn_b200.irreps import Irreps
    from se3gnn_b200.tp import get_plan
    rng = np.random.default_rng(ns * 1000 + e)
    torch.manual_seed(e)
    d = ns + 3 * nv
    h, hg = f"{ns}x0e+{nv}x1o", f"{ns + nv}x0e+{nv}x1o"
    tp1, tp2 = L1TPPort(f"{h}+{h}+2x0e", hg).double(), L1TPPort(h, hg).double()
    dst, src = _graph(rng, n_dst, n_all, e)
    dt, st = torch.from_numpy(dst), torch.from_numpy(src)
    x = torch.randn(n_all, d, dtype=torch.float64, requires_grad=True)
    y = torch.randn(e, 4, dtype=torch.float64)
    extra = torch.randn(e, 2, dtype=torch.float64)
    gagg = torch.randn(n_dst, d, dtype=torch.float64)
    m1 = _gate(tp1(torch.cat([x[dt.long()], x[st.long()], extra], 1), y), ns, nv)
    m2 = _gate(tp2(m1, y), ns, nv)
    agg_r = torch.zeros(n_dst, d, dtype=torch.float64).index_add(0, dt.long(), m2)
    (agg_r * gagg).sum().backward()

    c = lambda t: t.detach().float().cuda().contiguous()
    xg = c(x).requires_grad_()
    w1 = (c(tp1.weights_l0e).requires_grad_(), c(tp1.weights_l1o).requires_grad_())
    w2 = (c(tp2.weights_l0e).requires_grad_(), c(tp2.weights_l1o).requires_grad_())
    n1, n2 = (c(tp1.norm_l0e), c(tp1.norm_l1o)), (c(tp2.norm_l0e), c(tp2.norm_l1o))
    ei = msg.build_edge_index(dt.cuda(), st.cuda(), n_dst, n_all)
    t0 = capi.tc_launch_count()
    agg = msg.message_layer(xg, w1, n1, w2, n2, c(y), c(extra), ei, ns, nv, SILU_CST, SIGMOID_CST,
                            get_plan(Irreps(h), Irreps(hg)))
    (agg * c(gagg)).sum().backward()
    torch.cuda.synchronize()
    assert capi.tc_launch_count() - t0 >= 1
    rel = lambda a, b: float((a.detach().cpu().double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert rel(agg, agg_r.detach()) < 1e-5
    assert rel(xg.grad, x.grad) < 2e-5
    for wg, ref in ((w1[0], tp1.weights_l0e), (w1[1], tp1.weights_l1o), (w2[0], tp2.weights_l0e), (w2[1], tp2.weights_l1o)):
        assert rel(wg.grad, ref.grad) < 5e-5
