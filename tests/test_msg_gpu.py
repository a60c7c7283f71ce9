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


@pytest.mark.parametrize("ns,nv,n_dst,n_all,e", [(34, 10, 300, 300, 5003), (34, 10, 2000, 2100, 40000), (6, 3, 64, 80, 999),
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
