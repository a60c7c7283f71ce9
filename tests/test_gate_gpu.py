"""The stand-alone gate kernels (`se3_gate_forward/backward`, csrc/gate.cu) against the defining formula in fp64
(public SEGNN Gate: normalize2mom silu on the scalars, normalize2mom sigmoid gates on every l > 0 channel)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(raw, ns, blocks):
    from se3gnn_b200.gate import SIGMOID_CST, SILU_CST
    ng = sum(c for c, _ in blocks)
    g = SIGMOID_CST * torch.sigmoid(raw[:, ns:ns + ng])
    out = [SILU_CST * torch.nn.functional.silu(raw[:, :ns])]
    o, k = ns + ng, 0
    for c, d in blocks:
        out.append((raw[:, o:o + c * d].reshape(-1, c, d) * g[:, k:k + c, None]).reshape(len(raw), -1))
        o += c * d
        k += c
    return torch.cat(out, 1)


@pytest.mark.parametrize("rows,ns,blocks", [(1000, 23, [(7, 3), (4, 5)]), (37, 5, [(2, 3)]), (513, 0, [(3, 5), (1, 3), (2, 1)]),
                                            (64, 9, []), (1, 1, [(1, 3), (1, 5), (1, 3), (1, 5)])])
def test_gate_matches_formula(rows, ns, blocks):
    from se3gnn_b200 import capi
    from se3gnn_b200.gate import irreps_gate
    d_raw = ns + sum(c for c, _ in blocks) + sum(c * d for c, d in blocks)
    torch.manual_seed(rows)
    raw = (2.0 * torch.randn(rows, d_raw, dtype=torch.float64)).requires_grad_()
    want = _ref(raw, ns, blocks)
    cot = torch.randn(want.shape, dtype=torch.float64)
    (want * cot).sum().backward()
    x = raw.detach().float().cuda().requires_grad_()
    n0 = capi.launch_count()
    got = irreps_gate(x, ns, blocks)
    (got * cot.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    assert capi.launch_count() - n0 == 2
    assert got.shape == want.shape
    assert (got.detach().cpu().double() - want.detach()).abs().max() <= 1e-5 * max(1.0, want.detach().abs().max().item())
    assert (x.grad.cpu().double() - raw.grad).abs().max() <= 1e-5 * max(1.0, raw.grad.abs().max().item())


@pytest.mark.parametrize("n_seg,rows,ns,blocks", [(40, 700, 23, [(7, 3), (4, 5)]), (5, 9, 5, [(2, 3)]), (300, 200, 9, []),
                                                   (17, 400, 40, [(30, 3), (20, 5)])])
def test_gate_segment_sum_matches_formula(n_seg, rows, ns, blocks):
    """gate + aggregation over sorted destinations in one kernel (`se3_gate_segment_sum_*`) == index_add of the gated rows
    in fp64; segments without rows give zero rows."""
    from se3gnn_b200 import capi
    from se3gnn_b200.gate import irreps_gate_segment_sum
    d_raw = ns + sum(c for c, _ in blocks) + sum(c * d for c, d in blocks)
    g = torch.Generator().manual_seed(rows)
    seg = torch.randint(0, n_seg, (rows,), generator=g).sort().values
    seg[seg == 2] = 3                       # an empty segment
    rowptr = torch.zeros(n_seg + 1, dtype=torch.int64)
    rowptr[1:] = torch.bincount(seg, minlength=n_seg).cumsum(0)
    raw = (2.0 * torch.randn(rows, d_raw, generator=g, dtype=torch.float64)).requires_grad_()
    gated = _ref(raw, ns, blocks)
    want = torch.zeros(n_seg, gated.shape[1], dtype=torch.float64).index_add_(0, seg, gated)
    cot = torch.randn(want.shape, generator=g, dtype=torch.float64)
    (want * cot).sum().backward()
    x = raw.detach().float().cuda().requires_grad_()
    n0 = capi.launch_count()
    got = irreps_gate_segment_sum(x, ns, blocks, seg.int().cuda(), rowptr.cuda(), n_seg)
    (got * cot.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    assert capi.launch_count() - n0 == 2
    assert (got.detach().cpu().double() - want.detach()).abs().max() <= 1e-5 * max(1.0, want.detach().abs().max().item())
    assert (x.grad.cpu().double() - raw.grad).abs().max() <= 1e-5 * max(1.0, raw.grad.abs().max().item())
